// Matching-cost volume (pf:78-113):  L[d,h,w] = -<fl[h,w,:], fr[h,w-d,:]>  (w >= d),
// R[d,h,w] = L[d,h,w+d]  (w < W-d), the invalid triangles filled by the 3-tap mean recurrence
// (pf:94-95, :105-106).  Volumes are written in the HWD layout.
//
// k_cost_volume: one CTA per (row h, 32-pixel tile).  The 64-channel feature rows of the tile
// (left) and of the 32+D pixels it can match (right) are staged once in shared memory with
// 16-byte cp.async copies (XOR-swizzled so that the float4 reads below are conflict free).  Each
// thread owns a 4(w) x 8(d) register tile: along the diagonals w-d the right-image pixel is the
// same, so 11 right pixels feed 32 outputs.  Every product is computed once and stored twice:
// L directly from registers (two 16 B stores per pixel), R through a shared-memory transpose so
// that each warp writes a contiguous run of disparities of one right-image pixel.
#include "common.cuh"

namespace mccnn {

constexpr int CV_TWB = 32;      // pixels per CTA tile
constexpr int CV_C = 64;        // feature channels (model.py:38)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// float4 chunk c4 (0..15) of staged pixel r lives at chunk (c4 ^ ((r >> 2) & 7)) of row r.
__device__ __forceinline__ int swz(int r, int c4) { return r * 16 + (c4 ^ ((r >> 2) & 7)); }

__global__ void __launch_bounds__(512) k_cost_volume(const float *__restrict__ fl, const float *__restrict__ fr,
                                                     float *__restrict__ L, float *__restrict__ R, int H, int W, int D,
                                                     int Dp, int Dr) {
    extern __shared__ float4 smem4[];
    float4 *As = smem4;                         // [32][16] float4
    float4 *Bs = smem4 + CV_TWB * 16;           // [32 + Dr][16] float4 ; later reused as the output tile
    const int h = blockIdx.y, w0 = blockIdx.x * CV_TWB;
    const int xb0 = w0 - Dr;                    // first staged right-image pixel
    const int NB = CV_TWB + Dr;
    const int tid = threadIdx.x, nthr = blockDim.x;

    // ---- stage the feature rows
    const float4 *flrow = reinterpret_cast<const float4 *>(fl) + (size_t)h * W * 16;
    const float4 *frrow = reinterpret_cast<const float4 *>(fr) + (size_t)h * W * 16;
    for (int i = tid; i < CV_TWB * 16; i += nthr) {
        int r = i >> 4, c4 = i & 15, w = w0 + r;
        if (w < W) cp_async16(&As[swz(r, c4)], flrow + (size_t)w * 16 + c4);
        else As[swz(r, c4)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < NB * 16; i += nthr) {
        int r = i >> 4, c4 = i & 15, x = xb0 + r;
        if (x >= 0 && x < W) cp_async16(&Bs[swz(r, c4)], frrow + (size_t)x * 16 + c4);
        else Bs[swz(r, c4)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- 4 x 8 register tile: pixels w0 + 4wq + i, disparities 8dq + j
    const int wq = tid & 7, dq = tid >> 3;
    const int d0 = dq * 8;
    const bool active = d0 < Dr;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
    if (active) {
        const int ra = 4 * wq;                      // first left pixel (tile relative)
        const int rb = 4 * wq - d0 + Dr - 7;        // first right pixel: (w - d) for i = 0, j = 7
#pragma unroll 2
        for (int c4 = 0; c4 < 16; c4++) {
            float4 a[4], b[11];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[swz(ra + i, c4)];
#pragma unroll
            for (int k = 0; k < 11; k++) b[k] = Bs[swz(rb + k, c4)];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 y = b[i - j + 7];
                    acc[i][j] = fmaf(a[i].x, y.x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, y.y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, y.z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, y.w, acc[i][j]);
                }
        }
    }
    __syncthreads();                                // everyone is done reading Bs

    // ---- L straight from registers; the tile (negated, pf:111) also goes to shared memory for R
    const int SP = Dr + 4;                          // row pitch of the output tile in floats
    float *Ss = reinterpret_cast<float *>(Bs);
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int w = w0 + 4 * wq + i;
            float4 lo = make_float4(-acc[i][0], -acc[i][1], -acc[i][2], -acc[i][3]);
            float4 hi = make_float4(-acc[i][4], -acc[i][5], -acc[i][6], -acc[i][7]);
            float4 *srow = reinterpret_cast<float4 *>(Ss + (4 * wq + i) * SP + d0);
            srow[0] = lo;
            srow[1] = hi;
            if (w < W) {
                // cells with w < d are overwritten by k_cost_fill afterwards
                float4 *lrow = reinterpret_cast<float4 *>(L + ((size_t)h * W + w) * Dp + d0);
                if (d0 < Dp) lrow[0] = lo;
                if (d0 + 4 < Dp) lrow[1] = hi;
            }
        }
    }
    __syncthreads();

    // ---- R[h][x][d] = S(x + d, d): one warp per right-image pixel x, lanes over the run of d whose
    //      left pixel x + d lies in this tile
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int wend = min(w0 + CV_TWB, W);           // exclusive
    for (int x = max(w0 - (D - 1), 0) + warp; x < wend; x += nwarps) {
        const int dlo = max(0, w0 - x);
        const int dhi = min(D - 1, wend - 1 - x);
        const int d = dlo + lane;
        if (d <= dhi) R[((size_t)h * W + x) * Dp + d] = Ss[(x + d - w0) * SP + d];
    }
}

// Invalid triangles (pf:94-95 for L, pf:105-106 for R), in the already negated domain (negation
// commutes exactly with the mean).  One warp per (row, 32 disparities); lanes over d so that the
// cells written at each step are contiguous; each lane slides a 3-value window along w.
__global__ void k_cost_fill(float *__restrict__ L, float *__restrict__ R, int H, int W, int D, int Dp) {
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.y;
    const int d0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (d0 >= D) return;
    const int d = d0 + lane;
    const bool live = d < D && d >= 1;
    const int dmax = min(d0 + 31, D - 1);
    float *Lrow = L + (size_t)h * W * Dp;
    float *Rrow = R + (size_t)h * W * Dp;
    {   // L: columns d-1 .. 0, right to left
        float v1 = 0.f, v2 = 0.f, v3 = 0.f;          // values at c+1, c+2, c+3
        for (int c = dmax - 1; c >= 0; c--) {
            if (live && c == d - 1) {
                v1 = Lrow[(size_t)(c + 1) * Dp + d];
                v2 = Lrow[(size_t)(c + 2) * Dp + d];
                v3 = Lrow[(size_t)(c + 3) * Dp + d];
            }
            if (live && c <= d - 1) {
                float v = ((v1 + v2) + v3) / 3.0f;
                Lrow[(size_t)c * Dp + d] = v;
                v3 = v2; v2 = v1; v1 = v;
            }
        }
    }
    {   // R: columns W-d .. W-1, left to right
        float v1 = 0.f, v2 = 0.f, v3 = 0.f;          // values at c-1, c-2, c-3
        for (int c = W - dmax; c < W; c++) {
            if (live && c == W - d) {
                v1 = Rrow[(size_t)(c - 1) * Dp + d];
                v2 = Rrow[(size_t)(c - 2) * Dp + d];
                v3 = Rrow[(size_t)(c - 3) * Dp + d];
            }
            if (live && c >= W - d) {
                float v = ((v3 + v2) + v1) / 3.0f;
                Rrow[(size_t)c * Dp + d] = v;
                v3 = v2; v2 = v1; v1 = v;
            }
        }
    }
}

}  // namespace mccnn

using namespace mccnn;

extern "C" {

int mccnn_cost_volume(const float *fl, const float *fr, float *L, float *R, int H, int W, int C, int D, void *stream) {
    MCCNN_REQUIRE(fl && fr && L && R, "cost_volume: null pointer");
    MCCNN_REQUIRE(C == CV_C, "cost_volume: %d feature channels unsupported (the network emits 64, model.py:38)", C);
    MCCNN_REQUIRE(H >= 1 && D >= 1 && W >= D + 2, "cost_volume: need W >= ndisp + 2 (pf:94-95), got W=%d ndisp=%d", W, D);
    MCCNN_REQUIRE(D <= 512, "cost_volume: ndisp %d too large (max 512)", D);
    cudaStream_t s = (cudaStream_t)stream;
    const int Dp = dpitch(D);
    const int Dr = ((D + 7) / 8) * 8;
    int threads = ((8 * (Dr / 8) + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    const size_t smem = (size_t)(CV_TWB + CV_TWB + Dr) * 16 * sizeof(float4);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        MCCNN_CUDA(cudaFuncSetAttribute(k_cost_volume, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    dim3 grid(cdiv(W, CV_TWB), H);
    k_cost_volume<<<grid, threads, smem, s>>>(fl, fr, L, R, H, W, D, Dp, Dr);
    MCCNN_LAUNCHED("cost_volume");
    if (D > 1) {
        dim3 fgrid(cdiv(cdiv(D, 32), 4), H);
        k_cost_fill<<<fgrid, 128, 0, s>>>(L, R, H, W, D, Dp);
        MCCNN_LAUNCHED("cost_fill");
    }
    return MCCNN_OK;
}

}  // extern "C"
