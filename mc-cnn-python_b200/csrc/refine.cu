// WTA + refinement kernels (pf:239-470).  Compiled with -fmad=false: every float32 operation is
// separately rounded, which makes all five stages bit-exact against the reference's NumPy code.
#include "common.cuh"
#include <math_constants.h>

namespace mccnn {

// ------------------------------------------------------------------------------------------
// a8  disparity_prediction (pf:245-254): lowest d among the minima, stored as float32.
// HWD layout: GS lanes share one pixel, each lane walks float4 granules g = lane + GS*j, keeps its
// first minimum (strict <), then the group merges lexicographically on (value, d).
// HBM-bound: 4 B per cell read once, fully coalesced (a warp reads 512 contiguous bytes).
// ------------------------------------------------------------------------------------------
template <int GS, int NL>
__global__ void __launch_bounds__(256) k_wta(const float *__restrict__ vol, float *__restrict__ disp,
                                             float *__restrict__ minval, int dbase, int D, int Dp, long long P) {
    const int lane_in_group = threadIdx.x % GS;
    long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / GS;
    const bool live = p < P;
    float best = CUDART_INF_F;
    int bd = -1;
    if (live) {
        const float4 *row = reinterpret_cast<const float4 *>(vol + p * Dp);
        const int G = Dp >> 2;
        // all of this lane's granules first (NL independent 16-byte loads in flight), then the compares
        float4 v[NL];
#pragma unroll
        for (int i = 0; i < NL; i++) {
            const int g = lane_in_group + i * GS;
            v[i] = g < G ? __ldcs(row + g) : make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
        }
#pragma unroll
        for (int i = 0; i < NL; i++) {
            const int d = (lane_in_group + i * GS) << 2;
            if (d + 0 < D && v[i].x < best) { best = v[i].x; bd = d; }
            if (d + 1 < D && v[i].y < best) { best = v[i].y; bd = d + 1; }
            if (d + 2 < D && v[i].z < best) { best = v[i].z; bd = d + 2; }
            if (d + 3 < D && v[i].w < best) { best = v[i].w; bd = d + 3; }
        }
        for (int g = lane_in_group + NL * GS; g < G; g += GS) {                // (only for ndisp beyond NL * GS * 4)
            float4 w = __ldcs(row + g);
            int d = g << 2;
            if (d + 0 < D && w.x < best) { best = w.x; bd = d; }
            if (d + 1 < D && w.y < best) { best = w.y; bd = d + 1; }
            if (d + 2 < D && w.z < best) { best = w.z; bd = d + 2; }
            if (d + 3 < D && w.w < best) { best = w.w; bd = d + 3; }
        }
    }
#pragma unroll
    for (int off = GS / 2; off >= 1; off >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, off);
        int od = __shfl_xor_sync(0xffffffffu, bd, off);
        bool take = (od >= 0) && (bd < 0 || ov < best || (ov == best && od < bd));
        if (take) { best = ov; bd = od; }
    }
    if (live && lane_in_group == 0) {
        disp[p] = (float)(dbase + bd);                      // (a disparity slab reports the pair's disparity)
        if (minval) minval[p] = best;
    }
}

// ------------------------------------------------------------------------------------------
// a9  interpolation (pf:279-378), pass 1: consistency labels (pf:285-307).
// ------------------------------------------------------------------------------------------
__global__ void k_lr_labels(const float *__restrict__ dl, const float *__restrict__ dr, int32_t *__restrict__ lab,
                            int H, int W, int ndisp) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int h = blockIdx.y;
    if (w >= W) return;
    size_t row = (size_t)h * W;
    int ld = (int)dl[row + w];                                   // pf:287 int() truncation
    int label;
    if (w < ld) {
        label = 2;
    } else {
        int x = w - ld;
        x = x < 0 ? 0 : (x >= W ? W - 1 : x);                    // memory safety only (ld < 0 is outside the reference's domain)
        float rd = dr[row + x];
        if (fabsf((float)ld - rd) <= 1.0f) {
            label = 0;                                           // pf:294
        } else {
            int lim = (w + 1 < ndisp) ? w + 1 : ndisp;
            label = 2;
            for (int d = 0; d < lim; d++)
                if (fabsf((float)d - dr[row + (w - d)]) <= 1.0f) { label = 1; break; }
        }
    }
    lab[row + w] = label;
}

__device__ __forceinline__ float np_median_small(float *v, int n) {
    for (int i = 0; i < n; i++)
        if (isnan(v[i])) return CUDART_NAN_F;
    for (int i = 1; i < n; i++) {
        float x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; j--; }
        v[j + 1] = x;
    }
    if (n & 1) return v[n / 2];
    return (v[n / 2 - 1] + v[n / 2]) / 2.0f;
}

// pass 2: fill (pf:312-373), always reading the ORIGINAL left map and labels.
__global__ void k_lr_fill(const float *__restrict__ dl, const int32_t *__restrict__ lab, float *__restrict__ out,
                          int H, int W) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int h = blockIdx.y;
    if (w >= W) return;
    size_t p = (size_t)h * W + w;
    int l = lab[p];
    float v = dl[p];
    if (l == 1) {
        float nb[4];
        int cnt = 0;
        for (int x = w + 1; x < W; x++) if (lab[(size_t)h * W + x] == 0) { nb[cnt++] = dl[(size_t)h * W + x]; break; }
        for (int x = w - 1; x >= 0; x--) if (lab[(size_t)h * W + x] == 0) { nb[cnt++] = dl[(size_t)h * W + x]; break; }
        for (int y = h + 1; y < H; y++) if (lab[(size_t)y * W + w] == 0) { nb[cnt++] = dl[(size_t)y * W + w]; break; }
        for (int y = h - 1; y >= 0; y--) if (lab[(size_t)y * W + w] == 0) { nb[cnt++] = dl[(size_t)y * W + w]; break; }
        if (cnt) v = np_median_small(nb, cnt);                   // pf:353-356
    } else if (l == 2) {
        for (int x = w + 1; x < W; x++) if (lab[(size_t)h * W + x] == 0) { v = dl[(size_t)h * W + x]; break; }   // pf:365-373
    }
    out[p] = v;
}

// ------------------------------------------------------------------------------------------
// a10  subpixel_enhance (pf:381-400).  The three cells are adjacent in the HWD layout.
// ------------------------------------------------------------------------------------------
__global__ void k_subpixel(const float *__restrict__ disp, const float *__restrict__ vol, float *__restrict__ out,
                           int D, int Dp, long long P) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float d = disp[p];
    int im = (int)(d - 1.0f), ip = (int)(d + 1.0f), ic = (int)d;   // int() truncates toward zero
    if (im < 0 || ip >= D || !(d == d)) { out[p] = d; return; }
    ic = ic < 0 ? 0 : (ic >= D ? D - 1 : ic);
    const float *row = vol + p * Dp;
    float Cm = row[im], Cp = row[ip], C = row[ic];
    float num = Cp - Cm;
    float den = 2.0f * ((Cp - 2.0f * C) + Cm);
    out[p] = d - num / den;                                        // pf:396 (IEEE division; inf/NaN propagate)
}

// The same on a disparity-slab partition: triple [3][P] = (C[d-1], C[d], C[d+1]) restricted to this slab.
__global__ void k_subpixel_gather(const float *__restrict__ disp, const float *__restrict__ vol, float *__restrict__ triple,
                                  int D, int Dp, int dbase, int ndisp, long long P) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float d = disp[p];
    int im = (int)(d - 1.0f), ip = (int)(d + 1.0f), ic = (int)d;
    float Cm = 0.f, C = 0.f, Cp = 0.f;
    if (!(im < 0 || ip >= ndisp || !(d == d))) {
        ic = ic < 0 ? 0 : (ic >= ndisp ? ndisp - 1 : ic);
        const float *row = vol + p * Dp - dbase;
        if (im >= dbase && im < dbase + D) Cm = row[im];
        if (ic >= dbase && ic < dbase + D) C = row[ic];
        if (ip >= dbase && ip < dbase + D) Cp = row[ip];
    }
    triple[p] = Cm;
    triple[P + p] = C;
    triple[2 * P + p] = Cp;
}

__global__ void k_subpixel_triple(const float *__restrict__ disp, const float *__restrict__ triple, float *__restrict__ out,
                                  int ndisp, long long P) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float d = disp[p];
    const int im = (int)(d - 1.0f), ip = (int)(d + 1.0f);
    if (im < 0 || ip >= ndisp || !(d == d)) { out[p] = d; return; }
    const float Cm = triple[p], C = triple[P + p], Cp = triple[2 * P + p];
    const float num = Cp - Cm;
    const float den = 2.0f * ((Cp - 2.0f * C) + Cm);
    out[p] = d - num / den;                                        // pf:396
}

__global__ void k_wta_combine(const float *__restrict__ minvals, const float *__restrict__ disps, float *__restrict__ out,
                              int nslabs, long long stride, long long P) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float best = CUDART_INF_F, bd = -1.0f;
    for (int s = 0; s < nslabs; s++) {
        const float v = minvals[(size_t)s * stride + p];
        if (v < best) { best = v; bd = disps[(size_t)s * stride + p]; }
    }
    out[p] = bd;
}

// ------------------------------------------------------------------------------------------
// a11  median_filter (pf:403-421): border-clipped window, np.median.
// ------------------------------------------------------------------------------------------
#define MCCNN_MAX_WINDOW 121
// median of 25 NaN-free values held in registers: 99 compare-exchanges (Devillard's opt_med25 selection network,
// checked against np.median), no local memory
__device__ __forceinline__ float median25(float (&p)[25]) {
#define MCCNN_CX(a, b) { const float lo_ = fminf(p[a], p[b]), hi_ = fmaxf(p[a], p[b]); p[a] = lo_; p[b] = hi_; }
    MCCNN_CX(0, 1) MCCNN_CX(3, 4) MCCNN_CX(2, 4) MCCNN_CX(2, 3) MCCNN_CX(6, 7) MCCNN_CX(5, 7) MCCNN_CX(5, 6) MCCNN_CX(9, 10)
    MCCNN_CX(8, 10) MCCNN_CX(8, 9) MCCNN_CX(12, 13) MCCNN_CX(11, 13) MCCNN_CX(11, 12) MCCNN_CX(15, 16) MCCNN_CX(14, 16)
    MCCNN_CX(14, 15) MCCNN_CX(18, 19) MCCNN_CX(17, 19) MCCNN_CX(17, 18) MCCNN_CX(21, 22) MCCNN_CX(20, 22) MCCNN_CX(20, 21)
    MCCNN_CX(23, 24) MCCNN_CX(2, 5) MCCNN_CX(3, 6) MCCNN_CX(0, 6) MCCNN_CX(0, 3) MCCNN_CX(4, 7) MCCNN_CX(1, 7) MCCNN_CX(1, 4)
    MCCNN_CX(11, 14) MCCNN_CX(8, 14) MCCNN_CX(8, 11) MCCNN_CX(12, 15) MCCNN_CX(9, 15) MCCNN_CX(9, 12) MCCNN_CX(13, 16)
    MCCNN_CX(10, 16) MCCNN_CX(10, 13) MCCNN_CX(20, 23) MCCNN_CX(17, 23) MCCNN_CX(17, 20) MCCNN_CX(21, 24) MCCNN_CX(18, 24)
    MCCNN_CX(18, 21) MCCNN_CX(19, 22) MCCNN_CX(8, 17) MCCNN_CX(9, 18) MCCNN_CX(0, 18) MCCNN_CX(0, 9) MCCNN_CX(10, 19)
    MCCNN_CX(1, 19) MCCNN_CX(1, 10) MCCNN_CX(11, 20) MCCNN_CX(2, 20) MCCNN_CX(2, 11) MCCNN_CX(12, 21) MCCNN_CX(3, 21)
    MCCNN_CX(3, 12) MCCNN_CX(13, 22) MCCNN_CX(4, 22) MCCNN_CX(4, 13) MCCNN_CX(14, 23) MCCNN_CX(5, 23) MCCNN_CX(5, 14)
    MCCNN_CX(15, 24) MCCNN_CX(6, 24) MCCNN_CX(6, 15) MCCNN_CX(7, 16) MCCNN_CX(7, 19) MCCNN_CX(13, 21) MCCNN_CX(15, 23)
    MCCNN_CX(7, 13) MCCNN_CX(7, 15) MCCNN_CX(1, 9) MCCNN_CX(3, 11) MCCNN_CX(5, 17) MCCNN_CX(11, 17) MCCNN_CX(9, 17)
    MCCNN_CX(4, 10) MCCNN_CX(6, 12) MCCNN_CX(7, 14) MCCNN_CX(4, 6) MCCNN_CX(4, 7) MCCNN_CX(12, 14) MCCNN_CX(10, 14)
    MCCNN_CX(6, 7) MCCNN_CX(10, 12) MCCNN_CX(6, 10) MCCNN_CX(6, 17) MCCNN_CX(12, 17) MCCNN_CX(7, 17) MCCNN_CX(7, 10)
    MCCNN_CX(12, 18) MCCNN_CX(7, 12) MCCNN_CX(10, 18) MCCNN_CX(12, 20) MCCNN_CX(10, 20) MCCNN_CX(10, 12)
#undef MCCNN_CX
    return p[12];
}

__global__ void k_median(const float *__restrict__ in, float *__restrict__ out, int H, int W, int rh, int rw) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int h = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= W || h >= H) return;
    if (rh == 2 && rw == 2 && h >= 2 && h + 2 < H && w >= 2 && w + 2 < W) {
        // the common case (match.py:172 uses 5x5): a full window in registers
        float p[25];
        bool nan = false;
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
            for (int x = 0; x < 5; x++) {
                const float v = in[(size_t)(h + y - 2) * W + (w + x - 2)];
                p[y * 5 + x] = v;
                nan |= (v != v);
            }
        out[(size_t)h * W + w] = nan ? CUDART_NAN_F : median25(p);
        return;
    }
    float buf[MCCNN_MAX_WINDOW];
    int hs = max(h - rh, 0), he = min(h + rh + 1, H);
    int ws = max(w - rw, 0), we = min(w + rw + 1, W);
    int n = 0;
    for (int y = hs; y < he; y++)
        for (int x = ws; x < we; x++) buf[n++] = in[(size_t)y * W + x];
    out[(size_t)h * W + w] = np_median_small(buf, n);
}

// NumPy's float32 pairwise summation of a contiguous run of n < 128 elements
// (numpy/_core/src/umath/loops_utils.h): n < 8 sequential; else 8 strided accumulators, a fixed
// tree, then the remainder sequentially.  Windows here have n <= 121.
__device__ __forceinline__ float np_pairwise_sum_small(const float *a, int n) {
    if (n < 8) {
        float res = -0.0f;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    float r[8];
    int i;
#pragma unroll
    for (i = 0; i < 8; i++) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    }
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

// ------------------------------------------------------------------------------------------
// a12  bilateral_filter (pf:424-470).
// ------------------------------------------------------------------------------------------
__global__ void k_bilateral(const float *__restrict__ img, const float *__restrict__ in, float *__restrict__ out,
                            const float *__restrict__ table, int H, int W, int fh, int fw, float thr) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int h = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= W || h >= H) return;
    int ch = (fh - 1) / 2, cw = (fw - 1) / 2;
    float wts[MCCNN_MAX_WINDOW], prod[MCCNN_MAX_WINDOW];
    int hs = max(h - ch, 0), he = min(h + ch + 1, H);
    int ws = max(w - cw, 0), we = min(w + cw + 1, W);
    float cur = img[(size_t)h * W + w];
    int n = 0;
    for (int y = hs; y < he; y++)
        for (int x = ws; x < we; x++) {
            float diff = fabsf(img[(size_t)y * W + x] - cur);              // pf:458-459
            float mask = (diff < thr) ? 1.0f : 0.0f;                      // pf:460
            float wt = mask * table[(ch - (h - y)) * fw + (cw - (w - x))];  // pf:449-453, :462
            wts[n] = wt;
            prod[n] = wt * in[(size_t)y * W + x];                         // pf:465
            n++;
        }
    float wsum = np_pairwise_sum_small(wts, n);
    out[(size_t)h * W + w] = np_pairwise_sum_small(prod, n) / wsum;       // pf:466
}

}  // namespace mccnn

using namespace mccnn;

extern "C" {

static int wta_launch(const float *vol, float *disp, float *minval, int dbase, int D, int H, int W, void *stream) {
    MCCNN_REQUIRE(vol && disp && D >= 1 && H >= 1 && W >= 1 && dbase >= 0, "wta: bad arguments");
    long long P = (long long)H * W;
    int Dp = dpitch(D), G = Dp / 4;
    cudaStream_t s = (cudaStream_t)stream;
    const int T = 256;
#define WTA_CASE(GS, NL) k_wta<GS, NL><<<cdiv(P * GS, T), T, 0, s>>>(vol, disp, minval, dbase, D, Dp, P)
    if (G > 64) WTA_CASE(32, 4);
    else if (G >= 12) WTA_CASE(16, 4);
    else if (G >= 6) WTA_CASE(8, 2);
    else if (G >= 3) WTA_CASE(4, 2);
    else WTA_CASE(2, 1);
#undef WTA_CASE
    MCCNN_LAUNCHED("wta");
    return MCCNN_OK;
}

int mccnn_wta(const float *vol, float *disp, int D, int H, int W, void *stream) {
    return wta_launch(vol, disp, nullptr, 0, D, H, W, stream);
}

/* Disparity slab [d_base, d_base + D) of one big pair: the slab's first minimum and its cost. */
int mccnn_wta_slab(const float *vol, float *disp, float *minval, int D, int H, int W, int d_base, void *stream) {
    MCCNN_REQUIRE(minval, "wta_slab: null pointer");
    return wta_launch(vol, disp, minval, d_base, D, H, W, stream);
}

/* First minimum over the slabs, in slab order (strict <: the lowest disparity wins ties, pf:247-252). */
int mccnn_wta_combine(const float *minvals, const float *disps, float *out, int nslabs, long long slab_stride, int H,
                      int W, void *stream) {
    long long P = (long long)H * W;
    MCCNN_REQUIRE(minvals && disps && out && nslabs >= 1 && H >= 1 && W >= 1 && slab_stride >= P, "wta_combine: bad arguments");
    k_wta_combine<<<cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(minvals, disps, out, nslabs, slab_stride, P);
    MCCNN_LAUNCHED("wta_combine");
    return MCCNN_OK;
}

/* Sub-pixel on a disparity-slab partition: every slab contributes the cells of {d-1, d, d+1} it owns (0 for the
 * others, so that a sum over the slabs restores the three cells exactly), then the formula of pf:381-400. */
int mccnn_subpixel_gather(const float *disp, const float *vol, float *triple, int D, int H, int W, int d_base,
                          int ndisp, void *stream) {
    MCCNN_REQUIRE(disp && vol && triple && D >= 1 && H >= 1 && W >= 1 && d_base >= 0 && d_base + D <= ndisp,
                  "subpixel_gather: bad arguments");
    long long P = (long long)H * W;
    k_subpixel_gather<<<cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(disp, vol, triple, D, dpitch(D), d_base, ndisp, P);
    MCCNN_LAUNCHED("subpixel_gather");
    return MCCNN_OK;
}

int mccnn_subpixel_triple(const float *disp, const float *triple, float *out, int ndisp, int H, int W, void *stream) {
    MCCNN_REQUIRE(disp && triple && out && ndisp >= 1 && H >= 1 && W >= 1, "subpixel_triple: bad arguments");
    long long P = (long long)H * W;
    k_subpixel_triple<<<cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(disp, triple, out, ndisp, P);
    MCCNN_LAUNCHED("subpixel_triple");
    return MCCNN_OK;
}

int mccnn_lr_interp(const float *dl, const float *dr, float *out, int32_t *labels, int H, int W, int ndisp,
                    void *stream) {
    MCCNN_REQUIRE(dl && dr && out && labels && H >= 1 && W >= 1 && ndisp >= 1, "lr_interp: bad arguments");
    MCCNN_REQUIRE(out != dl, "lr_interp: out must not alias disp_left (the fill reads the original map, pf:312)");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(128), grid(cdiv(W, 128), H);
    k_lr_labels<<<grid, block, 0, s>>>(dl, dr, labels, H, W, ndisp);
    MCCNN_LAUNCHED("lr_labels");
    k_lr_fill<<<grid, block, 0, s>>>(dl, labels, out, H, W);
    MCCNN_LAUNCHED("lr_fill");
    return MCCNN_OK;
}

int mccnn_subpixel(const float *disp, const float *vol, float *out, int D, int H, int W, void *stream) {
    MCCNN_REQUIRE(disp && vol && out && D >= 1 && H >= 1 && W >= 1, "subpixel: bad arguments");
    long long P = (long long)H * W;
    k_subpixel<<<cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(disp, vol, out, D, dpitch(D), P);
    MCCNN_LAUNCHED("subpixel");
    return MCCNN_OK;
}

int mccnn_median(const float *in, float *out, int H, int W, int fh, int fw, void *stream) {
    MCCNN_REQUIRE(in && out && in != out && H >= 1 && W >= 1, "median: bad arguments");
    MCCNN_REQUIRE(fh >= 1 && fw >= 1 && fh * fw <= MCCNN_MAX_WINDOW, "median: window %dx%d unsupported (max %d cells)",
                  fh, fw, MCCNN_MAX_WINDOW);
    dim3 block(32, 4), grid(cdiv(W, 32), cdiv(H, 4));
    k_median<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, H, W, (fh - 1) / 2, (fw - 1) / 2);
    MCCNN_LAUNCHED("median");
    return MCCNN_OK;
}

int mccnn_bilateral(const float *img, const float *in, float *out, const float *table, int H, int W, int fh, int fw,
                    float blur_threshold, void *stream) {
    MCCNN_REQUIRE(img && in && out && table && in != out && H >= 1 && W >= 1, "bilateral: bad arguments");
    MCCNN_REQUIRE(fh >= 1 && fw >= 1 && fh * fw <= MCCNN_MAX_WINDOW, "bilateral: window %dx%d unsupported", fh, fw);
    dim3 block(32, 4), grid(cdiv(W, 32), cdiv(H, 4));
    k_bilateral<<<grid, block, 0, (cudaStream_t)stream>>>(img, in, out, table, H, W, fh, fw, blur_threshold);
    MCCNN_LAUNCHED("bilateral");
    return MCCNN_OK;
}

}  // extern "C"
