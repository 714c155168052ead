// Cross-based cost aggregation (pf:117-183, pf:571-657).  Compiled with -fmad=false.
//
// Support regions are kept as four arm lengths per pixel (4 B) instead of the reference's
// explicit coordinate list (6.3 KB per pixel, pf:638); |U| is a precomputed int32.
// The aggregation walks the region in the reference's own enumeration order (pf:640-650) with one
// float32 running sum per cell (pf:157-161), so the result is bit-identical to the reference.
// HWD layout makes that walk cheap: the lanes of a warp are the disparities of one or two
// pixels, so every lane follows the same arms (no divergence) and every load is a contiguous
// 16 B granule of the neighbour pixel's disparity row (fully coalesced for any region shape).
#include "common.cuh"
#include "cbca_tile.cuh"
#include "cbca_stream.cuh"
#include "cbca_march.cuh"
#include "cbca_fused.cuh"
#include <stdlib.h>

namespace mccnn {

// a4: arms.  One thread per pixel; each arm stops at the first neighbour whose intensity differs
// from the ANCHOR by >= tau (pf:588, :596, :615, :623) or after dist-1 pixels.
__global__ void k_cross_arms(const float *__restrict__ img, uchar4 *__restrict__ arms, int H, int W, float tau,
                             int dist) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int h = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= W || h >= H) return;
    const float cur = img[(size_t)h * W + w];
    int up = 0, down = 0, left = 0, right = 0, lim;
    lim = min(dist, h + 1);                                              // pf:585
    for (int b = 1; b < lim; b++) { if (fabsf(cur - img[(size_t)(h - b) * W + w]) >= tau) break; up = b; }
    lim = min(dist, H - h);                                              // pf:593
    for (int b = 1; b < lim; b++) { if (fabsf(cur - img[(size_t)(h + b) * W + w]) >= tau) break; down = b; }
    lim = min(dist, w + 1);                                              // pf:612
    for (int b = 1; b < lim; b++) { if (fabsf(cur - img[(size_t)h * W + w - b]) >= tau) break; left = b; }
    lim = min(dist, W - w);                                              // pf:620
    for (int b = 1; b < lim; b++) { if (fabsf(cur - img[(size_t)h * W + w + b]) >= tau) break; right = b; }
    arms[(size_t)h * W + w] = make_uchar4((unsigned char)up, (unsigned char)down, (unsigned char)left,
                                          (unsigned char)right);
}

// |U(h,w)| = sum over the vertical arm of (left + right + 1) of each spine pixel (pf:640-652).
__global__ void k_cross_count(const uchar4 *__restrict__ arms, int32_t *__restrict__ count, int H, int W) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int h = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= W || h >= H) return;
    uchar4 a = arms[(size_t)h * W + w];
    int n = 0;
    for (int hh = h - a.x; hh <= h + a.y; hh++) {
        uchar4 s = arms[(size_t)hh * W + w];
        n += s.z + s.w + 1;
    }
    count[(size_t)h * W + w] = n;
}

// Compatibility view: the reference's explicit list, in its own order (pf:640-655).
__global__ void k_cross_region_list(const uchar4 *__restrict__ arms, int32_t *__restrict__ region, int H, int W,
                                    int max_num) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int h = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= W || h >= H) return;
    int2 *out = reinterpret_cast<int2 *>(region) + ((size_t)h * W + w) * max_num;
    uchar4 a = arms[(size_t)h * W + w];
    int n = 0;
    for (int k = 0; k <= a.x + a.y; k++) {
        int hh = (k <= a.x) ? h - k : h + (k - a.x);
        uchar4 s = arms[(size_t)hh * W + w];
        for (int j = 0; j <= s.z + s.w; j++) {
            int ww = (j <= s.z) ? w - j : w + (j - s.z);
            out[n++] = make_int2(hh, ww);
        }
    }
    for (; n < max_num; n++) out[n] = make_int2(-1, -1);
}

// a5: one aggregation round.  A block owns a TH x TW pixel tile (so neighbouring regions share
// L1 lines); its threads stride over (pixel, granule) items; one item = 4 disparities of one pixel.
constexpr int CBCA_TH = 8, CBCA_TW = 16, CBCA_THREADS = 256;

__global__ void __launch_bounds__(CBCA_THREADS) k_cbca_round(const float4 *__restrict__ in, float4 *__restrict__ out,
                                                             const uchar4 *__restrict__ arms,
                                                             const int32_t *__restrict__ count, int G, int H, int W) {
    const int h0 = blockIdx.y * CBCA_TH, w0 = blockIdx.x * CBCA_TW;
    const int items = CBCA_TH * CBCA_TW * G;
    for (int item = threadIdx.x; item < items; item += CBCA_THREADS) {
        const int px = item / G, g = item - px * G;
        const int h = h0 + px / CBCA_TW, w = w0 + px % CBCA_TW;
        if (h >= H || w >= W) continue;
        const size_t p = (size_t)h * W + w;
        const uchar4 a = arms[p];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);                    // pf:157
        const int nrows = a.x + a.y;
        for (int k = 0; k <= nrows; k++) {
            const int hh = (k <= a.x) ? h - k : h + (k - a.x);            // spine order h, h-1.., h+1..
            const size_t rowp = (size_t)hh * W;
            const uchar4 s = arms[rowp + w];
            const float4 *src = in + (rowp + w) * G + g;
            // w, w-1, ..., w-left
            for (int j = 0; j <= s.z; j++) {
                float4 v = src[-(ptrdiff_t)j * G];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;  // pf:160
            }
            // w+1, ..., w+right
            for (int j = 1; j <= s.w; j++) {
                float4 v = src[(ptrdiff_t)j * G];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        const float n = (float)count[p];
        out[p * G + g] = make_float4(acc.x / n, acc.y / n, acc.z / n, acc.w / n);   // pf:161
    }
}

}  // namespace mccnn


namespace mccnn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// Tensor maps of one HWD volume for the staged-row boxes of k_cbca_round_tile:
// [level: 4, 2, 1 granules per slab][halo class] -> box {4*stride floats, 8 + 2*halo pixels, 1 row}.
static int build_cbca_maps(CtMaps &maps, const float *vol, int G, int H, int W) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        set_error("cbca: cuTensorMapEncodeTiled is not available from this driver");
        return MCCNN_ERR_CUDA;
    }
    const cuuint64_t Dp = (cuuint64_t)G * 4;
    const cuuint64_t gdim[3] = {Dp, (cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t gstr[2] = {Dp * 4, (cuuint64_t)W * Dp * 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int l = 0; l < CT_NLEVEL; l++) {
        const int gp = 4 >> l;
        for (int i = 0; i < CT_NHALO; i++) {
            const cuuint32_t box[3] = {(cuuint32_t)(4 * ct_stride(gp, G)), (cuuint32_t)(CT_TW + 2 * ct_halo(i)), 1};
            CUresult r = enc(&maps.m[l][i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)vol, gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cbca: cuTensorMapEncodeTiled failed (%d) for box %ux%u", (int)r, box[0], box[1]);
                return MCCNN_ERR_CUDA;
            }
        }
    }
    return MCCNN_OK;
}

// Tensor maps of one HWD volume for k_cbca_march: boxes {32 floats, WT | 2 | 13 pixels, 1 row}.
static int build_cm_maps(CmMaps &maps, const float *vol, int G, int H, int W, int WT) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        set_error("cbca: cuTensorMapEncodeTiled is not available from this driver");
        return MCCNN_ERR_CUDA;
    }
    const cuuint64_t Dp = (cuuint64_t)G * 4;
    const cuuint64_t gdim[3] = {Dp, (cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t gstr[2] = {Dp * 4, (cuuint64_t)W * Dp * 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap *m[3] = {&maps.centre, &maps.small, &maps.full};
    const cuuint32_t px[3] = {(cuuint32_t)WT, (cuuint32_t)CM_HSMALL, (cuuint32_t)CM_ARM};
    for (int i = 0; i < 3; i++) {
        const cuuint32_t box[3] = {(cuuint32_t)(4 * CM_GT), px[i], 1};
        CUresult r = enc(m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)vol, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cbca: cuTensorMapEncodeTiled failed (%d) for box %ux%u", (int)r, box[0], box[1]);
            return MCCNN_ERR_CUDA;
        }
    }
    return MCCNN_OK;
}

// The strip shapes k_cbca_march is built for: {strip width, TMA stages, CTAs per SM}.
struct CmVariant { int wt, nst, per_sm; };
static const CmVariant CM_VARIANTS[] = {{16, 3, 3}, {16, 6, 2}, {24, 3, 2}, {24, 4, 2}, {32, 5, 1}, {32, 3, 1}};
static const int CM_NVARIANTS = (int)(sizeof(CM_VARIANTS) / sizeof(CM_VARIANTS[0]));
static const int CM_DEFAULT_VARIANT = 0;

template <int WT, int NST, int MINB>
static int launch_march(const CmMaps &maps, float *dst, const uint8_t *arms, const int32_t *count, int G, int H, int W,
                        int nW, int nG, int nseg, int hseg, cudaStream_t s) {
    static bool attr_set = false;
    const int smem = (int)sizeof(CmSmem<WT, NST>);
    if (!attr_set) {
        MCCNN_CUDA(cudaFuncSetAttribute(k_cbca_march<WT, NST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    k_cbca_march<WT, NST, MINB><<<nW * nG * nseg, WT * CM_GT + 32, smem, s>>>(
        maps, reinterpret_cast<float4 *>(dst), reinterpret_cast<const uchar4 *>(arms), count, G, H, W, nW, nG, hseg);
    MCCNN_LAUNCHED("cbca_march");
    return MCCNN_OK;
}

// One round of the default mode: row sums src -> scratch, column sums scratch -> out.
static int stream_round(const float *src, float *scratch, float *out, const uint8_t *arms, const int32_t *count, int G, int H,
                        int W, cudaStream_t s) {
    dim3 grid(cdiv(G, CS_GC), cdiv(W, CS_PW), cdiv(H, CS_PH));
    k_cbca_pass<false, CS_ITEMS, 1><<<grid, CS_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(scratch),
                                                                reinterpret_cast<const uchar4 *>(arms), count, G, H, W);
    MCCNN_LAUNCHED("cbca_rows");
    k_cbca_pass<true, CS_ITEMS, 1><<<grid, CS_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(scratch), reinterpret_cast<float4 *>(out),
                                                               reinterpret_cast<const uchar4 *>(arms), count, G, H, W);
    MCCNN_LAUNCHED("cbca_cols");
    return MCCNN_OK;
}

// The last round of a call whose result is re-partitioned: the column pass stores into the row slabs of the owners.
static int stream_round_to(const float *src, float *scratch, const CsScatter &sc, const uint8_t *arms, const int32_t *count,
                           int G, int H, int W, cudaStream_t s) {
    dim3 grid(cdiv(G, CS_GC), cdiv(W, CS_PW), cdiv(H, CS_PH));
    k_cbca_pass<false, CS_ITEMS, 1><<<grid, CS_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(scratch),
                                                                reinterpret_cast<const uchar4 *>(arms), count, G, H, W);
    MCCNN_LAUNCHED("cbca_rows");
    k_cbca_pass<true, CS_ITEMS, 1, true><<<grid, CS_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(scratch), nullptr,
                                                                     reinterpret_cast<const uchar4 *>(arms), count, G, H, W, sc);
    MCCNN_LAUNCHED("cbca_cols_scatter");
    return MCCNN_OK;
}

}  // namespace mccnn

using namespace mccnn;

extern "C" {

int mccnn_cross_arms(const float *img, uint8_t *arms, int32_t *count, int H, int W, float tau, int dist,
                     void *stream) {
    MCCNN_REQUIRE(img && arms && count && H >= 1 && W >= 1, "cross_arms: bad arguments");
    MCCNN_REQUIRE(dist >= 1 && dist <= 255, "cross_arms: distance_threshold %d outside [1, 255]", dist);
    MCCNN_REQUIRE(tau > 0.0f, "cross_arms: intensity_threshold must be > 0 (asserts at pf:601, :628)");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(32, 8), grid(cdiv(W, 32), cdiv(H, 8));
    k_cross_arms<<<grid, block, 0, s>>>(img, reinterpret_cast<uchar4 *>(arms), H, W, tau, dist);
    MCCNN_LAUNCHED("cross_arms");
    k_cross_count<<<grid, block, 0, s>>>(reinterpret_cast<const uchar4 *>(arms), count, H, W);
    MCCNN_LAUNCHED("cross_count");
    return MCCNN_OK;
}

int mccnn_cross_region_list(const uint8_t *arms, int32_t *region, int H, int W, int dist, void *stream) {
    MCCNN_REQUIRE(arms && region && H >= 1 && W >= 1 && dist >= 1 && dist <= 255, "cross_region_list: bad arguments");
    dim3 block(32, 8), grid(cdiv(W, 32), cdiv(H, 8));
    k_cross_region_list<<<grid, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uchar4 *>(arms), region, H, W,
                                                                  (2 * dist) * (2 * dist));
    MCCNN_LAUNCHED("cross_region_list");
    return MCCNN_OK;
}

static const int CBCA_MAX_ROUNDS = 1024;       // one tile counter per round of a call

size_t mccnn_cbca_workspace_bytes(int H, int W) {
    if (H < 1 || W < 1) return 0;
    // tiled mode: per-tile schedule + one tile counter per round; L2 mode: ticket + two completion counters per band per round
    const size_t tiled = (size_t)cdiv(W, CT_TW) * cdiv(H, CT_TH) * sizeof(CbcaTileMeta) + CBCA_MAX_ROUNDS * sizeof(unsigned);
    const size_t fused = (size_t)CBCA_MAX_ROUNDS * (1 + 2 * (size_t)cdiv(H, CS_PH)) * sizeof(unsigned);
    return tiled > fused ? tiled : fused;
}

int mccnn_cbca_to(const float *in, float *out, float *scratch, const uint8_t *arms, const int32_t *count, int D, int H,
                  int W, int iters, int nparts, const int *row_bounds, float *const *dst, int g_offset, int g_total,
                  void *stream) {
    MCCNN_REQUIRE(in && out && scratch && arms && count && D >= 1 && H >= 1 && W >= 1 && iters >= 1, "cbca_to: bad arguments");
    MCCNN_REQUIRE(in != out && scratch != in && scratch != out, "cbca_to: in, out and scratch must differ");
    MCCNN_REQUIRE(nparts >= 1 && nparts <= CS_MAX_PARTS && row_bounds && dst, "cbca_to: 1 to %d parts", CS_MAX_PARTS);
    MCCNN_REQUIRE(row_bounds[0] == 0 && row_bounds[nparts] == H, "cbca_to: row bounds must tile [0, %d)", H);
    const int G = dpitch(D) / 4;
    MCCNN_REQUIRE(g_offset >= 0 && g_offset + G <= g_total, "cbca_to: granules [%d, %d) outside the destination pitch %d",
                  g_offset, g_offset + G, g_total);
    CsScatter sc;
    sc.nparts = nparts; sc.g_off = g_offset; sc.g_total = g_total;
    for (int r = 0; r <= nparts; r++) sc.lo[r] = row_bounds[r];
    for (int r = 0; r < nparts; r++) {
        MCCNN_REQUIRE(row_bounds[r] < row_bounds[r + 1] && dst[r], "cbca_to: empty part or null destination %d", r);
        sc.base[r] = reinterpret_cast<float4 *>(dst[r]);
    }
    cudaStream_t s = (cudaStream_t)stream;
    const float *src = in;
    for (int it = 0; it + 1 < iters; it++) {
        int rc = stream_round(src, scratch, out, arms, count, G, H, W, s);
        if (rc) return rc;
        src = out;
    }
    return stream_round_to(src, scratch, sc, arms, count, G, H, W, s);
}

int mccnn_cbca(const float *in, float *out, float *scratch, const uint8_t *arms, const int32_t *count, int D, int H,
               int W, int iters, int dist, int mode, void *workspace, void *stream) {
    MCCNN_REQUIRE(mode == MCCNN_CBCA_SEPARABLE || mode == MCCNN_CBCA_EXACT || mode == MCCNN_CBCA_SEPARABLE_TILED ||
                      mode == MCCNN_CBCA_SEPARABLE_MARCH || mode == MCCNN_CBCA_SEPARABLE_L2,
                  "cbca: unknown mode %d", mode);
    MCCNN_REQUIRE(dist >= 1 && dist <= 255, "cbca: distance_threshold %d outside [1, 255]", dist);
    MCCNN_REQUIRE(mode != MCCNN_CBCA_SEPARABLE_TILED || dist <= CT_MAXARM + 1,
                  "cbca: separable mode supports distance_threshold <= 14 (got %d); use MCCNN_CBCA_EXACT", dist);
    MCCNN_REQUIRE(in && out && arms && count && D >= 1 && H >= 1 && W >= 1 && iters >= 0, "cbca: bad arguments");
    MCCNN_REQUIRE(H <= 65535 && W <= 65535, "cbca: image too large");
    MCCNN_REQUIRE(in != out, "cbca: in and out must differ (the reference leaves its input untouched, pf:119)");
    MCCNN_REQUIRE(mode != MCCNN_CBCA_SEPARABLE_TILED || iters <= CBCA_MAX_ROUNDS, "cbca: at most %d rounds per call", CBCA_MAX_ROUNDS);
    MCCNN_REQUIRE(mode != MCCNN_CBCA_SEPARABLE_TILED || iters == 0 || workspace,
                  "cbca: separable mode needs a workspace of mccnn_cbca_workspace_bytes(H, W) bytes");
    cudaStream_t s = (cudaStream_t)stream;
    const int Dp = dpitch(D), G = Dp / 4;
    // the marching kernel's boxes are 8 granules x up to 13 halo pixels: other shapes take the two streaming passes
    if (mode == MCCNN_CBCA_SEPARABLE_MARCH && (dist > CM_ARM + 1 || G < CM_GT)) mode = MCCNN_CBCA_SEPARABLE;
    const bool two_pass = mode == MCCNN_CBCA_SEPARABLE || mode == MCCNN_CBCA_SEPARABLE_L2;
    MCCNN_REQUIRE(iters < (two_pass ? 1 : 2) || (scratch && scratch != in && scratch != out),
                  "cbca: scratch volume required (separable: any round; other modes: iters >= 2)");
    MCCNN_REQUIRE(mode != MCCNN_CBCA_SEPARABLE_L2 || iters == 0 || (workspace && iters <= CBCA_MAX_ROUNDS),
                  "cbca: L2 mode needs a workspace of mccnn_cbca_workspace_bytes(H, W) bytes and at most %d rounds", CBCA_MAX_ROUNDS);
    if (iters == 0) {
        MCCNN_CUDA(cudaMemcpyAsync(out, in, (size_t)H * W * Dp * sizeof(float), cudaMemcpyDeviceToDevice, s));
        return MCCNN_OK;
    }
    if (mode == MCCNN_CBCA_SEPARABLE) {
        // every round: row sums src -> scratch, column sums scratch -> out; the next round reads out
        const float *src = in;
        for (int it = 0; it < iters; it++) {
            int rc = stream_round(src, scratch, out, arms, count, G, H, W, s);
            if (rc) return rc;
            src = out;
        }
        return MCCNN_OK;
    }
    if (mode == MCCNN_CBCA_SEPARABLE_L2) {
        // one persistent kernel per round; row sums in the first min(H, 64) rows of scratch, which stay in L2
        static int fused_grid = 0;
        if (fused_grid == 0) {
            int dev = 0, sms = 0, per_sm = 0;
            MCCNN_CUDA(cudaGetDevice(&dev));
            MCCNN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            MCCNN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cbca_fused, CS_THREADS, 0));
            fused_grid = sms * (per_sm > 0 ? per_sm : 1);
        }
        CfSched sc;
        sc.nbx = cdiv(W, CS_PW); sc.nz = cdiv(G, CS_GC); sc.nb = cdiv(H, CS_PH);
        sc.lag = sc.nb < CF_LAG ? sc.nb : CF_LAG;
        sc.items_band = sc.nbx * sc.nz;
        sc.chunk = 1;
        if (const char *e = getenv("MCCNN_CBCA_L2_CHUNK")) { const int c = atoi(e); if (c >= 1 && c <= 64) sc.chunk = c; }
        sc.chunks_band = cdiv(sc.items_band, sc.chunk);
        const long long total = 2ll * sc.nb * sc.chunks_band;
        MCCNN_REQUIRE(total < (1ll << 31), "cbca: image too large");
        sc.total = (int)total;
        const size_t per_round = 1 + 2 * (size_t)sc.nb;
        unsigned *counters = reinterpret_cast<unsigned *>(workspace);
        MCCNN_CUDA(cudaMemsetAsync(counters, 0, (size_t)iters * per_round * sizeof(unsigned), s));
        const int grid = total < fused_grid ? (int)total : fused_grid;
        const float *src = in;
        for (int it = 0; it < iters; it++) {
            k_cbca_fused<<<grid, CS_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(scratch),
                                                     reinterpret_cast<float4 *>(out), reinterpret_cast<const uchar4 *>(arms), count,
                                                     G, H, W, sc, counters + (size_t)it * per_round);
            MCCNN_LAUNCHED("cbca_fused");
            src = out;
        }
        return MCCNN_OK;
    }
    // ping-pong so that the last round lands in `out`
    float *buf[2];
    buf[(iters - 1) & 1] = out;
    buf[iters & 1] = scratch;
    const float *src = in;
    if (mode == MCCNN_CBCA_SEPARABLE_MARCH) {
        static int num_sms_m = 0;
        if (num_sms_m == 0) {
            int dev = 0;
            MCCNN_CUDA(cudaGetDevice(&dev));
            MCCNN_CUDA(cudaDeviceGetAttribute(&num_sms_m, cudaDevAttrMultiProcessorCount, dev));
        }
        // MCCNN_CBCA_MARCH="variant[,row segments]" overrides the strip shape / segmentation (tuning and tests)
        int variant = CM_DEFAULT_VARIANT, nseg = 0;
        if (const char *e = getenv("MCCNN_CBCA_MARCH")) {
            int v = -1, n = 0;
            const int got = sscanf(e, "%d,%d", &v, &n);
            if (got >= 1 && v >= 0 && v < CM_NVARIANTS) variant = v;
            if (got >= 2 && n >= 1) nseg = n;
        }
        const CmVariant cv = CM_VARIANTS[variant];
        const int nW = cdiv(W, cv.wt), nG = cdiv(G, CM_GT);
        if (nseg == 0) {
            // fewest (waves x rows marched per CTA): a segment re-forms the 13 row sums above and below it
            const long long slots = (long long)num_sms_m * cv.per_sm;
            long long best = -1;
            for (int n = 1; n <= 16 && n <= H; n++) {
                const long long units = (long long)nW * nG * n;
                const long long cost = ((units + slots - 1) / slots) * (cdiv(H, n) + 2 * CM_ARM);
                if (best < 0 || cost < best) { best = cost; nseg = n; }
            }
        }
        if (nseg > H) nseg = H;
        const int hseg = cdiv(H, nseg);
        nseg = cdiv(H, hseg);
        const float *vols[3] = {in, iters >= 2 ? buf[0] : nullptr, iters >= 3 ? buf[1] : nullptr};
        CmMaps maps[3];
        for (int v = 0; v < 3; v++)
            if (vols[v]) {
                int rc = build_cm_maps(maps[v], vols[v], G, H, W, cv.wt);
                if (rc) return rc;
            }
        for (int it = 0; it < iters; it++) {
            float *dst = buf[it & 1];
            const CmMaps &m = maps[(src == in) ? 0 : (src == buf[0] ? 1 : 2)];
            int rc;
            switch (variant) {
                case 0: rc = launch_march<16, 3, 3>(m, dst, arms, count, G, H, W, nW, nG, nseg, hseg, s); break;
                case 1: rc = launch_march<16, 6, 2>(m, dst, arms, count, G, H, W, nW, nG, nseg, hseg, s); break;
                case 2: rc = launch_march<24, 3, 2>(m, dst, arms, count, G, H, W, nW, nG, nseg, hseg, s); break;
                case 3: rc = launch_march<24, 4, 2>(m, dst, arms, count, G, H, W, nW, nG, nseg, hseg, s); break;
                case 4: rc = launch_march<32, 5, 1>(m, dst, arms, count, G, H, W, nW, nG, nseg, hseg, s); break;
                default: rc = launch_march<32, 3, 1>(m, dst, arms, count, G, H, W, nW, nG, nseg, hseg, s); break;
            }
            if (rc) return rc;
            src = dst;
        }
        return MCCNN_OK;
    }
    if (mode == MCCNN_CBCA_EXACT) {
        dim3 grid(cdiv(W, CBCA_TW), cdiv(H, CBCA_TH));
        for (int it = 0; it < iters; it++) {
            float *dst = buf[it & 1];
            k_cbca_round<<<grid, CBCA_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(src),
                                                        reinterpret_cast<float4 *>(dst),
                                                        reinterpret_cast<const uchar4 *>(arms), count, G, H, W);
            MCCNN_LAUNCHED("cbca_round");
            src = dst;
        }
        return MCCNN_OK;
    }

    int gp_top = 1;
    while (gp_top < G && gp_top < CT_GPMAX) gp_top <<= 1;
    const int tilesX = cdiv(W, CT_TW), tilesY = cdiv(H, CT_TH), ntiles = tilesX * tilesY;
    static int num_sms = 0;
    static bool smem_set = false;
    if (num_sms == 0) {
        int dev = 0;
        MCCNN_CUDA(cudaGetDevice(&dev));
        MCCNN_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (!smem_set) {
        MCCNN_CUDA(cudaFuncSetAttribute(k_cbca_round_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_BYTES));
        smem_set = true;
    }
    const int tgrid = ntiles < num_sms * CT_CTAS_PER_SM ? ntiles : num_sms * CT_CTAS_PER_SM;
    CbcaTileMeta *meta = reinterpret_cast<CbcaTileMeta *>(workspace);
    unsigned *counters = reinterpret_cast<unsigned *>(meta + ntiles);
    MCCNN_CUDA(cudaMemsetAsync(counters, 0, (size_t)iters * sizeof(unsigned), s));
    k_cbca_tile_meta<<<dim3(tilesX, tilesY), 64, 0, s>>>(reinterpret_cast<const uchar4 *>(arms), count, meta, G, H, W, gp_top);
    MCCNN_LAUNCHED("cbca_tile_meta");
    // tensor maps of the (at most three) volumes this call reads
    const float *vols[3] = {in, iters >= 2 ? buf[0] : nullptr, iters >= 3 ? buf[1] : nullptr};
    CtMaps maps[3];
    for (int v = 0; v < 3; v++)
        if (vols[v]) {
            int rc = build_cbca_maps(maps[v], vols[v], G, H, W);
            if (rc) return rc;
        }
    for (int it = 0; it < iters; it++) {
        float *dst = buf[it & 1];
        const int v = (src == in) ? 0 : (src == buf[0] ? 1 : 2);
        k_cbca_round_tile<<<tgrid, CT_THREADS, CT_SMEM_BYTES, s>>>(maps[v], reinterpret_cast<float4 *>(dst), meta, G, H, W,
                                                                    ntiles, counters + it);
        MCCNN_LAUNCHED("cbca_round_tile");
        src = dst;
    }
    return MCCNN_OK;
}

}  // extern "C"
