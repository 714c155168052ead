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
#include "cbca_stream.cuh"
#include "cbca_chain.cuh"

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

// sum over the image of up + down (what decides between the two bit-identical separable schedules, process_functional.py)
__global__ void k_arms_vertical_sum(const uchar4 *__restrict__ arms, unsigned long long *__restrict__ sum, long long P) {
    unsigned v = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const uchar4 a = arms[p];
        v += (unsigned)a.x + (unsigned)a.y;
    }
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(sum, (unsigned long long)v);
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


// The closing column pass of a separable call: plain, scattered to the row slabs of other ranks (sc), or with the
// winner-take-all folded in (wt).
static int closing_cols(const float *hs, float *out, const CsScatter *sc, const CsWta *wt, const uint8_t *arms, const int32_t *count,
                        int G, int H, int W, cudaStream_t s) {
    dim3 grid(cdiv(G, CS_GC), cdiv(W, CS_PW), cdiv(H, CS_PH));
    const float4 *src = reinterpret_cast<const float4 *>(hs);
    const uchar4 *a4 = reinterpret_cast<const uchar4 *>(arms);
    if (wt) {
        MCCNN_CUDA(cudaMemsetAsync(wt->keys, 0xff, (size_t)H * W * sizeof(unsigned long long), s));
        k_cbca_pass<true, CS_ITEMS, 1, false, true><<<grid, CS_THREADS, 0, s>>>(src, reinterpret_cast<float4 *>(out), a4, count, G, H, W,
                                                                               CsScatter(), *wt);
        MCCNN_LAUNCHED("cbca_cols_wta");
    } else if (sc) {
        k_cbca_pass<true, CS_ITEMS, 1, true><<<grid, CS_THREADS, 0, s>>>(src, nullptr, a4, count, G, H, W, *sc);
        MCCNN_LAUNCHED("cbca_cols_scatter");
    } else {
        k_cbca_pass<true, CS_ITEMS, 1><<<grid, CS_THREADS, 0, s>>>(src, reinterpret_cast<float4 *>(out), a4, count, G, H, W);
        MCCNN_LAUNCHED("cbca_cols");
    }
    return MCCNN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// One round of the separable mode: row sums src -> scratch, column sums scratch -> out (or scattered, sc != NULL).
static int stream_round(const float *src, float *scratch, float *out, const CsScatter *sc, const CsWta *wt, const uint8_t *arms,
                        const int32_t *count, int G, int H, int W, cudaStream_t s) {
    dim3 grid(cdiv(G, CS_GC), cdiv(W, CS_PW), cdiv(H, CS_PH));
    k_cbca_pass<false, CS_ITEMS, 1><<<grid, CS_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(scratch),
                                                                reinterpret_cast<const uchar4 *>(arms), count, G, H, W);
    MCCNN_LAUNCHED("cbca_rows");
    return closing_cols(scratch, out, sc, wt, arms, count, G, H, W, s);
}

// n >= 1 rounds as two streaming passes each; every round reads the previous round's `out`; the last column pass may scatter
static int two_pass_rounds(const float *in, float *out, float *scratch, const CsScatter *sc, const CsWta *wt, const uint8_t *arms,
                           const int32_t *count, int G, int H, int W, int iters, cudaStream_t s) {
    const float *src = in;
    for (int it = 0; it < iters; it++) {
        const bool last = it + 1 == iters;
        int rc = stream_round(src, scratch, out, last ? sc : nullptr, last ? wt : nullptr, arms, count, G, H, W, s);
        if (rc) return rc;
        src = out;
    }
    return MCCNN_OK;
}

// The chained round's shapes (cbca_chain.cuh): pixels per segment, threads, resident CTAs asked for, granules per thread.
// Measured at 1024 x 1024, natural image, ms per round of a 16-round call (B200):
// (all before pixels without arms were dropped from the passes that find them settled: G3 is 0.269 on those, 0.333 on the first)
//   ndisp 192 (48 granules): G1 0.362, G3 0.333 (other GPT = 3 shapes: 14 px x 128 thr 0.334, 30 x 256 0.338, 22 x 128 0.356,
//                            30 x 128 0.383; without the L2 look-ahead G3 is 0.364); round 2's first cp.async kernel 0.438,
//                            two streaming passes 0.580
//   ndisp 256 (64 granules): G2 0.449 (before the look-ahead: G1 0.508, G2 0.481, four granules per thread 0.522)
//   ndisp 400 (100 granules): G1 0.826 (before the look-ahead: G1 0.894, G2 0.985, G3 1.170; granule groups are padded to
//                            16 x GPT: dead lanes)
typedef CgShape<30, 128, 8, 1> CgG1;         // 25 KB, 8 CTAs per SM
typedef CgShape<14, 128, 8, 2> CgG2;         // 25 KB
typedef CgShape<22, 192, 4, 3> CgG3;         // 55 KB, 4 CTAs of 6 warps per SM

template <class C>
static int launch_colrow_g(const float *hs_in, float *hs_out, const uint8_t *arms, const int32_t *count, int G, int H, int W,
                           int settled, cudaStream_t s) {
    CUtensorMap map, map_row;
    int rc = tc_encode_map_3d(map, hs_in, (unsigned long long)G * 4, W, H, CS_GC * 4 * C::GPT, C::NP, false, "cbca", 3);
    if (rc) return rc;
    rc = tc_encode_map_3d(map_row, hs_in, (unsigned long long)G * 4, W, H, CS_GC * 4 * C::GPT, C::NP, false, "cbca", 1);
    if (rc) return rc;
    const int ahead = 8;      // rows of L2 look-ahead; measured at C3: 0 -> 0.362 ms per round, 4 .. 24 -> 0.339, 32 -> 0.350, 64 -> 0.384
    // per device and cheap: set on every launch rather than cached in a process-wide static
    MCCNN_CUDA(cudaFuncSetAttribute(k_cbca_colrow_g<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    dim3 grid(cdiv(G, CS_GC * C::GPT), cdiv(W, C::S), H);
    k_cbca_colrow_g<C><<<grid, C::NT, C::SMEM, s>>>(map, map_row, reinterpret_cast<const float4 *>(hs_in), reinterpret_cast<float4 *>(hs_out),
                                                    reinterpret_cast<const uchar4 *>(arms), count, G, H, W, ahead, settled);
    MCCNN_LAUNCHED("cbca_colrow");
    return MCCNN_OK;
}

// granules per thread for a pitch of G granules: the cheapest padded cover, cost per 16 granules 1 / 0.947 / 0.943 for
// GPT = 1 / 2 / 3 from the table above; 0 if no shape holds the far halo of arms up to hm pixels (two passes then)
static int colrow_gpt(int G, int hm) {
    const double c[4] = {0.0, 1.0, 0.947, 0.943};
    const bool ok[4] = {false, CgG1::supports(hm), CgG2::supports(hm), CgG3::supports(hm)};
    int best = 0;
    double bc = 1e30;
    for (int gpt = 1; gpt <= 3; gpt++) {
        const double cost = cdiv(G, CS_GC * gpt) * gpt * c[gpt];
        if (ok[gpt] && cost < bc) { bc = cost; best = gpt; }
    }
    return best;
}

template <class C>
static int launch_close_g(const float *hs_in, float *out, const CsWta *wt, const uint8_t *arms, const int32_t *count, int G, int H,
                          int W, cudaStream_t s) {
    CUtensorMap map, map_row;
    int rc = tc_encode_map_3d(map, hs_in, (unsigned long long)G * 4, W, H, CS_GC * 4 * C::GPT, C::NP, false, "cbca", 3);
    if (rc) return rc;
    rc = tc_encode_map_3d(map_row, hs_in, (unsigned long long)G * 4, W, H, CS_GC * 4 * C::GPT, C::NP, false, "cbca", 1);
    if (rc) return rc;
    MCCNN_CUDA(cudaFuncSetAttribute(k_cbca_close_g<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    CsWta w = CsWta();
    if (wt) {
        w = *wt;
        MCCNN_CUDA(cudaMemsetAsync(w.keys, 0xff, (size_t)H * W * sizeof(unsigned long long), s));
    }
    dim3 grid(cdiv(G, CS_GC * C::GPT), cdiv(W, C::S), H);
    k_cbca_close_g<C><<<grid, C::NT, C::SMEM, s>>>(map, map_row, reinterpret_cast<const float4 *>(hs_in), reinterpret_cast<float4 *>(out),
                                                   reinterpret_cast<const uchar4 *>(arms), count, G, H, W, 8, 1, wt ? wt->store : 1, w);
    MCCNN_LAUNCHED("cbca_close");
    return MCCNN_OK;
}

static int launch_close(int gpt, const float *hs_in, float *out, const CsWta *wt, const uint8_t *arms, const int32_t *count, int G, int H,
                        int W, cudaStream_t s) {
    if (gpt == 3) return launch_close_g<CgG3>(hs_in, out, wt, arms, count, G, H, W, s);
    if (gpt == 2) return launch_close_g<CgG2>(hs_in, out, wt, arms, count, G, H, W, s);
    return launch_close_g<CgG1>(hs_in, out, wt, arms, count, G, H, W, s);
}

static int launch_colrow(int gpt, const float *hs_in, float *hs_out, const uint8_t *arms, const int32_t *count, int G, int H, int W,
                         int settled, cudaStream_t s) {
    if (gpt == 3) return launch_colrow_g<CgG3>(hs_in, hs_out, arms, count, G, H, W, settled, s);
    if (gpt == 2) return launch_colrow_g<CgG2>(hs_in, hs_out, arms, count, G, H, W, settled, s);
    return launch_colrow_g<CgG1>(hs_in, hs_out, arms, count, G, H, W, settled, s);
}

// n >= 2 rounds, chained: rows | (n-1) x colrow | cols.  The row sums ping-pong between `out` and `scratch` so that
// the last ones sit in `scratch` and the closing column pass can write `out` (or scatter, sc != NULL).
static int chained_rounds(const float *in, float *out, float *scratch, const CsScatter *sc, const CsWta *wt, const uint8_t *arms,
                          const int32_t *count, int G, int H, int W, int iters, int gpt, cudaStream_t s) {
    dim3 grid(cdiv(G, CS_GC), cdiv(W, CS_PW), cdiv(H, CS_PH));
    float *hs[2];
    hs[(iters - 1) & 1] = scratch;
    hs[iters & 1] = out;
    k_cbca_pass<false, CS_ITEMS, 1><<<grid, CS_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(in), reinterpret_cast<float4 *>(hs[0]),
                                                                reinterpret_cast<const uchar4 *>(arms), count, G, H, W);
    MCCNN_LAUNCHED("cbca_rows");
    for (int k = 1; k < iters; k++) {
        // hs[k & 1] was last written two passes ago (k = 2: by the row pass): from k = 2 on, pixels without arms are settled
        int rc = launch_colrow(gpt, hs[(k - 1) & 1], hs[k & 1], arms, count, G, H, W, k >= 2, s);
        if (rc) return rc;
    }
    if (sc) return closing_cols(scratch, out, sc, wt, arms, count, G, H, W, s);
    // `out` was last written two passes ago (iters = 2: by the row pass): it already holds the pixels without arms
    return launch_close(gpt, scratch, out, wt, arms, count, G, H, W, s);
}

// the default: chained rounds wherever they apply (two rounds or more; the shared-memory tile grows with the arm limit)
static int separable_rounds(const float *in, float *out, float *scratch, const CsScatter *sc, const CsWta *wt, const uint8_t *arms,
                            const int32_t *count, int G, int H, int W, int iters, int hm, cudaStream_t s) {
    const int gpt = colrow_gpt(G, hm);
    if (iters >= 2 && gpt) return chained_rounds(in, out, scratch, sc, wt, arms, count, G, H, W, iters, gpt, s);
    return two_pass_rounds(in, out, scratch, sc, wt, arms, count, G, H, W, iters, s);
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

int mccnn_arms_vertical_sum(const uint8_t *arms, int H, int W, unsigned long long *sum, void *stream) {
    MCCNN_REQUIRE(arms && sum && H >= 1 && W >= 1, "arms_vertical_sum: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    MCCNN_CUDA(cudaMemsetAsync(sum, 0, sizeof(unsigned long long), s));
    const long long P = (long long)H * W;
    k_arms_vertical_sum<<<(int)(P < 148 * 1024 ? (P + 255) / 256 : 592), 256, 0, s>>>(reinterpret_cast<const uchar4 *>(arms), sum, P);
    MCCNN_LAUNCHED("arms_vertical_sum");
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

int mccnn_cbca_to(const float *in, float *out, float *scratch, const uint8_t *arms, const int32_t *count, int D, int H,
                  int W, int iters, int dist, int nparts, const int *row_bounds, float *const *dst, int g_offset, int g_total,
                  void *stream) {
    MCCNN_REQUIRE(in && out && scratch && arms && count && D >= 1 && H >= 1 && W >= 1 && iters >= 1, "cbca_to: bad arguments");
    MCCNN_REQUIRE(H <= 65535 && W <= 65535, "cbca_to: image too large");
    MCCNN_REQUIRE(dist >= 1 && dist <= 255, "cbca_to: distance_threshold %d outside [1, 255]", dist);
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
    return separable_rounds(in, out, scratch, &sc, nullptr, arms, count, G, H, W, iters, dist - 1, (cudaStream_t)stream);
}

int mccnn_cbca_wta(const float *in, float *out, float *scratch, const uint8_t *arms, const int32_t *count, int D, int H, int W,
                   int iters, int dist, int mode, int store_volume, void *keys, float *disp, void *stream) {
    MCCNN_REQUIRE(mode == MCCNN_CBCA_SEPARABLE || mode == MCCNN_CBCA_SEPARABLE_TWO_PASS,
                  "cbca_wta: mode %d has no fused winner-take-all (separable modes only)", mode);
    MCCNN_REQUIRE(dist >= 1 && dist <= 255, "cbca_wta: distance_threshold %d outside [1, 255]", dist);
    MCCNN_REQUIRE(in && out && scratch && arms && count && keys && disp && D >= 1 && H >= 1 && W >= 1 && iters >= 1,
                  "cbca_wta: bad arguments");
    MCCNN_REQUIRE(H <= 65535 && W <= 65535, "cbca_wta: image too large");
    MCCNN_REQUIRE(in != out && scratch != in && scratch != out, "cbca_wta: in, out and scratch must differ");
    MCCNN_REQUIRE(((uintptr_t)keys & 7) == 0, "cbca_wta: keys must be 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const int G = dpitch(D) / 4;
    CsWta wt;
    wt.keys = reinterpret_cast<unsigned long long *>(keys); wt.D = D; wt.store = store_volume != 0;
    int rc = mode == MCCNN_CBCA_SEPARABLE ? separable_rounds(in, out, scratch, nullptr, &wt, arms, count, G, H, W, iters, dist - 1, s)
                                          : two_pass_rounds(in, out, scratch, nullptr, &wt, arms, count, G, H, W, iters, s);
    if (rc) return rc;
    const long long P = (long long)H * W;
    k_wta_decode<<<cdiv(P, 256), 256, 0, s>>>(wt.keys, disp, P);
    MCCNN_LAUNCHED("wta_decode");
    return MCCNN_OK;
}

int mccnn_cbca(const float *in, float *out, float *scratch, const uint8_t *arms, const int32_t *count, int D, int H,
               int W, int iters, int dist, int mode, void *stream) {
    MCCNN_REQUIRE(mode == MCCNN_CBCA_SEPARABLE || mode == MCCNN_CBCA_EXACT || mode == MCCNN_CBCA_SEPARABLE_TWO_PASS,
                  "cbca: unknown mode %d", mode);
    MCCNN_REQUIRE(dist >= 1 && dist <= 255, "cbca: distance_threshold %d outside [1, 255]", dist);
    MCCNN_REQUIRE(in && out && arms && count && D >= 1 && H >= 1 && W >= 1 && iters >= 0, "cbca: bad arguments");
    MCCNN_REQUIRE(H <= 65535 && W <= 65535, "cbca: image too large");
    MCCNN_REQUIRE(in != out, "cbca: in and out must differ (the reference leaves its input untouched, pf:119)");
    cudaStream_t s = (cudaStream_t)stream;
    const int Dp = dpitch(D), G = Dp / 4;
    const bool separable = mode != MCCNN_CBCA_EXACT;
    MCCNN_REQUIRE(iters < (separable ? 1 : 2) || (scratch && scratch != in && scratch != out),
                  "cbca: scratch volume required (separable: any round; exact: iters >= 2)");
    if (iters == 0) {
        MCCNN_CUDA(cudaMemcpyAsync(out, in, (size_t)H * W * Dp * sizeof(float), cudaMemcpyDeviceToDevice, s));
        return MCCNN_OK;
    }
    if (mode == MCCNN_CBCA_SEPARABLE) return separable_rounds(in, out, scratch, nullptr, nullptr, arms, count, G, H, W, iters, dist - 1, s);
    if (mode == MCCNN_CBCA_SEPARABLE_TWO_PASS) return two_pass_rounds(in, out, scratch, nullptr, nullptr, arms, count, G, H, W, iters, s);
    // exact: ping-pong so that the last round lands in `out`
    float *buf[2];
    buf[(iters - 1) & 1] = out;
    buf[iters & 1] = scratch;
    const float *src = in;
    dim3 grid(cdiv(W, CBCA_TW), cdiv(H, CBCA_TH));
    for (int it = 0; it < iters; it++) {
        float *dst = buf[it & 1];
        k_cbca_round<<<grid, CBCA_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst),
                                                    reinterpret_cast<const uchar4 *>(arms), count, G, H, W);
        MCCNN_LAUNCHED("cbca_round");
        src = dst;
    }
    return MCCNN_OK;
}

}  // extern "C"
