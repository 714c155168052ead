// Separable cross-based aggregation as two streaming passes (default fast mode of mccnn_cbca).
//
// One round is  out(h,w) = ( sum_{h' in spine(h,w)} Hs(h',w) ) / |U(h,w)|,  Hs(h',w) = sum_{w' in arm(h',w)} in(h',w')
// (pf:640-650): k_cbca_pass<rows> writes Hs, k_cbca_pass<cols> adds it along the spine and divides.  Both are
// pure gathers with no barriers: the lanes of a warp are consecutive disparity granules of one pixel (two
// pixels per warp), so every load and store is a contiguous run and an arm walk is warp uniform; neighbouring
// pixels are served by L2 (a CTA covers an 8x4 pixel patch per 64 disparities).  HBM traffic is 16 B per cell
// per round -- twice the fused minimum -- but the passes run closer to copy speed than any of the seven fused
// kernels tried in rounds 1 and 2 (profiles/r2_cbca_chained_experiment.md): the round is bound by
// bytes in flight and instruction issue, and every way of keeping a round's result on chip pays in one of the two.
// Summation order inside a row and along the spine is the reference's; only the association
// (row sums first) differs: ~1e-7 relative.
#pragma once
#include <math_constants.h>
#include "common.cuh"

namespace mccnn {

constexpr int CS_PH = 8, CS_PW = 4, CS_GC = 16, CS_THREADS = 128;   // patch 8x4 pixels x 16 granules per CTA
constexpr int CS_MIN_BLOCKS = 8;   // <= 64 registers: 8 CTAs per SM.  Measured at C3 (ms per round): 6 -> 0.673, 7 -> 0.628,
                                   // 8 -> 0.608, 9 -> 0.647, 10 -> 0.740 (rows / columns tuned separately: 8 / 8 is still best); other shapes (items per thread, staged
                                   // neighbours): (4,1) 0.608, (3,1) 0.621, (2,1) 0.704, (4,2) 0.825, (2,2) 0.735; cp.async.ca 0.737

__device__ __forceinline__ void cs_add(float4 &acc, const float4 v) {
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
}

constexpr int CS_ITEMS = (CS_PH * CS_PW * CS_GC) / CS_THREADS;    // (pixel, granule) items per thread: 4

__device__ __forceinline__ void cs_cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}

// a / n for n = |U| (an integer <= 729), y = RN(1/n): q = RN(a*y); r = a - n*q (exact, FMA); RN(q + r*y):
// correctly rounded (Markstein), i.e. the IEEE division of pf:161; odd magnitudes take the plain division.
__device__ __forceinline__ float cs_div1(float a, float n, float y) {
    const float q = a * y;
    const float r = fmaf(-n, q, a);
    return fmaf(r, y, q);
}

// One pass, rows (COLS = false: Hs = sum along the row arm, stride one pixel) or columns (COLS = true: sum of Hs
// along the spine, stride one image row, then / |U|).  A CTA is 128 threads = 16 granules x 8 pixel slots and covers
// a (2 * ITEMS) x 4 pixel patch, ITEMS (pixel, granule) items per thread.  The centre cell and the NB nearest cells
// on either side of every item travel global -> shared with cp.async: bytes in flight that cost no registers (these
// kernels are bound by memory-level parallelism, and most arms are 0 or 1 long, so the common walk needs nothing
// else).  A thread reads back only its own slots: no barrier.
// Where the column pass of a round stores when the volume is re-partitioned right after it (one big pair over
// several GPUs, slab.py): row h of this disparity slab goes to the rank that owns row h, into its row slab
// [rows_r][W][g_total] at granule offset g_off -- peer memory over NVLink, no packing or staging afterwards.
constexpr int CS_MAX_PARTS = 8;
struct CsScatter {
    int nparts, g_off, g_total;
    int lo[CS_MAX_PARTS + 1];                 // row bounds of the parts
    float4 *base[CS_MAX_PARTS];
};

// Winner-take-all (pf:239-272) folded into the closing column pass of an aggregation: the 16 lanes that share a pixel
// reduce their first minimum over the 64 disparities the CTA holds and merge it into a per-pixel 64-bit key
// (order-preserving image of the cost << 32 | disparity) with one atomicMin per (pixel, granule group); the smallest
// key is the smallest cost and, among equal costs, the lowest disparity: np.argmin's first-minimum rule.  Saves
// re-reading the volume (4 B per cell).  keys must be preset to all ones; k_wta_decode turns them into the map.
struct CsWta {
    unsigned long long *keys;     // [H * W]
    int D;                        // disparities that exist (the pitch G * 4 may be larger)
    int store;                    // 0: the aggregated volume itself is not needed afterwards (right volume of match.py)
};
__device__ __forceinline__ unsigned cs_fkey(float f) {             // order-preserving: a < b  <=>  key(a) < key(b)
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <bool COLS, int ITEMS, int NB, bool SC = false, bool WT = false>
__global__ void __launch_bounds__(CS_THREADS, CS_MIN_BLOCKS) k_cbca_pass(const float4 *__restrict__ src, float4 *__restrict__ dst,
                                                          const uchar4 *__restrict__ arms, const int32_t *__restrict__ count,
                                                          int G, int H, int W, const CsScatter sc = CsScatter(),
                                                          const CsWta wt = CsWta()) {
    __shared__ float4 stage[2 * NB + 1][ITEMS][CS_THREADS];      // [centre, -1, +1, -2, +2, ..]
    constexpr int PH = 2 * ITEMS;
    // grid = (granule groups, patch columns, patch rows): the granule groups of a patch are consecutive CTAs, so a
    // pixel's 4*Dp bytes are fetched together (DRAM pages)
    const int bx = blockIdx.y, by = blockIdx.z;
    const int gi = threadIdx.x % CS_GC, g = blockIdx.x * CS_GC + gi;
    if (!WT && g >= G) return;                                   // (with the winner-take-all every lane stays for the shuffles)
    const bool gok = g < G;
    const ptrdiff_t stride = COLS ? (ptrdiff_t)W * G : (ptrdiff_t)G;
    size_t p[ITEMS];
    bool ok[ITEMS];
    uchar4 a[ITEMS];
    float n[ITEMS];
#pragma unroll
    for (int s = 0; s < ITEMS; s++) {
        const int pi = s * (CS_THREADS / CS_GC) + threadIdx.x / CS_GC;
        const int h = by * PH + pi / CS_PW, w = bx * CS_PW + pi % CS_PW;
        ok[s] = h < H && w < W && gok;
        p[s] = ok[s] ? (size_t)h * W + w : 0;
        if (ok[s]) {
            const float4 *c = src + p[s] * G + g;
            const int x = COLS ? h : w, lim = COLS ? H : W;
            cs_cp_async16(&stage[0][s][threadIdx.x], c);
#pragma unroll
            for (int k = 1; k <= NB; k++) {
                if (x - k >= 0) cs_cp_async16(&stage[2 * k - 1][s][threadIdx.x], c - k * stride);
                if (x + k < lim) cs_cp_async16(&stage[2 * k][s][threadIdx.x], c + k * stride);
            }
        }
        a[s] = arms[p[s]];
        n[s] = COLS ? (float)count[p[s]] : 1.0f;
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#pragma unroll
    for (int s = 0; s < ITEMS; s++) {
        if (!WT && !ok[s]) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[s]) {
            const float4 *c = src + p[s] * G + g;
            const int lo = COLS ? a[s].x : a[s].z, hi = COLS ? a[s].y : a[s].w;     // (up, down) | (left, right)
            cs_add(acc, stage[0][s][threadIdx.x]);                   // x, x-1, .., x-lo, then x+1, .., x+hi (pf:640-650)
#pragma unroll
            for (int k = 1; k <= NB; k++)
                if (lo >= k) cs_add(acc, stage[2 * k - 1][s][threadIdx.x]);
            for (int k = NB + 1; k <= lo; k++) cs_add(acc, c[-k * stride]);
#pragma unroll
            for (int k = 1; k <= NB; k++)
                if (hi >= k) cs_add(acc, stage[2 * k][s][threadIdx.x]);
            for (int k = NB + 1; k <= hi; k++) cs_add(acc, c[k * stride]);
            if (COLS) {
                const float vmax = fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)));
                const float vmin = fminf(fminf(fabsf(acc.x), fabsf(acc.y)), fminf(fabsf(acc.z), fabsf(acc.w)));
                if (vmax < 1e30f && vmin > 1e-30f) {
                    const float y = 1.0f / n[s];
                    acc = make_float4(cs_div1(acc.x, n[s], y), cs_div1(acc.y, n[s], y), cs_div1(acc.z, n[s], y), cs_div1(acc.w, n[s], y));
                } else {
                    acc = make_float4(acc.x / n[s], acc.y / n[s], acc.z / n[s], acc.w / n[s]);   // pf:161
                }
            }
        }
        if (WT) {
            // this lane's first minimum (strict <, cells d >= D do not exist; NaN and +inf never win, as in k_wta), as a
            // 64-bit key; then the minimum over the 16 lanes of the pixel: xor shuffles stay inside a half warp, and all 32
            // lanes take part (a reduction under a half-warp mask makes the two halves of the warp run one after the other).
            // Measured at C3: this pass + decode 0.46 ms, the plain pass + k_wta 0.46 ms -- at one granule per thread the
            // search costs what re-reading the volume costs; the chained calls close with k_cbca_close_g instead.
            const int d0 = g << 2;
            float best = CUDART_INF_F;
            int bd = 0;
            if (ok[s]) {
                if (acc.x < best) { best = acc.x; bd = d0; }
                if (d0 + 1 < wt.D && acc.y < best) { best = acc.y; bd = d0 + 1; }
                if (d0 + 2 < wt.D && acc.z < best) { best = acc.z; bd = d0 + 2; }
                if (d0 + 3 < wt.D && acc.w < best) { best = acc.w; bd = d0 + 3; }
            }
            unsigned long long key = best < CUDART_INF_F ? ((unsigned long long)cs_fkey(best + 0.0f) << 32 | (unsigned)bd) : ~0ull;   // (-0 counts as +0, like <)
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
                key = o < key ? o : key;
            }
            if (gi == 0 && key != ~0ull) atomicMin(wt.keys + p[s], key);
        }
        if (!ok[s]) continue;
        if (WT && !wt.store) continue;
        if (!SC) {
            dst[p[s] * G + g] = acc;
        } else {
            const int h = (int)(p[s] / W), w = (int)(p[s] - (size_t)h * W);
            int r = 0;
            while (r + 1 < sc.nparts && h >= sc.lo[r + 1]) r++;
            sc.base[r][((size_t)(h - sc.lo[r]) * W + w) * sc.g_total + sc.g_off + g] = acc;
        }
    }
}

// map = disparity of the smallest key (-1 where no finite cost exists, like k_wta), optionally the cost itself
__global__ void k_wta_decode(const unsigned long long *__restrict__ keys, float *__restrict__ disp, long long P) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const unsigned long long k = keys[p];
    disp[p] = k == ~0ull ? -1.0f : (float)(unsigned)(k & 0xffffffffu);
}

}  // namespace mccnn
