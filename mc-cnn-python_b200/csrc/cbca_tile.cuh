// Separable cross-based aggregation round, shared-memory staged (fast mode of mccnn_cbca).
//
// The support region of pf:571-657 is "for every spine row h' in [h-up, h+down]: the horizontal
// arm of (h', w)", so one round is exactly  out(h,w) = ( sum_{h'} Hs(h',w) ) / |U(h,w)|  with
// Hs(h',w) = sum_{w' in arm(h',w)} in(h',w').  Row sums are formed once and re-used down each
// column: <= 27 + 27 additions per cell instead of <= 729.  Only the association of the float32
// sum differs from the reference ((row sums) summed, instead of one running sum): ~1e-7 relative.
//
// HBM-bound stage (8 B per cell per round), but three things keep a straightforward kernel far
// from that bound (each measured, see profiles/): the dependent-load latency of walking arms in
// global memory; warp divergence (arm lengths are heavy tailed: most are 0-1, a few are 13, so a
// warp whose lanes are 32 neighbouring pixels almost always waits for one long walk); and the
// per-(pixel, granule) instruction overhead of short data-dependent loops.  Hence:
//   k_cbca_tile_meta (once per call): for every 16x8 pixel tile
//     - the halo its regions actually need: rows above/below (largest up/down arm reaching out of
//       the tile) and, per needed row, pixels left/right (largest left/right arm of that row's 8
//       spine pixels), plus the row offsets of the packed shared-memory image.  Natural images
//       need 0-2 halo pixels, flat ones 13;
//     - a lane schedule: the (row, pixel) items of the tile grouped into 8 bank classes (the slot
//       of the pixel in the staged image mod 8) and, inside each class, sorted by arm length.
//   k_cbca_round_tile: persistent CTAs (3 per SM) walk the work items (tile, slab of 4 disparity
//       granules = 16 disparities = 64 B per pixel) in launch order with a two-stage software
//       pipeline: while item i is being summed, all input cells of item i+1 (tile + halo, arms,
//       counts) and the schedule of item i+2 are in flight as cp.async copies into the other
//       shared-memory buffer, so the HBM latency of a tile is never exposed (every cell of the
//       volume is read from HBM once; halo re-reads hit L2).
//       One thread = one (row, pixel) item with all 4 granules in registers: the walk, its loop
//       and index overhead are shared by 16 disparities, and the float32 adds are issued as
//       packed FADD2.  The 8 lanes of a quarter warp take items of the 8 different bank classes
//       (pixel stride is 5 float4, so class c, granule k lives in bank group (5c + k) mod 8 and
//       every 16-byte access is conflict free at any walk offset), and the lanes of a warp take
//       items of equal rank in the sorted classes, so walks inside a warp have similar length.
//       Phase A writes row sums to a dense [row][8] image, phase B adds them along the spine,
//       divides by |U| and stores 64 contiguous bytes per pixel with two 256-bit stores.
//       A tile whose halo does not fit the shared-memory budget (flat image areas) runs the same
//       code over 2 or 1 granules at a time instead of 4 (the extra sub-passes are not pipelined).
//   The division by the integer |U| <= 729 is a*y refined by one FMA remainder step with
//   y = RN(1/|U|): correctly rounded (Markstein), identical to the IEEE division the reference
//   performs (checked against IEEE division for every |U| over 1e8 numerators); non-finite, zero
//   and extreme magnitudes take the plain division.
#pragma once
#include "common.cuh"

namespace mccnn {

constexpr int CT_TH = 16, CT_TW = 8, CT_GPMAX = 4, CT_THREADS = 128, CT_NW = CT_THREADS / 32;
constexpr int CT_CTAS_PER_SM = 3;
constexpr int CT_MAXARM = 13;                                  // distance_threshold <= 14 in this mode
constexpr int CT_MAXROWS = CT_TH + 2 * CT_MAXARM;              // 42
constexpr int CT_IN_BYTES = 23 * 1024;                         // one staged tile + halo image
constexpr int CT_HS_BYTES = 14 * 1024;                         // dense row sums [row][8]
constexpr int CT_NONE = 0xff;

__host__ __device__ constexpr int ct_stride(int gp) { return gp == 1 ? 1 : gp + 1; }   // float4 per staged pixel

struct __align__(16) CbcaTileMeta {                            // 672 bytes
    uint8_t up, down, nrows, gp;                               // halo rows, rows staged, granules per thread per sub-pass
    uint16_t r0, total_px;                                     // first staged row, staged pixels
    uint8_t left[CT_MAXROWS + 2], right[CT_MAXROWS + 2];       // per staged row: halo pixels
    uint16_t centre[CT_MAXROWS + 2];                           // per staged row: slot of the tile's first column
    uint8_t permA[8][CT_MAXROWS + 2];                          // [bank class][rank] -> staged row, longest row arms first
    uint8_t permB[CT_TW][CT_TH];                               // [tile column][rank] -> tile row, longest column arms first
    uint8_t pad[8];
};
static_assert(sizeof(CbcaTileMeta) == 672, "CbcaTileMeta layout");

struct __align__(16) CtStage {                                 // everything one work item needs, filled by cp.async
    float4 in[CT_IN_BYTES / 16];
    uchar4 arms[CT_MAXROWS][CT_TW];
    int32_t count[CT_TH][CT_TW];
};
struct __align__(16) CtSmem {
    CtStage stage[2];
    float4 hs[CT_HS_BYTES / 16];
    CbcaTileMeta meta[3];
};
constexpr int CT_SMEM_BYTES = (int)sizeof(CtSmem);

__global__ void __launch_bounds__(64) k_cbca_tile_meta(const uchar4 *__restrict__ arms, CbcaTileMeta *__restrict__ meta,
                                                       int H, int W) {
    __shared__ int s_up, s_down;
    __shared__ CbcaTileMeta m;
    __shared__ uint8_t key[CT_MAXROWS + 2][CT_TW];
    const int tid = threadIdx.x;
    const int w0 = blockIdx.x * CT_TW, h0 = blockIdx.y * CT_TH;
    const int wend = min(w0 + CT_TW, W), hend = min(h0 + CT_TH, H), tw = wend - w0;
    if (tid == 0) { s_up = 0; s_down = 0; }
    __syncthreads();
    for (int i = tid; i < CT_TH * CT_TW; i += 64) {
        const int h = h0 + i / CT_TW, w = w0 + i % CT_TW;
        if (h < hend && w < wend) {
            const uchar4 a = arms[(size_t)h * W + w];
            const int nu = (int)a.x - (h - h0), nd = (int)a.y - (hend - 1 - h);
            if (nu > 0) atomicMax(&s_up, nu);
            if (nd > 0) atomicMax(&s_down, nd);
        }
    }
    __syncthreads();
    const int r0 = h0 - s_up;                                   // arms never leave the image (pf:585, :593)
    const int nrows = hend + s_down - r0;
    if (tid < nrows) {
        int L = 0, R = 0;
        for (int w = w0; w < wend; w++) {
            const uchar4 a = arms[(size_t)(r0 + tid) * W + w];
            L = max(L, (int)a.z - (w - w0));
            R = max(R, (int)a.w - (wend - 1 - w));
            key[tid][w - w0] = (uint8_t)(a.z + a.w);
        }
        m.left[tid] = (uint8_t)L;
        m.right[tid] = (uint8_t)R;
    }
    __syncthreads();
    if (tid == 0) {
        int off = 0;
        for (int r = 0; r < nrows; r++) {
            m.centre[r] = (uint16_t)(off + m.left[r]);
            off += tw + m.left[r] + m.right[r];
        }
        m.up = (uint8_t)s_up; m.down = (uint8_t)s_down; m.nrows = (uint8_t)nrows;
        m.r0 = (uint16_t)r0; m.total_px = (uint16_t)off;
        int gp = CT_GPMAX;                                      // staged image and dense row sums must fit
        while (gp > 1 && (off * ct_stride(gp) * 16 > CT_IN_BYTES || nrows * CT_TW * ct_stride(gp) * 16 > CT_HS_BYTES))
            gp >>= 1;
        m.gp = (uint8_t)gp;
    }
    __syncthreads();
    if (tid < 8) {
        // phase A schedule of bank class q: one candidate per staged row, pixel (q - centre[r]) mod 8
        const int q = tid;
        int n = 0;
        for (int r = 0; r < nrows; r++) {
            const int px = (q - m.centre[r]) & 7;
            if (px >= tw) continue;
            const int k = key[r][px];
            int i = n++;
            while (i > 0 && key[m.permA[q][i - 1]][(q - m.centre[m.permA[q][i - 1]]) & 7] < k) {
                m.permA[q][i] = m.permA[q][i - 1];
                i--;
            }
            m.permA[q][i] = (uint8_t)r;
        }
        for (; n < CT_MAXROWS + 2; n++) m.permA[q][n] = CT_NONE;
    } else if (tid < 8 + CT_TW) {
        // phase B schedule of tile column px: tile rows sorted by up + down
        const int px = tid - 8;
        int n = 0;
        uint8_t kk[CT_TH];
        if (px < tw) {
            for (int rt = 0; rt < hend - h0; rt++) {
                const uchar4 a = arms[(size_t)(h0 + rt) * W + w0 + px];
                const int k = a.x + a.y;
                int i = n++;
                while (i > 0 && kk[i - 1] < k) {
                    m.permB[px][i] = m.permB[px][i - 1];
                    kk[i] = kk[i - 1];
                    i--;
                }
                m.permB[px][i] = (uint8_t)rt;
                kk[i] = (uint8_t)k;
            }
        }
        for (; n < CT_TH; n++) m.permB[px][n] = CT_NONE;
    }
    __syncthreads();
    uint32_t *dst = reinterpret_cast<uint32_t *>(meta + (size_t)blockIdx.y * gridDim.x + blockIdx.x);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&m);
    for (int i = tid; i < (int)(sizeof(CbcaTileMeta) / 4); i += 64) dst[i] = src[i];
}

__device__ __forceinline__ void ct_cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void ct_cp_async4(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void ct_cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

struct ct_f4 { float2 lo, hi; };                                // one granule as two packed pairs

__device__ __forceinline__ void ct_add(ct_f4 &acc, const float4 v) {
    acc.lo = __fadd2_rn(acc.lo, make_float2(v.x, v.y));
    acc.hi = __fadd2_rn(acc.hi, make_float2(v.z, v.w));
}

__device__ __forceinline__ void ct_st256(float4 *dst, const ct_f4 a, const ct_f4 b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(a.lo.x), "f"(a.lo.y), "f"(a.hi.x),
                 "f"(a.hi.y), "f"(b.lo.x), "f"(b.lo.y), "f"(b.hi.x), "f"(b.hi.y)
                 : "memory");
}

// a / n for n = |U| (an integer <= 729), y = RN(1/n): q = RN(a*y); r = a - n*q (exact, FMA); RN(q + r*y).
__device__ __forceinline__ float2 ct_div2(float2 a, float2 nn, float2 y) {      // nn = (-n, -n)
    const float2 q = __fmul2_rn(a, y);
    const float2 r = __ffma2_rn(nn, q, a);
    return __ffma2_rn(r, y, q);
}

struct CtGeom { int G, H, W, tilesX, nslab, gp_top; };
struct CtItem { int tile, w0, h0, tw, gslab; };

__device__ __forceinline__ CtItem ct_item(int it, const CtGeom &ge) {
    CtItem x;
    x.tile = it / ge.nslab;
    x.gslab = (it - x.tile * ge.nslab) * ge.gp_top;
    const int ty = x.tile / ge.tilesX, tx = x.tile - ty * ge.tilesX;
    x.w0 = tx * CT_TW; x.h0 = ty * CT_TH;
    x.tw = min(x.w0 + CT_TW, ge.W) - x.w0;
    return x;
}

// cp.async the cells of one sub-pass (GP granules from g0, ng live) of an item: one warp per row,
// lanes = (pixel, granule), granule fastest.  No wait here.
template <int GP>
__device__ __forceinline__ void ct_stage_cells(const float4 *__restrict__ in, float4 *in_s, const CbcaTileMeta &m,
                                               const CtItem &x, int g0, int ng, const CtGeom &ge) {
    constexpr int ST = ct_stride(GP), PPW = 32 / GP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = lane % GP, sub = lane / GP;
    if (k >= ng) return;
    const int nrows = m.nrows;
    const float4 *base = in + ((size_t)m.r0 * ge.W + x.w0 + sub) * ge.G + g0 + k;
    for (int r = warp; r < nrows; r += CT_NW) {
        const int L = m.left[r];
        const int npx = x.tw + L + m.right[r];
        const float4 *src = base + ((ptrdiff_t)r * ge.W - L) * (ptrdiff_t)ge.G;
        float4 *dst = in_s + (m.centre[r] - L + sub) * ST + k;
        for (int px = sub; px < npx; px += PPW) {
            ct_cp_async16(dst, src);
            src += (size_t)PPW * ge.G;
            dst += PPW * ST;
        }
    }
}
__device__ __forceinline__ void ct_stage_cells_gp(int gp, const float4 *__restrict__ in, float4 *in_s,
                                                  const CbcaTileMeta &m, const CtItem &x, int g0, int ng,
                                                  const CtGeom &ge) {
    if (gp == 4) ct_stage_cells<4>(in, in_s, m, x, g0, ng, ge);
    else if (gp == 2) ct_stage_cells<2>(in, in_s, m, x, g0, ng, ge);
    else ct_stage_cells<1>(in, in_s, m, x, g0, ng, ge);
}

// arms of the staged rows and |U| of the tile pixels (4-byte cp.async: no alignment assumption on W)
__device__ __forceinline__ void ct_stage_aux(const uchar4 *__restrict__ arms, const int32_t *__restrict__ count,
                                             CtStage &st, const CbcaTileMeta &m, const CtItem &x, const CtGeom &ge) {
    const int tid = threadIdx.x;
    const int n = m.nrows * CT_TW;
    for (int i = tid; i < n; i += CT_THREADS) {
        const int r = i / CT_TW, px = i % CT_TW;
        if (px < x.tw) ct_cp_async4(&st.arms[r][px], arms + (size_t)(m.r0 + r) * ge.W + x.w0 + px);
    }
    {
        const int r = tid / CT_TW, px = tid % CT_TW;
        if (x.h0 + r < ge.H && px < x.tw) ct_cp_async4(&st.count[r][px], count + (size_t)(x.h0 + r) * ge.W + x.w0 + px);
    }
}

// phase A + B of one sub-pass over GP granules per pixel starting at granule g0 (ng <= GP of them live);
// the caller has made the staged cells visible (wait + barrier).
template <int GP>
__device__ __forceinline__ void ct_compute(float4 *__restrict__ out, const CtStage &st, float4 *hs_s, const CbcaTileMeta &m,
                                           const CtItem &x, int g0, int ng, const CtGeom &ge) {
    constexpr int ST = ct_stride(GP);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nrows = m.nrows;
    const int q = lane & 7, rk = lane >> 3;                      // bank class / tile column, rank within the warp step
    // ---- phase A: row sums, order w, w-1, .., w-left, w+1, .., w+right (pf:645-650)
    for (int t = warp; t * 4 < nrows; t += CT_NW) {
        const int r = m.permA[q][t * 4 + rk];
        if (r == CT_NONE) continue;
        const int c = m.centre[r];
        const int px = (q - c) & 7;
        const uchar4 a = st.arms[r][px];
        const float4 *p0 = st.in + (c + px) * ST;
        ct_f4 acc[GP];
#pragma unroll
        for (int k = 0; k < GP; k++) acc[k].lo = acc[k].hi = make_float2(0.f, 0.f);
        const float4 *p = p0;
#pragma unroll 1
        for (int j = a.z; j >= 0; j--, p -= ST) {
#pragma unroll
            for (int k = 0; k < GP; k++) ct_add(acc[k], p[k]);
        }
        p = p0 + ST;
#pragma unroll 1
        for (int j = a.w; j > 0; j--, p += ST) {
#pragma unroll
            for (int k = 0; k < GP; k++) ct_add(acc[k], p[k]);
        }
        float4 *d = hs_s + (r * CT_TW + px) * ST;
#pragma unroll
        for (int k = 0; k < GP; k++) d[k] = make_float4(acc[k].lo.x, acc[k].lo.y, acc[k].hi.x, acc[k].hi.y);
    }
    __syncthreads();

    // ---- phase B: column sums in spine order h, h-1, .., h-up, h+1, .., h+down (pf:640-644), / |U|
    const int rt = m.permB[q][warp * 4 + rk];                    // 16 ranks = 4 warps x 4
    if (rt != CT_NONE) {
        const int px = q, r = rt + m.up;
        const uchar4 a = st.arms[r][px];
        const float4 *p0 = hs_s + (r * CT_TW + px) * ST;
        ct_f4 acc[GP];
#pragma unroll
        for (int k = 0; k < GP; k++) acc[k].lo = acc[k].hi = make_float2(0.f, 0.f);
        const float4 *p = p0;
#pragma unroll 1
        for (int j = a.x; j >= 0; j--, p -= CT_TW * ST) {
#pragma unroll
            for (int k = 0; k < GP; k++) ct_add(acc[k], p[k]);
        }
        p = p0 + CT_TW * ST;
#pragma unroll 1
        for (int j = a.y; j > 0; j--, p += CT_TW * ST) {
#pragma unroll
            for (int k = 0; k < GP; k++) ct_add(acc[k], p[k]);
        }
        const float n = (float)st.count[rt][px], y = 1.0f / n;
        float hi = 0.f, lo = 3e38f;
#pragma unroll
        for (int k = 0; k < GP; k++) {
            hi = fmaxf(hi, fmaxf(fmaxf(fabsf(acc[k].lo.x), fabsf(acc[k].lo.y)), fmaxf(fabsf(acc[k].hi.x), fabsf(acc[k].hi.y))));
            lo = fminf(lo, fminf(fminf(fabsf(acc[k].lo.x), fabsf(acc[k].lo.y)), fminf(fabsf(acc[k].hi.x), fabsf(acc[k].hi.y))));
        }
        if (hi < 1e30f && lo > 1e-30f) {
            const float2 nn = make_float2(-n, -n), yy = make_float2(y, y);
#pragma unroll
            for (int k = 0; k < GP; k++) { acc[k].lo = ct_div2(acc[k].lo, nn, yy); acc[k].hi = ct_div2(acc[k].hi, nn, yy); }
        } else {
#pragma unroll
            for (int k = 0; k < GP; k++) {                       // pf:161
                acc[k].lo = make_float2(acc[k].lo.x / n, acc[k].lo.y / n);
                acc[k].hi = make_float2(acc[k].hi.x / n, acc[k].hi.y / n);
            }
        }
        float4 *dst = out + ((size_t)(x.h0 + rt) * ge.W + x.w0 + px) * ge.G + g0;
        if (GP >= 2 && ng == GP && !(ge.G & 1)) {
#pragma unroll
            for (int k = 0; k < GP; k += 2) ct_st256(dst + k, acc[k], acc[k + 1 < GP ? k + 1 : k]);
        } else {
#pragma unroll
            for (int k = 0; k < GP; k++)
                if (k < ng) dst[k] = make_float4(acc[k].lo.x, acc[k].lo.y, acc[k].hi.x, acc[k].hi.y);
        }
    }
}
__device__ __forceinline__ void ct_compute_gp(int gp, float4 *__restrict__ out, const CtStage &st, float4 *hs_s,
                                              const CbcaTileMeta &m, const CtItem &x, int g0, int ng, const CtGeom &ge) {
    if (gp == 4) ct_compute<4>(out, st, hs_s, m, x, g0, ng, ge);
    else if (gp == 2) ct_compute<2>(out, st, hs_s, m, x, g0, ng, ge);
    else ct_compute<1>(out, st, hs_s, m, x, g0, ng, ge);
}

// Persistent: CTA b takes work items b, b + gridDim.x, ...; item = (tile, slab), the slabs of a tile
// adjacent in the order so that the pieces of a pixel's disparity row are read and written close
// together in time.  gp_top = granules per slab (4, or the next power of two >= G for small ndisp).
__global__ void __launch_bounds__(CT_THREADS, CT_CTAS_PER_SM)
k_cbca_round_tile(const float4 *__restrict__ in, float4 *__restrict__ out, const uchar4 *__restrict__ arms,
                  const int32_t *__restrict__ count, const CbcaTileMeta *__restrict__ meta, int G, int H, int W,
                  int tilesX, int nitems, int nslab, int gp_top) {
    extern __shared__ __align__(16) unsigned char ct_raw[];
    CtSmem &sm = *reinterpret_cast<CtSmem *>(ct_raw);
    const int tid = threadIdx.x;
    CtGeom ge;
    ge.G = G; ge.H = H; ge.W = W; ge.tilesX = tilesX; ge.nslab = nslab; ge.gp_top = gp_top;
    const int first = blockIdx.x, step = gridDim.x;
    if (first >= nitems) return;

    auto fetch_meta = [&](int it, int slot) {                    // 42 x 16 B
        const float4 *src = reinterpret_cast<const float4 *>(meta + it / nslab);
        float4 *dst = reinterpret_cast<float4 *>(&sm.meta[slot]);
        if (tid < (int)(sizeof(CbcaTileMeta) / 16)) ct_cp_async16(dst + tid, src + tid);
    };
    auto stage_item = [&](int it, int n) {                       // first sub-pass of item `it` (its n-th of this CTA)
        const CbcaTileMeta &m = sm.meta[n % 3];
        const CtItem x = ct_item(it, ge);
        const int gp = min((int)m.gp, gp_top);
        ct_stage_cells_gp(gp, in, sm.stage[n & 1].in, m, x, x.gslab, min(gp, G - x.gslab), ge);
        ct_stage_aux(arms, count, sm.stage[n & 1], m, x, ge);
    };

    // prologue: schedule of items 0 and 1, cells of item 0
    fetch_meta(first, 0);
    if (first + step < nitems) fetch_meta(first + step, 1);
    ct_cp_async_wait_all();
    __syncthreads();
    stage_item(first, 0);

    int n = 0;
    for (int it = first; it < nitems; it += step, n++) {
        ct_cp_async_wait_all();
        __syncthreads();                                        // item n staged, schedule n+1 present, buffers of n-1 free
        if (it + step < nitems) stage_item(it + step, n + 1);
        if (it + 2 * step < nitems) fetch_meta(it + 2 * step, (n + 2) % 3);
        const CbcaTileMeta &m = sm.meta[n % 3];
        const CtItem x = ct_item(it, ge);
        const int gp = min((int)m.gp, gp_top);
        CtStage &st = sm.stage[n & 1];
        ct_compute_gp(gp, out, st, sm.hs, m, x, x.gslab, min(gp, G - x.gslab), ge);
        // remaining sub-passes of a tile that did not fit at gp_top granules (not pipelined)
        for (int sub = gp; sub < gp_top && x.gslab + sub < G; sub += gp) {
            const int g0 = x.gslab + sub;
            __syncthreads();                                    // everyone done with st.in and hs
            ct_stage_cells_gp(gp, in, st.in, m, x, g0, min(gp, G - g0), ge);
            ct_cp_async_wait_all();
            __syncthreads();
            ct_compute_gp(gp, out, st, sm.hs, m, x, g0, min(gp, G - g0), ge);
        }
    }
    ct_cp_async_wait_all();
}

}  // namespace mccnn
