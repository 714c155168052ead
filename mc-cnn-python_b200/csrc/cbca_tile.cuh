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
//   k_cbca_round_tile: persistent CTAs walk the tiles; for each tile they walk the disparity range
//       in slabs of 4 granules (16 disparities, 64 B per pixel) with a two-stage TMA pipeline:
//       while slab s is being summed, the cells of slab s+1 (or of the next tile's first slab)
//       land in the other shared-memory buffer through per-row cp.async.bulk.tensor box copies
//       signalled on an mbarrier, so the HBM latency of a tile is never exposed and staging costs
//       one instruction per row instead of address arithmetic per 16 bytes (every cell of the
//       volume is read from HBM once; halo re-reads hit L2).  Out-of-image halo columns are
//       zero-filled by the TMA unit and never summed (arms stop at the border, pf:585-620).
//       One thread = one (row, pixel) item with all 4 granules in registers; its arm lengths and
//       shared-memory offsets are decoded ONCE per tile and reused for every slab; the float32
//       adds are issued as packed FADD2.  The 8 lanes of a quarter warp take items of the 8
//       different bank classes (the box is 5 float4 wide per pixel, the 5th being padding that
//       makes class c, granule k live in bank group (5c + k) mod 8, so every 16-byte access is
//       conflict free at any walk offset), and the lanes of a warp take items of equal rank in
//       the sorted classes, so walks inside a warp have similar length.
//       Phase A writes row sums to a dense [row][8] image, phase B adds them along the spine,
//       divides by |U| and stores 64 contiguous bytes per pixel with two 256-bit stores.
//       A tile whose halo does not fit the shared-memory budget (flat image areas) runs the same
//       code over 2 or 1 granules per slab instead of 4.
//   The division by the integer |U| <= 729 is a*y refined by one FMA remainder step with
//   y = RN(1/|U|): correctly rounded (Markstein), identical to the IEEE division the reference
//   performs (checked against IEEE division for every |U| over 1e8 numerators); non-finite, zero
//   and extreme magnitudes take the plain division.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace mccnn {

constexpr int CT_TH = 16, CT_TW = 8, CT_GPMAX = 4, CT_THREADS = 128, CT_NW = CT_THREADS / 32;
constexpr int CT_ASTEPS = (11 + CT_NW - 1) / CT_NW;   // phase A warp steps a warp may have to take
constexpr int CT_CTAS_PER_SM = 3;
constexpr int CT_MAXARM = 13;                                  // distance_threshold <= 14 in this mode
constexpr int CT_MAXROWS = CT_TH + 2 * CT_MAXARM;              // 42
constexpr int CT_ROWS_PAD = CT_MAXROWS + 2;                    // 44
constexpr int CT_IN_BYTES = 27 * 1024;                         // one staged slab (tile + halo), rows 128 B aligned
constexpr int CT_HS_BYTES = 13 * 1024;                         // dense row sums [row][8]
constexpr int CT_NONE = 0xff;
constexpr int CT_NLEVEL = 3;                                   // gp = 4, 2, 1
constexpr int CT_NHALO = 6;                                    // staged row halo is rounded up to one of these
__host__ __device__ constexpr int ct_halo(int i) { return i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 2 : i == 3 ? 4 : i == 4 ? 8 : 13; }
__host__ __device__ constexpr int ct_level(int gp) { return gp == 4 ? 0 : gp == 2 ? 1 : 2; }
// float4 per staged pixel: gp granules + one of padding (bank spreading), clamped to the volume's pitch
__host__ __device__ constexpr int ct_stride(int gp, int G) { return gp == 1 ? 1 : (gp + 1 < G ? gp + 1 : G); }

struct CtMaps { CUtensorMap m[CT_NLEVEL][CT_NHALO]; };         // box = [4*stride floats][8 + 2*halo pixels][1 row]

struct __align__(16) CbcaTileMeta {                            // 3088 bytes, copied to shared memory as is
    uint8_t up, down, nrows, gp;                               // halo rows, rows staged, granules per slab
    uint16_t r0, w0, h0;                                       // first staged row, tile origin
    uint8_t tw, th;                                            // valid tile extent
    uint32_t stage_bytes;                                      // bytes one slab's box copies deliver
    uint8_t hwi[CT_ROWS_PAD];                                  // per staged row: halo class (ct_halo)
    uint16_t rowb[CT_ROWS_PAD];                                // per staged row: base, in 16-byte units
    uint16_t centre[CT_ROWS_PAD];                              // per staged row: tile's first column, in 16-byte units
    uint8_t permA[8][CT_ROWS_PAD];                             // [bank class][rank] -> staged row, longest row arms first
    uint8_t permB[CT_TW][CT_TH];                               // [tile column][rank] -> tile row, longest column arms first
    uchar4 arms[CT_MAXROWS][CT_TW];                            // arms of the staged rows at the tile's columns
    float2 cnt[CT_TH][CT_TW];                                  // (|U|, RN(1/|U|)) of the tile's pixels
};
static_assert(sizeof(CbcaTileMeta) == 3088, "CbcaTileMeta layout");

struct __align__(128) CtSmem {
    unsigned char in[2][CT_IN_BYTES];
    float4 hs[CT_HS_BYTES / 16];
    CbcaTileMeta ti[2];
    unsigned long long mbar[2];
    int next[2];
};
constexpr int CT_SMEM_BYTES = (int)sizeof(CtSmem);

// pixel of bank class q in a row whose first tile column sits at 16-byte unit c: (c + st*px) = q (mod 8)
__device__ __forceinline__ int ct_class_px(int q, int c, int st) {
    return (st & 1) ? ((st * (q - c)) & 7) : ((q - c) & 7);      // 1, 3, 5 are their own inverses mod 8
}

__global__ void __launch_bounds__(64) k_cbca_tile_meta(const uchar4 *__restrict__ arms, const int32_t *__restrict__ count,
                                                       CbcaTileMeta *__restrict__ meta, int G, int H, int W, int gp_top) {
    __shared__ int s_up, s_down;
    __shared__ CbcaTileMeta m;
    __shared__ uint8_t key[CT_ROWS_PAD][CT_TW];
    __shared__ uint8_t need[CT_ROWS_PAD];
    const int tid = threadIdx.x;
    const int w0 = blockIdx.x * CT_TW, h0 = blockIdx.y * CT_TH;
    const int wend = min(w0 + CT_TW, W), hend = min(h0 + CT_TH, H), tw = wend - w0;
    if (tid == 0) { s_up = 0; s_down = 0; }
    __syncthreads();
    for (int i = tid; i < CT_TH * CT_TW; i += 64) {
        const int h = h0 + i / CT_TW, w = w0 + i % CT_TW;
        float2 c = make_float2(1.f, 1.f);
        if (h < hend && w < wend) {
            const uchar4 a = arms[(size_t)h * W + w];
            const int nu = (int)a.x - (h - h0), nd = (int)a.y - (hend - 1 - h);
            if (nu > 0) atomicMax(&s_up, nu);
            if (nd > 0) atomicMax(&s_down, nd);
            const float n = (float)count[(size_t)h * W + w];
            c = make_float2(n, 1.0f / n);
        }
        m.cnt[i / CT_TW][i % CT_TW] = c;
    }
    __syncthreads();
    const int r0 = h0 - s_up;                                   // arms never leave the image (pf:585, :593)
    const int nrows = hend + s_down - r0;
    if (tid < nrows) {
        int L = 0, R = 0;
        for (int px = 0; px < CT_TW; px++) {
            uchar4 a = make_uchar4(0, 0, 0, 0);
            if (px < tw) {
                a = arms[(size_t)(r0 + tid) * W + w0 + px];
                L = max(L, (int)a.z - px);
                R = max(R, (int)a.w - (tw - 1 - px));
            }
            m.arms[tid][px] = a;
            key[tid][px] = (uint8_t)(a.z + a.w);
        }
        int hi = 0;
        while (ct_halo(hi) < max(L, R)) hi++;
        need[tid] = (uint8_t)hi;
    }
    __syncthreads();
    if (tid == 0) {
        int gp = gp_top, st = 1, off = 0;
        for (;; gp >>= 1) {                                     // staged slab and dense row sums must fit
            st = ct_stride(gp, G);
            off = 0;
            for (int r = 0; r < nrows; r++) off += ((CT_TW + 2 * ct_halo(need[r])) * st + 7) & ~7;
            if (gp == 1 || (off * 16 <= CT_IN_BYTES && nrows * CT_TW * st * 16 <= CT_HS_BYTES)) break;
        }
        off = 0;
        unsigned bytes = 0;
        for (int r = 0; r < nrows; r++) {
            const int bw = CT_TW + 2 * ct_halo(need[r]);
            m.hwi[r] = need[r];
            m.rowb[r] = (uint16_t)off;
            m.centre[r] = (uint16_t)(off + ct_halo(need[r]) * st);
            off += (bw * st + 7) & ~7;
            bytes += (unsigned)bw * st * 16;
        }
        m.up = (uint8_t)s_up; m.down = (uint8_t)s_down; m.nrows = (uint8_t)nrows; m.gp = (uint8_t)gp;
        m.r0 = (uint16_t)r0; m.w0 = (uint16_t)w0; m.h0 = (uint16_t)h0;
        m.tw = (uint8_t)tw; m.th = (uint8_t)(hend - h0);
        m.stage_bytes = bytes;
    }
    __syncthreads();
    if (tid < 8) {
        // phase A schedule of bank class q: one candidate pixel per staged row
        const int q = tid, st = ct_stride(m.gp, G);
        int n = 0;
        uint8_t kk[CT_ROWS_PAD];
        for (int r = 0; r < nrows; r++) {
            const int px = ct_class_px(q, m.centre[r], st);
            if (px >= tw) continue;
            const int k = key[r][px];
            int i = n++;
            while (i > 0 && kk[i - 1] < k) {
                m.permA[q][i] = m.permA[q][i - 1];
                kk[i] = kk[i - 1];
                i--;
            }
            m.permA[q][i] = (uint8_t)r;
            kk[i] = (uint8_t)k;
        }
        for (; n < CT_ROWS_PAD; n++) m.permA[q][n] = CT_NONE;
    } else if (tid < 8 + CT_TW) {
        // phase B schedule of tile column px: tile rows sorted by up + down
        const int px = tid - 8;
        int n = 0;
        uint8_t kk[CT_TH];
        if (px < tw) {
            for (int rt = 0; rt < hend - h0; rt++) {
                const uchar4 a = m.arms[rt + s_up][px];
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

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ct_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(ct_smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void ct_cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void ct_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(ct_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ct_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(ct_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ct_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CT_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
        "@p bra CT_DONE_%=;\n"
        "bra CT_WAIT_%=;\n"
        "CT_DONE_%=:\n"
        "}\n" ::"r"(ct_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void ct_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// box copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void ct_tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2,
                                               unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
            "r"(ct_smem_u32(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2),
        "r"(ct_smem_u32(bar))
        : "memory");
}

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

// per-thread work of one tile, decoded once and reused for every slab
struct CtWork {
    unsigned a_in[CT_ASTEPS], a_hs[CT_ASTEPS];      // phase A items: 16-byte-unit offsets of the centre cell / of the row-sum cell
    unsigned a_lr[CT_ASTEPS];               // left | right << 8 | valid << 16
    unsigned b_hs, b_ud;            // phase B item: row-sum cell; up | down << 8 | valid << 16
    float b_n, b_y;
    float4 *b_out;                  // output cell of slab 0
};

// warp 0: issue the box copies of one slab (granules from g0) of tile `m` into `dst`
__device__ __forceinline__ void ct_issue_slab(const CtMaps &maps, const CbcaTileMeta &m, int g0, int G,
                                              unsigned char *dst, unsigned long long *bar) {
    const int lane = threadIdx.x & 31;
    const int level = ct_level(m.gp);
    for (int r = lane; r < m.nrows; r += 32) {
        const int hi = m.hwi[r];
        ct_tma_load_3d(dst + (size_t)m.rowb[r] * 16, &maps.m[level][hi], 4 * g0, (int)m.w0 - ct_halo(hi), (int)m.r0 + r, bar);
    }
    if (lane == 0) ct_mbar_expect_tx(bar, m.stage_bytes);
}

// acc = sum of the n cells of a walk, in the reference's order: p0, p0 - step, .., p0 - l*step, p0 + step, .., the
// adds strictly sequential per element (same rounding as a plain loop).
template <int GP>
__device__ __forceinline__ void ct_walk(ct_f4 (&acc)[GP], const float4 *p0, int step, int l, int n) {
#pragma unroll
    for (int k = 0; k < GP; k++) acc[k].lo = acc[k].hi = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int i = 0; i < n; i += 2) {
        const int i1 = i + 1;
        const bool two = i1 < n;
        const int o0 = (i <= l) ? -i : i - l;
        const int o1 = two ? ((i1 <= l) ? -i1 : i1 - l) : o0;
        const float4 *pa = p0 + o0 * step, *pb = p0 + o1 * step;
        float4 a[GP], b[GP];
#pragma unroll
        for (int k = 0; k < GP; k++) a[k] = pa[k];
#pragma unroll
        for (int k = 0; k < GP; k++) b[k] = pb[k];
#pragma unroll
        for (int k = 0; k < GP; k++) ct_add(acc[k], a[k]);
        if (two) {
#pragma unroll
            for (int k = 0; k < GP; k++) ct_add(acc[k], b[k]);
        }
    }
}

template <int GP>
__device__ __forceinline__ void ct_slab(const float4 *in_s, float4 *hs_s, const CtWork &wk, int st, int slab, int ng, int G) {
    // ---- phase A: row sums, order w, w-1, .., w-left, w+1, .., w+right (pf:645-650).  One loop over the
    //      left + right + 1 cells of the walk (lanes of a warp have similar totals, not similar splits),
    //      two cells per trip with the loads issued ahead of the adds.
#pragma unroll
    for (int i = 0; i < CT_ASTEPS; i++) {
        if (!(wk.a_lr[i] >> 16)) continue;
        const float4 *p0 = in_s + wk.a_in[i];
        const int l = wk.a_lr[i] & 0xff, n = l + ((wk.a_lr[i] >> 8) & 0xff) + 1;
        ct_f4 acc[GP];
        ct_walk<GP>(acc, p0, st, l, n);
        float4 *d = hs_s + wk.a_hs[i];
#pragma unroll
        for (int k = 0; k < GP; k++)                             // (a pixel has min(GP + 1, G) cells: never write past ng)
            if (k < ng) d[k] = make_float4(acc[k].lo.x, acc[k].lo.y, acc[k].hi.x, acc[k].hi.y);
    }
    __syncthreads();

    // ---- phase B: column sums in spine order h, h-1, .., h-up, h+1, .., h+down (pf:640-644), / |U|
    if (wk.b_ud >> 16) {
        const float4 *p0 = hs_s + wk.b_hs;
        const int u = wk.b_ud & 0xff, cells = u + ((wk.b_ud >> 8) & 0xff) + 1;
        ct_f4 acc[GP];
        ct_walk<GP>(acc, p0, CT_TW * st, u, cells);
        const float n = wk.b_n, y = wk.b_y;
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
        float4 *dst = wk.b_out + slab * GP;
        if (GP >= 2 && ng == GP && !(G & 1)) {
#pragma unroll
            for (int k = 0; k < GP; k += 2) ct_st256(dst + k, acc[k], acc[k + 1 < GP ? k + 1 : k]);
        } else {
#pragma unroll
            for (int k = 0; k < GP; k++)
                if (k < ng) dst[k] = make_float4(acc[k].lo.x, acc[k].lo.y, acc[k].hi.x, acc[k].hi.y);
        }
    }
}

// Persistent: CTA b starts with tile b and then takes tiles from a global counter (tiles differ a lot in
// cost: flat image areas have long arms), two tiles ahead of the one being summed.  Neighbouring tiles run
// at the same time on different SMs, so halo re-reads hit L2.  Each tile is walked over all its slabs.
__global__ void __launch_bounds__(CT_THREADS, CT_CTAS_PER_SM)
k_cbca_round_tile(const __grid_constant__ CtMaps maps, float4 *__restrict__ out, const CbcaTileMeta *__restrict__ meta,
                  int G, int H, int W, int ntiles, unsigned *__restrict__ counter) {
    extern __shared__ __align__(128) unsigned char ct_raw[];
    CtSmem &sm = *reinterpret_cast<CtSmem *>(ct_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int first = blockIdx.x, step = gridDim.x;
    if (first >= ntiles) return;

    auto fetch_ti = [&](int tile, int slot) {                    // 193 x 16 B
        const float4 *src = reinterpret_cast<const float4 *>(meta + tile);
        float4 *dst = reinterpret_cast<float4 *>(&sm.ti[slot]);
        for (int i = tid; i < (int)(sizeof(CbcaTileMeta) / 16); i += CT_THREADS) ct_cp_async16(dst + i, src + i);
    };

    if (tid == 0) {
        ct_mbar_init(&sm.mbar[0], 1);
        ct_mbar_init(&sm.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid == 0) sm.next[0] = (int)atomicAdd(counter, 1u) + step;
    fetch_ti(first, 0);
    ct_cp_async_wait_all();
    __syncthreads();
    unsigned c = 0;                                             // slabs issued so far: buffer c & 1, phase (c >> 1) & 1
    if (warp == 0) ct_issue_slab(maps, sm.ti[0], 0, G, sm.in[0], &sm.mbar[0]);

    int tile = first;
    for (int n = 0; tile < ntiles; n++) {
        const CbcaTileMeta &m = sm.ti[n & 1];
        const int next_tile = sm.next[n & 1];
        const bool has_next = next_tile < ntiles;
        if (has_next) fetch_ti(next_tile, (n + 1) & 1);          // lands while this tile's slabs are summed
        int after_next = 0;
        if (tid == 0) after_next = (int)atomicAdd(counter, 1u) + step;
        const int gp = m.gp, st = ct_stride(gp, G);
        const int nslab = (G + gp - 1) / gp;

        // ---- decode this thread's items once per tile
        CtWork wk;
        const int q = lane & 7, rk = lane >> 3;
#pragma unroll
        for (int i = 0; i < CT_ASTEPS; i++) {
            const int t = warp + i * CT_NW;
            wk.a_lr[i] = 0; wk.a_in[i] = 0; wk.a_hs[i] = 0;
            if (t * 4 < m.nrows) {
                const int r = m.permA[q][t * 4 + rk];
                if (r != CT_NONE) {
                    const int cc = m.centre[r];
                    const int px = ct_class_px(q, cc, st);
                    const uchar4 a = m.arms[r][px];
                    wk.a_in[i] = cc + px * st;
                    wk.a_hs[i] = (r * CT_TW + px) * st;
                    wk.a_lr[i] = a.z | (a.w << 8) | (1u << 16);
                }
            }
        }
        {
            const int rt = warp < 4 ? m.permB[q][warp * 4 + rk] : CT_NONE;   // 16 ranks = 4 warps x 4
            wk.b_ud = 0; wk.b_hs = 0; wk.b_n = 1.f; wk.b_y = 1.f; wk.b_out = out;
            if (rt != CT_NONE) {
                const int r = rt + m.up;
                const uchar4 a = m.arms[r][q];
                wk.b_hs = (r * CT_TW + q) * st;
                wk.b_ud = a.x | (a.y << 8) | (1u << 16);
                wk.b_n = m.cnt[rt][q].x; wk.b_y = m.cnt[rt][q].y;
                wk.b_out = out + ((size_t)(m.h0 + rt) * W + m.w0 + q) * G;
            }
        }

        for (int s = 0; s < nslab; s++, c++) {
            ct_mbar_wait(&sm.mbar[c & 1], (c >> 1) & 1);         // this slab's cells have landed
            ct_cp_async_wait_all();                             // (this thread's part of the next tile's schedule)
            if (tid == 0 && s == nslab - 1) sm.next[(n + 1) & 1] = after_next;
            __syncthreads();                                    // previous slab fully consumed; next schedule visible
            if (warp == 0) {
                if (s + 1 < nslab) ct_issue_slab(maps, m, (s + 1) * gp, G, sm.in[(c + 1) & 1], &sm.mbar[(c + 1) & 1]);
                else if (has_next) ct_issue_slab(maps, sm.ti[(n + 1) & 1], 0, G, sm.in[(c + 1) & 1], &sm.mbar[(c + 1) & 1]);
            }
            const float4 *in_s = reinterpret_cast<const float4 *>(sm.in[c & 1]);
            const int ng = min(gp, G - s * gp);
            if (gp == 4) ct_slab<4>(in_s, sm.hs, wk, st, s, ng, G);
            else if (gp == 2) ct_slab<2>(in_s, sm.hs, wk, st, s, ng, G);
            else ct_slab<1>(in_s, sm.hs, wk, st, s, ng, G);
        }
        tile = next_tile;
    }
}

}  // namespace mccnn
