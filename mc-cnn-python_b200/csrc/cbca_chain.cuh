// Separable cross-based aggregation, fused ACROSS rounds.
//
// One round is  out_k(h,w) = ( sum_{h' in spine(h,w)} Hs_k(h',w) ) / |U(h,w)|,  Hs_k(h,w) = sum_{w' in arm(h,w)} out_{k-1}(h,w')
// (pf:640-650, :157-161).  As two streaming passes (cbca_stream.cuh) a round moves 16 B per cell through HBM: Hs_k out
// and back, out_k out and back.  k_cbca_colrow_g runs the column pass of round k and the row pass of round k+1 as ONE
// kernel with out_k in shared memory only:
//     Hs_k -> [out_k] -> Hs_{k+1},   8 B per cell per round;   a call of n rounds is  rows | (n-1) x colrow | close
// (rows = k_cbca_pass<rows>, close = k_cbca_close_g below: the last column pass in the same style).
// Fusing this way needs no vertical on-chip state (fusing the two passes of one round needs up to 27 row sums per
// column on chip).
//
// A CTA owns a segment of S pixels of ONE image row x 16 * GPT disparity granules (NT threads = NT / 16 pixel slots x
// 16 lanes; a thread handles granules gi, gi + 16, .. of its pixels).
//   load    ONE cp.async.bulk.tensor (TMA) per CTA brings Hs_k of rows h-1, h, h+1 for the segment and one halo pixel
//           per side: a 3 x NP x 64 GPT box; cells outside the volume (image borders, granules beyond Dp) arrive as
//           zeros and are never used.  No per-thread load instructions or addresses; an mbarrier publishes the box.
//           Meanwhile one thread per pixel reads the pixel's arms and |U| and leaves (arms, |U|, RN(1/|U|)) in shared
//           memory: the per-pixel work is done once, not by each of the 16 lanes.
//   column  out_k = (Hs_k(h) + up to `up` rows above + up to `down` rows below) / |U| into the tile T; rows beyond
//           +-1 (4 % of the arms of a natural image) come from global memory, two rows' loads in flight.
//   row     Hs_{k+1} = sum of out_k along the horizontal arm, from T; stored.
// The next row pass reaches a data-dependent halo left and right of the segment (at most distance_threshold - 1
// pixels, usually 0 or 1): one halo pixel per side is always computed; a segment whose arms reach further computes the
// far halo pixels straight from global memory into the (by then dead) ends of the h-1 / h+1 buffers.
// Several granules per thread (GPT): the per-pixel work (arm decode, addresses, loop control, branches) is paid once
// per 4 * GPT cells and the GPT accumulators are independent chains, so a warp needs fewer instructions per cell and
// stalls less per instruction; the price is GPT times the shared-memory tile.  GPT = 3 is the whole disparity row at
// ndisp 192.  Same additions in the same order as the two-pass form => bit-identical to it (tests: *_match_two_pass).
// L2 look-ahead: rows h-1 and h of a box were fetched by the CTAs of the rows above, row h+1 is first touched by this
// CTA; each CTA asks L2 (cp.async.bulk.prefetch.tensor) for the row that the CTA `ahead` rows below will be first to touch.
// Settled pixels: a pixel whose four arms are zero is its own region (out_k = Hs_k = out_{k-1} in every round); from the
// third pass over the volume on both ping-pong buffers hold its value, so it gets no work and no store, and the pixels
// that do have work are compacted into a list.
#pragma once
#include "cbca_stream.cuh"
#include "tc_common.cuh"

namespace mccnn {

__device__ __forceinline__ float4 cc_lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 cc_lds128u(unsigned a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void cc_sts128(unsigned a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cc_sts128u(unsigned a, const uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// the quotient of cs_divide with the reciprocal supplied (y = 1.0f / n, once per pixel)
__device__ __forceinline__ float4 cc_divide(float4 acc, const float n, const float y) {
    const float vmax = fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)));
    const float vmin = fminf(fminf(fabsf(acc.x), fabsf(acc.y)), fminf(fabsf(acc.z), fabsf(acc.w)));
    if (vmax < 1e30f && vmin > 1e-30f)
        return make_float4(cs_div1(acc.x, n, y), cs_div1(acc.y, n, y), cs_div1(acc.z, n, y), cs_div1(acc.w, n, y));
    return make_float4(acc.x / n, acc.y / n, acc.z / n, acc.w / n);
}

// map: float32 tensor {Dp, W, H} of src, box {64 * GPT, NP, 3}, no swizzle.
template <int S_, int NT_, int MINB_, int GPT_>
struct CgShape {
    static constexpr int S = S_, NT = NT_, MINB = MINB_, GPT = GPT_, HL = 1, SLOTS = NT / CS_GC, NP = S + 2 * HL;
    static constexpr int PB = GPT * 256;                                 // bytes per staged pixel
    static constexpr size_t SMEM = (size_t)3 * NP * PB + (size_t)NP * 16;
    static bool supports(int hm) { return hm - HL <= NP; }
};

// per-pixel information, one thread per staged pixel: (arms, |U|, RN(1/|U|)) into shared memory; returns whether some row
// arm of the segment leaves the staged halo (and how far, in lneed / rneed)
template <class C>
__device__ __forceinline__ bool cc_pixel_info(const uchar4 *__restrict__ arms, const int32_t *__restrict__ count, const size_t rowp,
                                              const int w0, const int tid, const int p_lo, const int np, const int sv,
                                              const unsigned sP, int &lneed, int &rneed, unsigned &packed) {
    constexpr int HL = C::HL;
    lneed = 0; rneed = 0; packed = 0;
    if (tid >= p_lo && tid < np) {
        const size_t q = rowp + w0 - HL + tid;
        const uchar4 a = arms[q];
        const float n = (float)count[q];
        packed = (unsigned)a.x | (unsigned)a.y << 8 | (unsigned)a.z << 16 | (unsigned)a.w << 24;
        cc_sts128u(sP + tid * 16, make_uint4(packed, __float_as_uint(n), __float_as_uint(1.0f / n), 0u));
        const int px = tid - HL;
        if (px >= 0 && px < sv) { lneed = (int)a.z - px; rneed = (int)a.w - (sv - 1 - px); }
    }
    return lneed > HL || rneed > HL;
}

template <int GPT>
__device__ __forceinline__ void cg_divide(float4 (&acc)[GPT], const float n, const float y) {
#pragma unroll
    for (int j = 0; j < GPT; j++) acc[j] = cc_divide(acc[j], n, y);
}

// acc[j] += c[k * stride + 16 j] for k = k0 .. k1 in that order (the granules jm names), two rows' loads in flight
template <int GPT>
__device__ __forceinline__ void cg_walk(float4 (&acc)[GPT], const float4 *__restrict__ c, const ptrdiff_t stride, int k0, const int k1,
                                        const unsigned jm) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (; k0 <= k1; k0 += 2) {
        const bool two = k0 + 1 <= k1;
        float4 v0[GPT], v1[GPT];
#pragma unroll
        for (int j = 0; j < GPT; j++) {
            const bool ok = (jm >> j) & 1;
            v0[j] = ok ? c[k0 * stride + CS_GC * j] : z;
            v1[j] = ok && two ? c[(k0 + 1) * stride + CS_GC * j] : z;
        }
#pragma unroll
        for (int j = 0; j < GPT; j++) {
            cs_add(acc[j], v0[j]);
            if (two) cs_add(acc[j], v1[j]);
        }
    }
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB) k_cbca_colrow_g(const __grid_constant__ CUtensorMap map,
                                                                  const __grid_constant__ CUtensorMap map_row,
                                                                  const float4 *__restrict__ src,
                                                                  float4 *__restrict__ dst, const uchar4 *__restrict__ arms,
                                                                  const int32_t *__restrict__ count, int G, int H, int W, int ahead,
                                                                  int settled) {
    constexpr int S = C::S, NP = C::NP, HL = C::HL, GPT = C::GPT, PB = C::PB, SLOTS = C::SLOTS;
    static_assert(NP <= 32 && C::NT >= 64, "one warp makes the per-pixel information; whole sweeps of the pixel slots");
    extern __shared__ __align__(128) unsigned char cc_raw[];
    __shared__ int reach[2];
    __shared__ int nlive;
    __shared__ unsigned char live[32];                                // staged pixels that have work this round, compacted
    __shared__ __align__(8) unsigned long long bar;
    const unsigned sU = (unsigned)__cvta_generic_to_shared(cc_raw);   // [NP][PB]  Hs_k(h-1)
    const unsigned sT = sU + NP * PB;                                 // [NP][PB]  Hs_k(h), then out_k
    const unsigned sP = sU + 3 * NP * PB;                             // [NP][16 B]  arms | |U| | 1/|U|
    const int tid = threadIdx.x, gi = tid & 15, slot = tid >> 4;
    const int g = blockIdx.x * (CS_GC * GPT) + gi, w0 = blockIdx.y * S, h = blockIdx.z;
    if (tid == 0) {
        tc_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        tc_mbar_expect_tx(&bar, 3 * NP * PB);
        tc_tma_load_3d(cc_raw, &map, blockIdx.x * (CS_GC * GPT * 4), w0 - HL, h - 1, &bar);
        // Rows h-1 and h were fetched by the CTAs of the rows above (L2 hits); row h+1 is first touched here, so the box
        // would wait for DRAM.  Ask L2 for the row that the CTA `ahead` rows below will be the first to touch: by the
        // time it runs, all three of its rows are L2 hits (map_row: the same tensor, box of one row).
        if (ahead > 0 && h + 1 + ahead < H)
            asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];\n" ::"l"(
                             reinterpret_cast<unsigned long long>(&map_row)),
                         "r"((int)(blockIdx.x * (CS_GC * GPT * 4))), "r"(w0 - HL), "r"(h + 1 + ahead)
                         : "memory");
        reach[0] = 0; reach[1] = 0;
    }
    unsigned jm = 0;                                                  // this thread's granules that exist
#pragma unroll
    for (int j = 0; j < GPT; j++) jm |= (g + CS_GC * j < G ? 1u : 0u) << j;
    const int sv = min(S, W - w0);
    const int np = min(NP, W - w0 + HL);
    const int p_lo = w0 > 0 ? 0 : HL;
    const ptrdiff_t stride = (ptrdiff_t)W * G;
    const size_t rowp = (size_t)h * W;
    const float4 *cbase = src + (rowp + w0 - HL) * G + g;              // staged pixel 0, this lane's first granule
    const unsigned my = gi * 16;
    int lneed, rneed;
    // (the last warp, so that its loads are not queued behind thread 0's TMA set-up)
    unsigned packed;
    const int ti = tid - (C::NT - 32);
    const bool far = cc_pixel_info<C>(arms, count, rowp, w0, ti, p_lo, np, sv, sP, lneed, rneed, packed);
    if (ti >= 0) {
        // A pixel whose four arms are all zero is its own region: out_k = Hs_k = out_{k-1}, bit for bit, in every round.
        // From the third pass over the volume on (settled) both ping-pong buffers hold that value: T already is out_k and
        // there is nothing to store -- on a natural image that is every second pixel.  The pixels that do have work are
        // compacted into a list, so the two phases sweep over half as many.
        const bool on = ti >= p_lo && ti < np && !(settled && packed == 0);
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) live[__popc(m & ((1u << ti) - 1u))] = (unsigned char)ti;
        if (ti == 0) nlive = __popc(m);
    }
    const int any_far = __syncthreads_or(far);                        // publishes the per-pixel information and the mbarrier
    tc_mbar_wait(&bar, 0);

    // ---- column phase: out_k of the staged pixels into T
    const int nl_ = nlive;
    if (jm) {
#pragma unroll 1
        for (int i = slot; i < nl_; i += SLOTS) {
            const int p = live[i];
            const unsigned t = sT + p * PB + my;
            const uint4 pi = cc_lds128u(sP + p * 16);
            const int up = pi.x & 0xff, down = (pi.x >> 8) & 0xff;
            float4 acc[GPT];
#pragma unroll
            for (int j = 0; j < GPT; j++) { acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); cs_add(acc[j], cc_lds128(t + j * 256)); }
            if (up >= 1) {
#pragma unroll
                for (int j = 0; j < GPT; j++) cs_add(acc[j], cc_lds128(t - NP * PB + j * 256));
                if (up >= 2) cg_walk<GPT>(acc, cbase + (ptrdiff_t)p * G, -stride, 2, up, jm);
            }
            if (down >= 1) {
#pragma unroll
                for (int j = 0; j < GPT; j++) cs_add(acc[j], cc_lds128(t + NP * PB + j * 256));
                if (down >= 2) cg_walk<GPT>(acc, cbase + (ptrdiff_t)p * G, stride, 2, down, jm);
            }
            cg_divide<GPT>(acc, __uint_as_float(pi.y), __uint_as_float(pi.z));
#pragma unroll
            for (int j = 0; j < GPT; j++) cc_sts128(t + j * 256, acc[j]);
        }
    }
    if (any_far) {
        // rare: out_k of the halo pixels beyond the staged one, straight from global memory, into the ends of the
        // h-1 / h+1 buffers next to T (every warp is past its column phase after this barrier)
        if (far) { atomicMax(&reach[0], lneed); atomicMax(&reach[1], rneed); }
        __syncthreads();
        const int nl = max(reach[0] - HL, 0), nr = max(reach[1] - HL, 0);
        for (int it = tid; it < (nl + nr) * CS_GC; it += C::NT) {
            const int q = it >> 4;
            const int fx = q < nl ? -HL - 1 - q : S + HL + (q - nl);               // segment-relative column
            const int x = w0 + fx;
            if (!jm || x < 0 || x >= W) continue;
            const uchar4 a = arms[rowp + x];
            const float n = (float)count[rowp + x];
            const float4 *c = src + (rowp + x) * G + g;
            float4 acc[GPT];
#pragma unroll
            for (int j = 0; j < GPT; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            cg_walk<GPT>(acc, c, 0, 0, 0, jm);
            cg_walk<GPT>(acc, c, -stride, 1, a.x, jm);
            cg_walk<GPT>(acc, c, stride, 1, a.y, jm);
            cg_divide<GPT>(acc, n, 1.0f / n);
#pragma unroll
            for (int j = 0; j < GPT; j++) cc_sts128(sT + (fx + HL) * PB + my + j * 256, acc[j]);
        }
    }
    __syncthreads();

    // ---- row phase: Hs_{k+1} of the segment from shared memory
    if (jm) {
        float4 *out0 = dst + (rowp + w0) * G + g;
#pragma unroll 1
        for (int i = slot; i < nl_; i += SLOTS) {
            const int p = live[i], px = p - HL;
            if (px < 0 || px >= sv) continue;                             // (a halo pixel)
            const unsigned t0 = sT + p * PB + my;
            float4 *out = out0 + (ptrdiff_t)px * G;
            unsigned a;
            asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(a) : "r"(sP + p * 16));
            float4 acc[GPT];
#pragma unroll
            for (int j = 0; j < GPT; j++) { acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); cs_add(acc[j], cc_lds128(t0 + j * 256)); }
            const unsigned tl = t0 - ((a >> 16) & 0xff) * PB, tr = t0 + (a >> 24) * PB;
#pragma unroll 1
            for (unsigned t = t0; t != tl;) {
                t -= PB;
#pragma unroll
                for (int j = 0; j < GPT; j++) cs_add(acc[j], cc_lds128(t + j * 256));
            }
#pragma unroll 1
            for (unsigned t = t0; t != tr;) {
                t += PB;
#pragma unroll
                for (int j = 0; j < GPT; j++) cs_add(acc[j], cc_lds128(t + j * 256));
            }
#pragma unroll
            for (int j = 0; j < GPT; j++)
                if ((jm >> j) & 1) out[CS_GC * j] = acc[j];
        }
    }
}

// The closing pass of a chained call in the same style: out_n = colsum(Hs_n) / |U| for a segment, straight to the volume,
// with the winner-take-all of pf:239-272 folded in (wt.keys != NULL).  A thread holds 4 * GPT disparities of a pixel, so
// the first-minimum search costs a third of what it does with one granule per thread (k_cbca_pass<cols, WT>), and at
// ndisp <= 64 * GPT one CTA sees the whole disparity row.  The smallest cost of the pixel over the CTA's disparities, then
// the first disparity that has it (cells d >= D do not exist; NaN and +inf never win, -0 = +0: k_wta's rules), merged into
// the per-pixel 64-bit key with one atomicMin.  Pixels without arms (armless != 0: the call had >= 2 rounds, so `dst`
// already holds their value, see k_cbca_colrow_g) are read for the minimum but neither recomputed nor stored.
template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB) k_cbca_close_g(const __grid_constant__ CUtensorMap map,
                                                                 const __grid_constant__ CUtensorMap map_row,
                                                                 const float4 *__restrict__ src, float4 *__restrict__ dst,
                                                                 const uchar4 *__restrict__ arms, const int32_t *__restrict__ count,
                                                                 int G, int H, int W, int ahead, int armless, int store, const CsWta wt) {
    constexpr int S = C::S, NP = C::NP, HL = C::HL, GPT = C::GPT, PB = C::PB, SLOTS = C::SLOTS;
    extern __shared__ __align__(128) unsigned char cc_raw[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned sU = (unsigned)__cvta_generic_to_shared(cc_raw);
    const unsigned sT = sU + NP * PB, sP = sU + 3 * NP * PB;
    const int tid = threadIdx.x, gi = tid & 15, slot = tid >> 4;
    const int g = blockIdx.x * (CS_GC * GPT) + gi, w0 = blockIdx.y * S, h = blockIdx.z;
    if (tid == 0) {
        tc_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        tc_mbar_expect_tx(&bar, 3 * NP * PB);
        tc_tma_load_3d(cc_raw, &map, blockIdx.x * (CS_GC * GPT * 4), w0 - HL, h - 1, &bar);
        if (ahead > 0 && h + 1 + ahead < H)
            asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];\n" ::"l"(
                             reinterpret_cast<unsigned long long>(&map_row)),
                         "r"((int)(blockIdx.x * (CS_GC * GPT * 4))), "r"(w0 - HL), "r"(h + 1 + ahead)
                         : "memory");
    }
    unsigned jm = 0;
#pragma unroll
    for (int j = 0; j < GPT; j++) jm |= (g + CS_GC * j < G ? 1u : 0u) << j;
    const int sv = min(S, W - w0);
    const int np = min(NP, W - w0 + HL);
    const int p_lo = w0 > 0 ? 0 : HL;
    const ptrdiff_t stride = (ptrdiff_t)W * G;
    const size_t rowp = (size_t)h * W;
    const float4 *cbase = src + (rowp + w0 - HL) * G + g;
    const unsigned my = gi * 16;
    int lneed, rneed;
    unsigned packed;
    cc_pixel_info<C>(arms, count, rowp, w0, tid - (C::NT - 32), p_lo, np, sv, sP, lneed, rneed, packed);
    __syncthreads();                                                  // publishes the per-pixel information and the mbarrier
    tc_mbar_wait(&bar, 0);

    float4 *out0 = dst + (rowp + w0) * G + g;
    unsigned long long *keys = wt.keys ? wt.keys + rowp + w0 : nullptr;
    // (every lane makes every sweep: the shuffles below need the whole warp)
#pragma unroll 1
    for (int px0 = 0; px0 < sv; px0 += SLOTS) {
        const int px = px0 + slot, p = px + HL;
        const bool valid = px < sv;
        float4 acc[GPT];
        bool keep = false;                                            // armless and settled: nothing to compute or store
        if (valid) {
            const unsigned t = sT + p * PB + my;
            const uint4 pi = cc_lds128u(sP + p * 16);
            keep = armless && pi.x == 0;
            const int up = pi.x & 0xff, down = (pi.x >> 8) & 0xff;
#pragma unroll
            for (int j = 0; j < GPT; j++) { acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); cs_add(acc[j], cc_lds128(t + j * 256)); }
            if (!keep && jm) {
                if (up >= 1) {
#pragma unroll
                    for (int j = 0; j < GPT; j++) cs_add(acc[j], cc_lds128(t - NP * PB + j * 256));
                    if (up >= 2) cg_walk<GPT>(acc, cbase + (ptrdiff_t)p * G, -stride, 2, up, jm);
                }
                if (down >= 1) {
#pragma unroll
                    for (int j = 0; j < GPT; j++) cs_add(acc[j], cc_lds128(t + NP * PB + j * 256));
                    if (down >= 2) cg_walk<GPT>(acc, cbase + (ptrdiff_t)p * G, stride, 2, down, jm);
                }
                cg_divide<GPT>(acc, __uint_as_float(pi.y), __uint_as_float(pi.z));
#pragma unroll
                for (int j = 0; j < GPT; j++)
                    if (store && ((jm >> j) & 1)) out0[(ptrdiff_t)px * G + CS_GC * j] = acc[j];
            }
        }
        if (keys) {
            float m = CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < GPT; j++) {
                const int d0 = (g + CS_GC * j) << 2;
                const bool on = valid && ((jm >> j) & 1);
                if (!on) acc[j] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
                if (d0 + 1 >= wt.D) acc[j].y = CUDART_INF_F;
                if (d0 + 2 >= wt.D) acc[j].z = CUDART_INF_F;
                if (d0 + 3 >= wt.D) acc[j].w = CUDART_INF_F;
                m = fminf(m, fminf(fminf(acc[j].x, acc[j].y), fminf(acc[j].z, acc[j].w)));
            }
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, off));
            unsigned bd = 0xffffffffu;
#pragma unroll
            for (int j = GPT - 1; j >= 0; j--) {
                const unsigned d0 = (unsigned)(g + CS_GC * j) << 2;
                bd = acc[j].w == m ? d0 + 3 : bd;
                bd = acc[j].z == m ? d0 + 2 : bd;
                bd = acc[j].y == m ? d0 + 1 : bd;
                bd = acc[j].x == m ? d0 : bd;
            }
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) bd = min(bd, __shfl_xor_sync(0xffffffffu, bd, off));
            if (gi == 0 && valid && m < CUDART_INF_F) atomicMin(keys + px, (unsigned long long)cs_fkey(m + 0.0f) << 32 | bd);
        }
    }
}

}  // namespace mccnn
