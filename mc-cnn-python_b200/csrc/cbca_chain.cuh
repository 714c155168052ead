// Separable cross-based aggregation, fused ACROSS rounds.
//
// One round is  out_k(h,w) = ( sum_{h' in spine(h,w)} Hs_k(h',w) ) / |U(h,w)|,  Hs_k(h,w) = sum_{w' in arm(h,w)} out_{k-1}(h,w')
// (pf:640-650, :157-161).  As two streaming passes (cbca_stream.cuh) a round moves 16 B per cell through HBM: Hs_k out
// and back, out_k out and back.  k_cbca_colrow runs the column pass of round k and the row pass of round k+1 as ONE
// kernel with out_k in shared memory only:
//     Hs_k -> [out_k] -> Hs_{k+1},   8 B per cell per round;   a call of n rounds is  rows | (n-1) x colrow | cols.
// Fusing this way needs no vertical on-chip state (fusing the two passes of one round needs up to 27 row sums per
// column on chip).
//
// A CTA owns a segment of S pixels of ONE image row x 16 disparity granules (NT threads = NT / 16 pixel slots x 16 lanes).
//   load    every (pixel, granule) item of the segment and of one halo pixel per side: Hs_k of rows h-1, h, h+1 by
//           cp.async into shared memory, unconditionally (no dependent step; rows h-1 / h+1 are other CTAs' centre
//           rows, i.e. L2 hits).  Meanwhile one thread per pixel reads the pixel's arms and |U| and leaves
//           (arms, |U|, RN(1/|U|)) in shared memory: the per-pixel work is done once, not by each of the 16 lanes.
//   column  out_k = (Hs_k(h) + up to `up` rows above + up to `down` rows below) / |U| into the tile T; rows beyond
//           +-1 (4 % of the arms of a natural image) come from global memory, four loads at a time.
//   row     Hs_{k+1} = sum of out_k along the horizontal arm, from T; stored.
// The next row pass reaches a data-dependent halo left and right of the segment (at most distance_threshold - 1
// pixels, usually 0 or 1): one halo pixel per side is always computed; a segment whose arms reach further computes the
// far halo pixels straight from global memory (one segment in ten).
// Same additions in the same order as the two-pass form => bit-identical to it (tests: *_match_two_pass).
#pragma once
#include "cbca_stream.cuh"
#include "tc_common.cuh"

namespace mccnn {

// NP = S + 2 * HL staged pixels (the segment + HL halo pixels per side) = four sweeps of the NT / 16 pixel slots
template <int S_, int NT_, int MINB_, int HL_>
struct CcShape {
    static constexpr int S = S_, NT = NT_, MINB = MINB_, HL = HL_, SLOTS = NT / CS_GC, NP = S + 2 * HL;
    // dynamic shared memory: Hs_k(h-1) | Hs_k(h) -> out_k | Hs_k(h+1) | per-pixel info.  The far halo of out_k (rare, up to
    // arm limit - 1 pixels per side) is written over the neighbouring ends of the h-1 / h+1 buffers, which are dead by then.
    static constexpr size_t SMEM = (size_t)3 * NP * 256 + (size_t)NP * 16;
    static bool supports(int hm) { return hm - HL <= NP; }
};
typedef CcShape<30, 128, 8, 1> CcNarrow;     // 25 KB, 8 CTAs of 4 warps per SM (the one used: shorter waits at the two barriers)
typedef CcShape<62, 256, 4, 1> CcWide;       // 50 KB, 4 CTAs of 8 warps per SM

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
__device__ __forceinline__ void cc_cp16(unsigned smem, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(g) : "memory");
}

// the quotient of cs_divide with the reciprocal supplied (y = 1.0f / n, once per pixel)
__device__ __forceinline__ float4 cc_divide(float4 acc, const float n, const float y) {
    const float vmax = fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)));
    const float vmin = fminf(fminf(fabsf(acc.x), fabsf(acc.y)), fminf(fabsf(acc.z), fabsf(acc.w)));
    if (vmax < 1e30f && vmin > 1e-30f)
        return make_float4(cs_div1(acc.x, n, y), cs_div1(acc.y, n, y), cs_div1(acc.z, n, y), cs_div1(acc.w, n, y));
    return make_float4(acc.x / n, acc.y / n, acc.z / n, acc.w / n);
}

// acc + c[k0 * stride] + .. + c[k1 * stride] added in that order, the loads of four steps issued together
__device__ __noinline__ float4 cc_walk(float4 acc, const float4 *__restrict__ c, const ptrdiff_t stride, int k0, const int k1) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (; k0 <= k1; k0 += 4) {
        const float4 v0 = c[k0 * stride];
        const float4 v1 = k0 + 1 <= k1 ? c[(k0 + 1) * stride] : z;
        const float4 v2 = k0 + 2 <= k1 ? c[(k0 + 2) * stride] : z;
        const float4 v3 = k0 + 3 <= k1 ? c[(k0 + 3) * stride] : z;
        cs_add(acc, v0);
        if (k0 + 1 <= k1) cs_add(acc, v1);
        if (k0 + 2 <= k1) cs_add(acc, v2);
        if (k0 + 3 <= k1) cs_add(acc, v3);
    }
    return acc;
}

// rows 2 .. n of an arm (n >= 2): rows 2 and 3 inline with both loads in flight (most long arms end there), the rest by cc_walk
__device__ __forceinline__ float4 cc_far_rows(float4 acc, const float4 *__restrict__ c, const ptrdiff_t stride, const int n) {
    const float4 v2 = c[2 * stride];
    if (n >= 3) {
        const float4 v3 = c[3 * stride];
        cs_add(acc, v2);
        cs_add(acc, v3);
        if (n >= 4) acc = cc_walk(acc, c, stride, 4, n);
    } else {
        cs_add(acc, v2);
    }
    return acc;
}

// ---- the three phases after the load, shared by the two kernels below (they differ in how the rows get into shared memory)

// column phase: out_k of the staged pixels p = slot, slot + SLOTS, ..  (p_lo <= p < np) into T
template <class C>
__device__ __forceinline__ void cc_column_phase(const float4 *__restrict__ cbase, const ptrdiff_t stride, const int G, const unsigned sT,
                                                const unsigned sP, const unsigned my, const int slot, const int p_lo, const int np) {
    constexpr int NP = C::NP, SLOTS = C::SLOTS;
    unsigned t = sT + slot * 256 + my, pa = sP + slot * 16;
#pragma unroll 1
    for (int p = slot; p < np; p += SLOTS, t += SLOTS * 256, pa += SLOTS * 16) {
        if (p < p_lo) continue;
        const uint4 pi = cc_lds128u(pa);
        const int up = pi.x & 0xff, down = (pi.x >> 8) & 0xff;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        cs_add(acc, cc_lds128(t));                                             // h, h-1, .., h-up, then h+1, .., h+down
        if (up >= 1) cs_add(acc, cc_lds128(t - NP * 256));
        if (up >= 2) acc = cc_far_rows(acc, cbase + (ptrdiff_t)p * G, -stride, up);
        if (down >= 1) cs_add(acc, cc_lds128(t + NP * 256));
        if (down >= 2) acc = cc_far_rows(acc, cbase + (ptrdiff_t)p * G, stride, down);
        cc_sts128(t, cc_divide(acc, __uint_as_float(pi.y), __uint_as_float(pi.z)));
    }
}

// rare: out_k of the halo pixels beyond the staged one, straight from global memory, into the ends of the
// h-1 / h+1 buffers next to T (every warp is past its column phase when this runs)
template <class C>
__device__ __forceinline__ void cc_far_halo(const float4 *__restrict__ src, const uchar4 *__restrict__ arms,
                                            const int32_t *__restrict__ count, const int nl, const int nr, const int tid,
                                            const bool gok, const int g, const int G, const int W, const int w0, const size_t rowp,
                                            const ptrdiff_t stride, const unsigned sT, const unsigned my) {
    constexpr int S = C::S, HL = C::HL;
    for (int it = tid; it < (nl + nr) * CS_GC; it += C::NT) {
        const int q = it >> 4;
        const int fx = q < nl ? -HL - 1 - q : S + HL + (q - nl);               // segment-relative column
        const int x = w0 + fx;
        if (!gok || x < 0 || x >= W) continue;
        const uchar4 a = arms[rowp + x];
        const float n = (float)count[rowp + x];
        const float4 *c = src + (rowp + x) * G + g;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        cs_add(acc, c[0]);
        acc = cc_walk(acc, c, -stride, 1, a.x);
        acc = cc_walk(acc, c, stride, 1, a.y);
        cc_sts128(sT + (fx + HL) * 256 + my, cc_divide(acc, n, 1.0f / n));
    }
}

// row phase: Hs_{k+1} of the segment from shared memory
template <class C>
__device__ __forceinline__ void cc_row_phase(float4 *__restrict__ dst, const size_t rowp, const int w0, const int g, const int G,
                                             const unsigned sT, const unsigned sP, const unsigned my, const int slot, const int sv) {
    constexpr int SLOTS = C::SLOTS, HL = C::HL;
    char *out = reinterpret_cast<char *>(dst + (rowp + w0 + slot) * G + g);
    const ptrdiff_t stepB = (ptrdiff_t)SLOTS * G * 16;
    unsigned t0 = sT + (slot + HL) * 256 + my, pa = sP + (slot + HL) * 16;
#pragma unroll 1
    for (int px = slot; px < sv; px += SLOTS, t0 += SLOTS * 256, pa += SLOTS * 16, out += stepB) {
        unsigned a;
        asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(a) : "r"(pa));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        cs_add(acc, cc_lds128(t0));                                            // w, w-1, .., w-left, then w+1, .., w+right
        const unsigned tl = t0 - ((a >> 16) & 0xff) * 256, tr = t0 + (a >> 24) * 256;
#pragma unroll 1
        for (unsigned t = t0; t != tl;) { t -= 256; cs_add(acc, cc_lds128(t)); }
#pragma unroll 1
        for (unsigned t = t0; t != tr;) { t += 256; cs_add(acc, cc_lds128(t)); }
        *reinterpret_cast<float4 *>(out) = acc;
    }
}

// per-pixel information, one thread per staged pixel: (arms, |U|, RN(1/|U|)) into shared memory; returns whether some row
// arm of the segment leaves the staged halo (and how far, in lneed / rneed)
template <class C>
__device__ __forceinline__ bool cc_pixel_info(const uchar4 *__restrict__ arms, const int32_t *__restrict__ count, const size_t rowp,
                                              const int w0, const int tid, const int p_lo, const int np, const int sv,
                                              const unsigned sP, int &lneed, int &rneed) {
    constexpr int HL = C::HL;
    lneed = 0; rneed = 0;
    if (tid >= p_lo && tid < np) {
        const size_t q = rowp + w0 - HL + tid;
        const uchar4 a = arms[q];
        const float n = (float)count[q];
        cc_sts128u(sP + tid * 16, make_uint4((unsigned)a.x | (unsigned)a.y << 8 | (unsigned)a.z << 16 | (unsigned)a.w << 24,
                                             __float_as_uint(n), __float_as_uint(1.0f / n), 0u));
        const int px = tid - HL;
        if (px >= 0 && px < sv) { lneed = (int)a.z - px; rneed = (int)a.w - (sv - 1 - px); }
    }
    return lneed > HL || rneed > HL;
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB) k_cbca_colrow(const float4 *__restrict__ src, float4 *__restrict__ dst,
                                                                const uchar4 *__restrict__ arms, const int32_t *__restrict__ count,
                                                                int G, int H, int W) {
    constexpr int S = C::S, NP = C::NP, SLOTS = C::SLOTS, HL = C::HL;
    extern __shared__ __align__(128) unsigned char cc_raw[];
    __shared__ int reach[2];
    // staged pixel p = 0 .. NP-1 is image column w0 - HL + p
    const unsigned sU = (unsigned)__cvta_generic_to_shared(cc_raw);   // [NP][256 B]  Hs_k(h-1)
    const unsigned sT = sU + NP * 256;                                // [NP][256 B]  Hs_k(h), then out_k
    const unsigned sD = sT + NP * 256;                                // [NP][256 B]  Hs_k(h+1)
    const unsigned sP = sD + NP * 256;                                // [NP][16 B]   arms | |U| | 1/|U|
    const int tid = threadIdx.x, gi = tid & 15, slot = tid >> 4;
    const int g = blockIdx.x * CS_GC + gi, w0 = blockIdx.y * S, h = blockIdx.z;
    const bool gok = g < G;
    const int sv = min(S, W - w0);                                    // valid pixels of the segment
    const int np = min(NP, W - w0 + HL);                              // staged pixels that exist on the right
    const int p_lo = w0 > 0 ? 0 : HL;                                 // ... and on the left
    const ptrdiff_t stride = (ptrdiff_t)W * G;
    const size_t rowp = (size_t)h * W;
    const float4 *cbase = src + (rowp + w0 - HL) * G + g;              // staged pixel 0, this lane's granule (not dereferenced if outside)
    const unsigned my = gi * 16;
    if (tid == 0) { reach[0] = 0; reach[1] = 0; }

    // ---- load: rows h-1, h, h+1 of every staged (pixel, granule) item, unconditionally (running pointers: the 64-bit
    //      address arithmetic is done once per thread, not once per item)
    if (gok) {
        const char *c = reinterpret_cast<const char *>(cbase + (ptrdiff_t)slot * G);
        const ptrdiff_t strideB = stride * 16, stepB = (ptrdiff_t)SLOTS * G * 16;
        const bool has_up = h >= 1, has_dn = h + 1 < H;
        unsigned t = sT + slot * 256 + my;
#pragma unroll
        for (int i = 0; i < (NP + SLOTS - 1) / SLOTS; i++, c += stepB, t += SLOTS * 256) {
            const int p = slot + SLOTS * i;
            if (p >= p_lo && p < np) {
                cc_cp16(t, c);
                if (has_up) cc_cp16(t - NP * 256, c - strideB);
                if (has_dn) cc_cp16(t + NP * 256, c + strideB);
            }
        }
    }
    int lneed, rneed;
    const bool far = cc_pixel_info<C>(arms, count, rowp, w0, tid, p_lo, np, sv, sP, lneed, rneed);
    // (the barrier only publishes the per-pixel information: it comes BEFORE the wait for the staged rows, so that a warp
    //  whose own rows have arrived does not wait for the slowest warp's)
    const int any_far = __syncthreads_or(far);
    asm volatile("cp.async.wait_all;\n" ::: "memory");

    if (gok) cc_column_phase<C>(cbase, stride, G, sT, sP, my, slot, p_lo, np);   // each thread in the slots it loaded itself
    if (any_far) {
        if (far) { atomicMax(&reach[0], lneed); atomicMax(&reach[1], rneed); }
        __syncthreads();
        cc_far_halo<C>(src, arms, count, max(reach[0] - HL, 0), max(reach[1] - HL, 0), tid, gok, g, G, W, w0, rowp, stride, sT, my);
    }
    __syncthreads();
    if (gok) cc_row_phase<C>(dst, rowp, w0, g, G, sT, sP, my, slot, sv);
}

// The same round with the load done by the TMA unit: ONE cp.async.bulk.tensor per CTA brings the 3 rows x NP pixels x 64
// disparities box (24 KB) of Hs_k; cells outside the volume (image borders, granules beyond Dp) arrive as zeros and are never
// used.  Removes the per-thread load loop (12 cp.async with their 64-bit addresses and predicates, a sixth of the kernel's
// instructions); the rows are published by an mbarrier instead of cp.async.wait_all.
// map: float32 tensor {Dp, W, H} of src, box {64, NP, 3}, no swizzle.
template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB) k_cbca_colrow_tma(const __grid_constant__ CUtensorMap map, const float4 *__restrict__ src,
                                                                    float4 *__restrict__ dst, const uchar4 *__restrict__ arms,
                                                                    const int32_t *__restrict__ count, int G, int H, int W) {
    constexpr int S = C::S, NP = C::NP, HL = C::HL;
    extern __shared__ __align__(128) unsigned char cc_raw[];
    __shared__ int reach[2];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned sU = (unsigned)__cvta_generic_to_shared(cc_raw);
    const unsigned sT = sU + NP * 256, sP = sU + 3 * NP * 256;
    const int tid = threadIdx.x, gi = tid & 15, slot = tid >> 4;
    const int g = blockIdx.x * CS_GC + gi, w0 = blockIdx.y * S, h = blockIdx.z;
    if (tid == 0) {
        tc_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        tc_mbar_expect_tx(&bar, 3 * NP * 256);
        tc_tma_load_3d(cc_raw, &map, blockIdx.x * (CS_GC * 4), w0 - HL, h - 1, &bar);
        reach[0] = 0; reach[1] = 0;
    }
    const bool gok = g < G;
    const int sv = min(S, W - w0);
    const int np = min(NP, W - w0 + HL);
    const int p_lo = w0 > 0 ? 0 : HL;
    const ptrdiff_t stride = (ptrdiff_t)W * G;
    const size_t rowp = (size_t)h * W;
    const float4 *cbase = src + (rowp + w0 - HL) * G + g;
    const unsigned my = gi * 16;
    int lneed, rneed;
    const bool far = cc_pixel_info<C>(arms, count, rowp, w0, tid, p_lo, np, sv, sP, lneed, rneed);
    const int any_far = __syncthreads_or(far);                        // publishes the per-pixel information and the mbarrier
    tc_mbar_wait(&bar, 0);

    if (gok) cc_column_phase<C>(cbase, stride, G, sT, sP, my, slot, p_lo, np);
    if (any_far) {
        if (far) { atomicMax(&reach[0], lneed); atomicMax(&reach[1], rneed); }
        __syncthreads();
        cc_far_halo<C>(src, arms, count, max(reach[0] - HL, 0), max(reach[1] - HL, 0), tid, gok, g, G, W, w0, rowp, stride, sT, my);
    }
    __syncthreads();
    if (gok) cc_row_phase<C>(dst, rowp, w0, g, G, sT, sP, my, slot, sv);
}

}  // namespace mccnn
