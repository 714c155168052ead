// Separable cross-based aggregation round as ONE row-marching kernel (mode MCCNN_CBCA_SEPARABLE_MARCH).
//
// Same sums in the same order as the two streaming passes of cbca_stream.cuh (pf:640-650):
//   out(h,w) = ( sum_{h' in spine(h,w)} Hs(h',w) ) / |U(h,w)|,   Hs(h',w) = sum_{w' in arm(h',w)} in(h',w'),
// but every cell is read from HBM once and written once (8 B per cell per round, the stage's minimum).
//
// A CTA owns a strip of WT pixel columns x 8 disparity granules (128 contiguous bytes per pixel) and marches
// down the rows of a row segment.  One consumer thread = one (column, granule) for the whole march:
//   * input rows arrive in a shared-memory stage ring through TMA box copies (cp.async.bulk.tensor) issued by
//     a producer warp NST rows ahead: full/empty mbarriers, no __syncthreads anywhere in the march;
//   * the thread forms Hs(r, w) by walking the row arm inside the staged row (the only cross-thread data, and it
//     is written by the TMA unit), and keeps the last 27 row sums of ITS column in a thread-private
//     shared-memory ring -- the spine walk of output row r-13 reads only that ring, so the vertical pass needs
//     no communication and no barrier at all;
//   * the producer warp also stages the arms and |U| of each row (4 + 8 bytes per pixel) and decides per row how
//     much halo the strip's arms actually reach into: none, 2 or 13 pixels per side, each a separate box copy
//     (natural images mostly need none; halo re-reads are served by L2 since neighbouring strips march in step).
// Lanes of a warp are the 8 granules of 4 neighbouring columns, so an arm walk diverges over 4 pixels only and
// every shared-memory access of a warp is one contiguous 512-byte run (conflict free for any walk offset).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "cbca_tile.cuh"
#include "cbca_stream.cuh"

namespace mccnn {

constexpr int CM_GT = 8;                        // granules per strip
constexpr int CM_ARM = 13;                      // longest arm: distance_threshold <= 14
constexpr int CM_RING = 2 * CM_ARM + 1;         // row sums kept per column
constexpr int CM_META = 24;                     // rows of arms / counts kept (>= CM_ARM + NST + 2)
constexpr int CM_HSMALL = 2;                    // the small halo box
constexpr int CM_PD = 8;                        // rows of arms / counts the producer prefetches into registers

struct CmMaps { CUtensorMap centre, small, full; };   // boxes [32 floats][WT | 2 | 13 pixels][1 row]

template <int WT, int NST>
struct __align__(128) CmSmem {
    float4 stage[NST][(WT + 2 * CM_ARM) * CM_GT];     // [halo 13 | strip | halo 13] pixels x 8 granules
    float4 ring[CM_RING][WT * CM_GT];
    float2 cnt[CM_META][WT];                          // (|U|, RN(1/|U|))
    uchar4 arms[CM_META][WT];
    unsigned long long full[NST], empty[NST];
};

// Shared-memory accesses by 32-bit shared address (the march keeps byte addresses, not pointers).  A granule is held
// as two packed f32x2 registers so that the adds are the packed FADD2 (each half rounded like a plain add) and can be
// issued under a predicate without a branch.
struct cm_p4 { unsigned long long lo, hi; };
__device__ __forceinline__ cm_p4 cm_lds(unsigned addr) {
    cm_p4 v;
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];\n" : "=l"(v.lo), "=l"(v.hi) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cm_sts(unsigned addr, const cm_p4 v) {
    asm volatile("st.shared.v2.b64 [%0], {%1,%2};\n" ::"r"(addr), "l"(v.lo), "l"(v.hi) : "memory");
}
__device__ __forceinline__ unsigned cm_lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 cm_lds64(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cm_acc(cm_p4 &acc, const cm_p4 v) {
    asm("add.rn.f32x2 %0, %0, %2;\n\tadd.rn.f32x2 %1, %1, %3;\n" : "+l"(acc.lo), "+l"(acc.hi) : "l"(v.lo), "l"(v.hi));
}
// acc += v if n >= K, as predicated instructions (no branch)
template <int K>
__device__ __forceinline__ void cm_acc_ge(cm_p4 &acc, const cm_p4 v, unsigned n) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %4, %5;\n\t@p add.rn.f32x2 %0, %0, %2;\n\t@p add.rn.f32x2 %1, %1, %3;\n\t}\n"
        : "+l"(acc.lo), "+l"(acc.hi)
        : "l"(v.lo), "l"(v.hi), "r"(n), "n"(K));
}
__device__ __forceinline__ unsigned long long cm_pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};\n" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float2 cm_unpack(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0,%1}, %2;\n" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// a / n for n = |U| (an integer <= 729), y = RN(1/n): q = RN(a*y); r = a - n*q (exact, FMA); RN(q + r*y), both halves
__device__ __forceinline__ unsigned long long cm_div2(unsigned long long a, unsigned long long nn, unsigned long long yy) {
    unsigned long long q, r;
    asm("mul.rn.f32x2 %0, %2, %3;\n\tfma.rn.f32x2 %1, %4, %0, %2;\n\tfma.rn.f32x2 %0, %1, %3, %0;\n"
        : "=&l"(q), "=&l"(r)
        : "l"(a), "l"(yy), "l"(nn));
    return q;
}
__device__ __forceinline__ void cm_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cm_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CM_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CM_DONE_%=;\n"
        "bra CM_WAIT_%=;\n"
        "CM_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

template <int WT, int NST, int MINB>
__global__ void __launch_bounds__(WT * CM_GT + 32, MINB)
k_cbca_march(const __grid_constant__ CmMaps maps, float4 *__restrict__ out, const uchar4 *__restrict__ arms,
             const int32_t *__restrict__ count, int G, int H, int W, int nW, int nG, int hseg) {
    extern __shared__ __align__(128) unsigned char cm_raw[];
    CmSmem<WT, NST> &sm = *reinterpret_cast<CmSmem<WT, NST> *>(cm_raw);
    constexpr int NC = WT * CM_GT;                    // consumer threads
    const int tid = threadIdx.x;
    int u = blockIdx.x;
    const int sw = u % nW;
    u /= nW;
    const int sg = u % nG, seg = u / nG;
    const int w0 = sw * WT, g0 = sg * CM_GT;
    const int hs0 = seg * hseg, hs1 = min(H, hs0 + hseg);           // output rows of this CTA
    const int rbeg = max(0, hs0 - CM_ARM), rend = hs1 + CM_ARM;     // march; input rows are [rbeg, rin)
    const int rin = min(H, rend);
    if (tid == 0) {
        for (int i = 0; i < NST; i++) {
            ct_mbar_init(&sm.full[i], 1);
            ct_mbar_init(&sm.empty[i], NC / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (tid >= NC) {
        // ---------------------------------------------------------------- producer warp
        const int lane = tid - NC;
        const int w = w0 + lane;
        const bool pv = lane < WT && w < W;
        // arms / |U| of the rows ahead are fetched a batch of CM_PD rows at a time, one batch ahead: the loads of a
        // batch are consumed together at the start of the next one, so the producer waits for global memory once per
        // CM_PD rows (loads waited for one by one share scoreboards with the newest load: a full memory latency per
        // row, which was the whole kernel's pace)
        uchar4 a_nxt[CM_PD], a_cur[CM_PD];
        int c_nxt[CM_PD], c_cur[CM_PD];
#pragma unroll
        for (int i = 0; i < CM_PD; i++) {
            a_nxt[i] = make_uchar4(0, 0, 0, 0);
            c_nxt[i] = 1;
            if (pv && rbeg + i < rin) { a_nxt[i] = arms[(size_t)(rbeg + i) * W + w]; c_nxt[i] = count[(size_t)(rbeg + i) * W + w]; }
        }
        int rm = rbeg % CM_META, slot = 0, it = 0;
        unsigned phase = 1;                                                 // parity of the consumers' previous release
        for (int rb = rbeg; rb < rin; rb += CM_PD) {
#pragma unroll
            for (int i = 0; i < CM_PD; i++) { a_cur[i] = a_nxt[i]; c_cur[i] = c_nxt[i]; }
#pragma unroll
            for (int i = 0; i < CM_PD; i++)
                if (pv && rb + CM_PD + i < rin) {
                    a_nxt[i] = arms[(size_t)(rb + CM_PD + i) * W + w];
                    c_nxt[i] = count[(size_t)(rb + CM_PD + i) * W + w];
                }
#pragma unroll
            for (int i = 0; i < CM_PD; i++) {
                const int r = rb + i;
                if (r >= rin) break;
                const uchar4 a = a_cur[i];
                const int c = c_cur[i];
                // everything that does not need the stage slot happens before the wait for it: the row's arms and |U|
                // go to the meta ring (its slot was released long ago: CM_META > NST + CM_ARM + 2) and the halo classes
                // are decided, so that the box copies leave right after the consumers' release
                if (lane < WT) {
                    const float n = (float)c;
                    sm.arms[rm][lane] = a;
                    sm.cnt[rm][lane] = make_float2(n, 1.0f / n);
                }
                int nl = lane < WT ? (int)a.z - lane : 0;                   // pixels the left arms reach past the strip
                int nr = lane < WT ? (int)a.w - (WT - 1 - lane) : 0;
                nl = __reduce_max_sync(0xffffffffu, nl);
                nr = __reduce_max_sync(0xffffffffu, nr);
                const int hl = nl <= 0 ? 0 : (nl <= CM_HSMALL ? CM_HSMALL : CM_ARM);
                const int hr = nr <= 0 ? 0 : (nr <= CM_HSMALL ? CM_HSMALL : CM_ARM);
                __syncwarp();
                if (it >= NST) cm_wait(ct_smem_u32(&sm.empty[slot]), phase);
                if (lane == 0) {
                    float4 *dst = sm.stage[slot];
                    ct_mbar_expect_tx(&sm.full[slot], (unsigned)((WT + hl + hr) * CM_GT * 16));
                    ct_tma_load_3d(dst + CM_ARM * CM_GT, &maps.centre, 4 * g0, w0, r, &sm.full[slot]);
                    if (hl == CM_HSMALL) ct_tma_load_3d(dst + (CM_ARM - CM_HSMALL) * CM_GT, &maps.small, 4 * g0, w0 - CM_HSMALL, r, &sm.full[slot]);
                    else if (hl) ct_tma_load_3d(dst, &maps.full, 4 * g0, w0 - CM_ARM, r, &sm.full[slot]);
                    if (hr == CM_HSMALL) ct_tma_load_3d(dst + (CM_ARM + WT) * CM_GT, &maps.small, 4 * g0, w0 + WT, r, &sm.full[slot]);
                    else if (hr) ct_tma_load_3d(dst + (CM_ARM + WT) * CM_GT, &maps.full, 4 * g0, w0 + WT, r, &sm.full[slot]);
                }
                if (++rm == CM_META) rm = 0;
                if (++slot == NST) { slot = 0; phase ^= 1; }
                it++;
            }
        }
        return;
    }

    // -------------------------------------------------------------------- consumers
    // Everything that changes per row is a loop-carried shared-memory byte address: no multiplies, divisions or
    // 64-bit index arithmetic inside the march (the kernel is bound by issue slots, not by bytes).
    const int wi = tid / CM_GT, gi = tid % CM_GT, lane = tid & 31;
    const int w = w0 + wi, g = g0 + gi;
    const bool st_ok = w < W && g < G;
    constexpr unsigned RS = NC * 16, RING_BYTES = CM_RING * RS;               // one ring slot / the whole ring
    constexpr unsigned SS = (WT + 2 * CM_ARM) * CM_GT * 16;                    // one stage
    constexpr unsigned PX = CM_GT * 16;                                        // one staged pixel
    const unsigned ring0 = ct_smem_u32(&sm.ring[0][tid]);
    const unsigned stage0 = ct_smem_u32(&sm.stage[0][(CM_ARM + wi) * CM_GT + gi]);
    const unsigned arms0 = ct_smem_u32(&sm.arms[0][wi]), cnt0 = ct_smem_u32(&sm.cnt[0][wi]);
    const unsigned full0 = ct_smem_u32(&sm.full[0]), empty0 = ct_smem_u32(&sm.empty[0]);
    float4 *outp = out + ((size_t)hs0 * W + w) * G + g;
    const size_t rowstride = (size_t)W * G;
    unsigned rr = (unsigned)(rbeg % CM_RING) * RS;                             // ring slot of row r
    int rm = rbeg % CM_META;                                                   // meta slot of row r
    int slot = 0;
    unsigned phase = 0;
    // 96 % of the arms are <= 2 long (80 % are 0), and the march is bound by the length of a step's dependent chain,
    // not by bytes or issue slots.  So the two nearest cells on either side are loaded unconditionally, all at once,
    // and added under predicates in the reference's order; longer arms continue in a (rare) loop.  The ring loads
    // of the output row do not depend on this step's row and are issued before the wait on the stage.
    // 96 % of the arms are <= 2 long (80 % are 0) and the march is bound by the length of a step's dependent chain:
    // the two nearest cells on either side are loaded unconditionally, all at once, and added under predicates in
    // the reference's order (no branches); only longer arms continue in a loop.  The ring loads of the output row
    // do not depend on this step's row and are issued before the wait on the stage.
    for (int r = rbeg; r < rend; r++) {
        const bool has_out = r - CM_ARM >= hs0;
        unsigned ro = rr + (CM_RING - CM_ARM) * RS;                            // ring slot of row r - 13
        if (ro >= RING_BYTES) ro -= RING_BYTES;
        unsigned ao = 0;
        float2 nc = make_float2(1.f, 1.f);
        cm_p4 o0, u1, u2, d1, d2;
        if (has_out) {
            int om = rm - CM_ARM;
            if (om < 0) om += CM_META;
            ao = cm_lds32(arms0 + om * (WT * 4));
            nc = cm_lds64(cnt0 + om * (WT * 8));
            const unsigned pu1 = ro == 0 ? RING_BYTES - RS : ro - RS, pu2 = pu1 == 0 ? RING_BYTES - RS : pu1 - RS;
            const unsigned pd1 = ro == RING_BYTES - RS ? 0 : ro + RS, pd2 = pd1 == RING_BYTES - RS ? 0 : pd1 + RS;
            o0 = cm_lds(ring0 + ro);
            u1 = cm_lds(ring0 + pu1);
            u2 = cm_lds(ring0 + pu2);
            d1 = cm_lds(ring0 + pd1);
            d2 = cm_lds(ring0 + pd2);
        }
        if (r < rin) {
            cm_wait(full0 + slot * 8, phase);
            const unsigned c = stage0 + slot * SS;
            const unsigned a = cm_lds32(arms0 + rm * (WT * 4));
            cm_p4 acc = cm_lds(c);                                             // w, w-1, .., w-left (pf:645-650)
            const cm_p4 l1 = cm_lds(c - PX), l2 = cm_lds(c - 2 * PX), r1 = cm_lds(c + PX), r2 = cm_lds(c + 2 * PX);
            const unsigned nl = (a >> 16) & 0xff, nr = a >> 24;
            cm_acc_ge<1>(acc, l1, nl);
            cm_acc_ge<2>(acc, l2, nl);
            if (nl > 2) {
#pragma unroll 1
                for (unsigned p = c - 2 * PX, e = c - nl * PX; p != e;) { p -= PX; cm_acc(acc, cm_lds(p)); }
            }
            cm_acc_ge<1>(acc, r1, nr);                                         // w+1, .., w+right
            cm_acc_ge<2>(acc, r2, nr);
            if (nr > 2) {
#pragma unroll 1
                for (unsigned p = c + 2 * PX, e = c + nr * PX; p != e;) { p += PX; cm_acc(acc, cm_lds(p)); }
            }
            cm_sts(ring0 + rr, acc);
            __syncwarp();
            if (lane == 0) cm_arrive(empty0 + slot * 8);
            if (++slot == NST) { slot = 0; phase ^= 1; }
        }
        if (has_out) {
            const unsigned nu = ao & 0xff, nd = (ao >> 8) & 0xff;
            cm_p4 acc = o0;                                                    // h, h-1, .., h-up (pf:640-644)
            cm_acc_ge<1>(acc, u1, nu);
            cm_acc_ge<2>(acc, u2, nu);
            if (nu > 2) {
                unsigned p = ro >= 2 * RS ? ro - 2 * RS : ro + RING_BYTES - 2 * RS;
#pragma unroll 1
                for (unsigned k = nu - 2; k != 0; k--) {
                    p = p == 0 ? RING_BYTES - RS : p - RS;
                    cm_acc(acc, cm_lds(ring0 + p));
                }
            }
            cm_acc_ge<1>(acc, d1, nd);                                         // h+1, .., h+down
            cm_acc_ge<2>(acc, d2, nd);
            if (nd > 2) {
                unsigned p = ro + 2 * RS;
                if (p >= RING_BYTES) p -= RING_BYTES;
#pragma unroll 1
                for (unsigned k = nd - 2; k != 0; k--) {
                    p = p == RING_BYTES - RS ? 0 : p + RS;
                    cm_acc(acc, cm_lds(ring0 + p));
                }
            }
            const float n = nc.x;
            const float2 s0 = cm_unpack(acc.lo), s1 = cm_unpack(acc.hi);
            const float hi = fmaxf(fmaxf(fabsf(s0.x), fabsf(s0.y)), fmaxf(fabsf(s1.x), fabsf(s1.y)));
            const float lo = fminf(fminf(fabsf(s0.x), fabsf(s0.y)), fminf(fabsf(s1.x), fabsf(s1.y)));
            float4 q;
            if (hi < 1e30f && lo > 1e-30f) {
                const unsigned long long nn = cm_pack(-n, -n), yy = cm_pack(nc.y, nc.y);
                const float2 q0 = cm_unpack(cm_div2(acc.lo, nn, yy)), q1 = cm_unpack(cm_div2(acc.hi, nn, yy));
                q = make_float4(q0.x, q0.y, q1.x, q1.y);
            } else {
                q = make_float4(s0.x / n, s0.y / n, s1.x / n, s1.y / n);       // pf:161
            }
            if (st_ok) *outp = q;
            outp += rowstride;
        }
        rr = rr == RING_BYTES - RS ? 0 : rr + RS;
        if (++rm == CM_META) rm = 0;
    }
}

}  // namespace mccnn
