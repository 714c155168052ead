// EXPERIMENT RECORD (round 2) -- not compiled into libmccnn_b200.so.  The v4 cross-round fused CBCA kernel measured in
// profiles/r2_cbca_chained_experiment.md (bit-identical to the two-pass form, 1.0x algorithmic DRAM traffic, but slower: 0.77-0.82 vs 0.58 ms per round).
// Separable cross-based aggregation, fused ACROSS rounds: the kernels of the default mode of mccnn_cbca.
//
// One round is  out_k(h,w) = ( sum_{h' in spine(h,w)} Hs_k(h',w) ) / |U(h,w)|,  Hs_k(h,w) = sum_{w' in arm(h,w)} out_{k-1}(h,w')
// (pf:640-650, :157-161).  As two streaming passes (cbca_stream.cuh) a round moves 16 B per cell through HBM (Hs_k
// out and back, out_k out and back) and costs ~260 instructions per 4 cells.  k_cbca_colrow runs the column pass of
// round k and the row pass of round k+1 as ONE kernel with out_k in shared memory only:
//     Hs_k -> [out_k] -> Hs_{k+1},   8 B per cell per round;   a call of n rounds is  rows | (n-1) x colrow | cols.
// Fusing this way needs no vertical on-chip state (fusing the two passes of one round needs up to 27 row sums per
// column on chip, which is what the three one-pass kernels of round 1 paid for with their occupancy).
//
// Unit of work: a SEGMENT = S pixels of one image row x 16 disparity granules; the next row pass reaches a
// data-dependent halo left and right of it (usually 0-2 pixels, at most distance_threshold - 1), so out_k is formed
// for segment + halo.  Everything that depends only on the image is computed ONCE per call by k_cbca_plan (the regions
// do not depend on the round or the disparity): per segment the halo [lo, hi) and per pixel a
// descriptor (arms, |U|, RN(1/|U|), position of its runs in a stage).  k_cbca_colrow is persistent (segments dealt
// round robin, so the CTAs sweep the image as one front and the rows above / below a segment are L2 hits) and
// software pipelined with cp.async groups, L = NS - 1 segments deep:
//     iteration k:  descriptors of segment k+2L+1  -> shared (cp.async)
//                   runs of segment k+L: for every pixel its 1 + up + down runs of 256 bytes of Hs_k, in the
//                   reference's summation order, global -> stage (k+L) % NS (cp.async)
//                   wait for segment k's group | column phase: add a pixel's runs in order, divide, out_k -> tile T
//                   barrier | row phase: add out_k along each pixel's horizontal arm from T, store Hs_{k+1}.
// Arithmetic never waits for DRAM, no thread keeps data in flight in registers, and the hot loops are a few
// instructions per run.  Pixels whose runs do not fit the stage (flat image regions: every arm at its limit) are
// gathered straight from global memory, eight loads at a time.
// Same additions in the same order as the two-pass form => bit-identical to it (tests: *_match_two_pass).
#pragma once
#include "cbca_stream.cuh"

namespace mccnn {

// S: segment pixels, NS: stages, CAP: 256-byte runs per stage, PER_SM: resident CTAs per SM
template <int S_, int NS_, int CAP_, int PER_SM_>
struct CcShape {
    static constexpr int S = S_, NS = NS_, CAP = CAP_, PER_SM = PER_SM_;
    static constexpr int L = NS - 1;                    // segments in flight ahead of the arithmetic
    static constexpr int DR = 2 * L + 2;                // descriptor ring slots
    static constexpr int NT = 256, SLOTS = NT / CS_GC;  // threads; pixels one sweep of the CTA covers
    static constexpr int EPOCH = 512;                   // segments per CTA between refills of the halo table
    static constexpr int TPMAX = 96;                    // tile pixels (S + 2 x maximum arm) the plan kernel handles
};

// Plan record of one segment (h, sx), tile pixel t <-> image column sx*S - hm + t, TP = S + 2*hm tile pixels:
//   CcPlanHdr | CcDesc[TP] | int run[CAP]
struct CcDesc {                 // 16 bytes
    unsigned arms;              // up | down << 8 | left << 16 | right << 24
    float n, y;                 // |U| and RN(1 / |U|)
    unsigned pk;                // first run in the stage | (up + down) << 16 | (1u << 31: not staged, gathered from global)
};
struct CcPlanHdr { int lo, hi, runs, pad; };           // active tile pixels [lo, hi); staged runs
// run[q]: float4 offset of run q from (row h, tile pixel 0, granule 0): t * G + dy * W * G, in the order the column
// phase adds them (pixel by pixel: h, h-1, .., h-up, h+1, .., h+down)
static inline size_t cc_record_bytes(int S, int CAP, int hm) { return sizeof(CcPlanHdr) + (size_t)(S + 2 * hm) * sizeof(CcDesc) + (size_t)CAP * 4; }

template <class C>
static inline size_t cc_smem_bytes(int hm) {
    const int TP = C::S + 2 * hm, RM = (TP + C::SLOTS - 1) / C::SLOTS;
    return (size_t)2 * TP * 256 + (size_t)C::NS * C::CAP * 256 + (size_t)C::DR * (RM * C::SLOTS * 16 + C::CAP * 4) + (size_t)2 * TP * 4 +
           (size_t)C::EPOCH * 4;
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per segment (h, sx): halo of the next row pass, runs per pixel, greedy packing of whole pixels into a stage.
template <class C>
__global__ void __launch_bounds__(128) k_cbca_plan(const uchar4 *__restrict__ arms, const int32_t *__restrict__ count,
                                                   unsigned char *__restrict__ plan, int G, int H, int W, int HM) {
    constexpr int S = C::S, J = C::TPMAX / 32;
    __shared__ int runs[4][C::TPMAX], offs[4][C::TPMAX];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nsx = (W + S - 1) / S;
    const int seg = blockIdx.x * 4 + wid;
    if (seg >= H * nsx) return;
    const int h = seg / nsx, sx = seg - h * nsx;
    const int w0 = sx * S, sv = min(S, W - w0), TP = S + 2 * HM;
    uchar4 a[J];
    int c[J];
    int lneed = 0, rneed = 0;
#pragma unroll
    for (int j = 0; j < J; j++) {
        const int t = lane + 32 * j, x = w0 - HM + t;
        const bool in = t < TP && x >= 0 && x < W;
        a[j] = in ? arms[(size_t)h * W + x] : make_uchar4(0, 0, 0, 0);
        c[j] = in ? count[(size_t)h * W + x] : 1;
        const int px = t - HM;
        if (px >= 0 && px < sv) {                       // how far this pixel's row arm leaves the segment
            lneed = max(lneed, (int)a[j].z - px);
            rneed = max(rneed, (int)a[j].w - (sv - 1 - px));
        }
    }
    lneed = __reduce_max_sync(0xffffffffu, lneed);
    rneed = __reduce_max_sync(0xffffffffu, rneed);
    const int lo = HM - lneed, hi = HM + sv + rneed;
#pragma unroll
    for (int j = 0; j < J; j++) {
        const int t = lane + 32 * j;
        if (t < TP) runs[wid][t] = (t >= lo && t < hi) ? 1 + a[j].x + a[j].y : 0;
    }
    __syncwarp();
    int used = 0;
    if (lane == 0) {
        for (int t = lo; t < hi; t++) {
            const int e = runs[wid][t];
            if (used + e <= C::CAP) { offs[wid][t] = used; used += e; }
            else offs[wid][t] = -1;
        }
    }
    __syncwarp();
    unsigned char *rec = plan + (size_t)seg * (sizeof(CcPlanHdr) + (size_t)TP * sizeof(CcDesc) + (size_t)C::CAP * 4);
    if (lane == 0) *reinterpret_cast<CcPlanHdr *>(rec) = CcPlanHdr{lo, hi, used, 0};
    CcDesc *desc = reinterpret_cast<CcDesc *>(rec + sizeof(CcPlanHdr));
    int *run = reinterpret_cast<int *>(desc + TP);
    const int vs = W * G;
#pragma unroll
    for (int j = 0; j < J; j++) {
        const int t = lane + 32 * j;
        if (t < TP) {
            const bool act = t >= lo && t < hi;
            const int off = act ? offs[wid][t] : -1;
            CcDesc d;
            d.arms = (unsigned)a[j].x | (unsigned)a[j].y << 8 | (unsigned)a[j].z << 16 | (unsigned)a[j].w << 24;
            d.n = (float)c[j];
            d.y = 1.0f / d.n;
            d.pk = (off >= 0 ? (unsigned)off : 0x80000000u) | (unsigned)(a[j].x + a[j].y) << 16;
            desc[t] = d;
            if (off >= 0) {
                int *r = run + off;
                r[0] = t * G;
                for (int k = 1; k <= a[j].x; k++) r[k] = t * G - k * vs;
                for (int k = 1; k <= a[j].y; k++) r[a[j].x + k] = t * G + k * vs;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shared memory through 32-bit addresses (no generic-pointer arithmetic in the hot loops)
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
__device__ __forceinline__ unsigned cc_lds32(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void cc_sts128(unsigned a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cc_sts32(unsigned a, const unsigned v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void cc_cp16(unsigned smem, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(g) : "memory");
}

// the quotient of cs_divide with the reciprocal supplied (y = 1.0f / n from the plan)
__device__ __forceinline__ float4 cc_divide(float4 acc, const float n, const float y) {
    const float vmax = fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)));
    const float vmin = fminf(fminf(fabsf(acc.x), fabsf(acc.y)), fminf(fabsf(acc.z), fabsf(acc.w)));
    if (vmax < 1e30f && vmin > 1e-30f)
        return make_float4(cs_div1(acc.x, n, y), cs_div1(acc.y, n, y), cs_div1(acc.z, n, y), cs_div1(acc.w, n, y));
    return make_float4(acc.x / n, acc.y / n, acc.z / n, acc.w / n);
}

// acc += c[k0 * stride], .., c[k1 * stride] in that order, the loads of eight steps issued together
__device__ __noinline__ void cc_walk(float4 &acc, const float4 *__restrict__ c, const int stride, int k0, const int k1) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (; k0 <= k1; k0 += 8) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = k0 + i <= k1 ? c[(ptrdiff_t)(k0 + i) * stride] : z;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (k0 + i <= k1) cs_add(acc, v[i]);
    }
}

// a CTA's k-th segment: seg = blockIdx.x + k * gridDim.x = (h * nsx + sx) * nz + gz, advanced without division;
// d / s: its slot of the descriptor ring / its stage
struct CcCursor {
    int k, gz, sx, h, d, s;
};

template <class C>
__global__ void __launch_bounds__(C::NT, C::PER_SM) k_cbca_colrow(const float4 *__restrict__ src, float4 *__restrict__ dst,
                                                                  const unsigned char *__restrict__ plan, int G, int H, int W, int HM) {
    constexpr int S = C::S, NS = C::NS, CAP = C::CAP, L = C::L, DR = C::DR, SLOTS = C::SLOTS, EPOCH = C::EPOCH;
    extern __shared__ __align__(128) unsigned char cc_raw[];
    const int TP = S + 2 * HM, RM = (TP + SLOTS - 1) / SLOTS;
    const unsigned DSZ = RM * SLOTS * 16 + CAP * 4;                   // one slot of the descriptor ring: descs | run list
    const unsigned sT = (unsigned)__cvta_generic_to_shared(cc_raw);   // [2][TP][256 B]   out_k of segment + halo
    const unsigned sST = sT + 2 * TP * 256;                           // [NS][CAP][256 B] runs of Hs_k
    const unsigned sDR = sST + NS * CAP * 256;                        // [DR][DSZ]
    const unsigned sTA = sDR + DR * DSZ;                              // [2][TP] arms of the tile pixels
    const unsigned sLH = sTA + 2 * TP * 4;                            // [EPOCH] lo | hi << 8 | runs << 16 of this CTA's segments

    const int tid = threadIdx.x, gi = tid % CS_GC, slot = tid / CS_GC;
    const int nz = (G + CS_GC - 1) / CS_GC, nsx = (W + S - 1) / S;
    const int nseg = nz * nsx * H;
    const int nmine = blockIdx.x < nseg ? (nseg - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const size_t rec_bytes = sizeof(CcPlanHdr) + (size_t)TP * sizeof(CcDesc) + (size_t)CAP * 4;
    const int vstride = W * G;
    // cursor step: gridDim.x segments further
    const int dgz = gridDim.x % nz, dsx = (gridDim.x / nz) % nsx, dh = gridDim.x / (nz * nsx);
    auto advance = [&](CcCursor &c) {
        c.k++;
        c.gz += dgz;
        if (c.gz >= nz) { c.gz -= nz; c.sx++; }
        c.sx += dsx;
        if (c.sx >= nsx) { c.sx -= nsx; c.h++; }
        c.h += dh;
        c.d = c.d + 1 == DR ? 0 : c.d + 1;
        c.s = c.s + 1 == NS ? 0 : c.s + 1;
    };

    CcCursor start;
    start.k = 0;
    start.gz = blockIdx.x % nz;
    start.sx = (blockIdx.x / nz) % nsx;
    start.h = blockIdx.x / (nz * nsx);
    start.d = 0;
    start.s = 0;
    int tb = 0;
    const unsigned my_desc = (gi * SLOTS + slot) * 16;      // where this lane parks the descriptor it fetches (pixel round gi)
    const unsigned my16 = gi * 16;

    for (int k0 = 0; k0 < nmine; k0 += EPOCH) {
        const int kend = min(nmine, k0 + EPOCH);
        // halo table of this epoch's segments
        __syncthreads();
        for (int k = k0 + tid; k < kend; k += C::NT) {
            const long long seg = (long long)blockIdx.x + (long long)k * gridDim.x;
            const int hs = (int)(seg / nz);                                           // = h * nsx + sx
            const CcPlanHdr hd = *reinterpret_cast<const CcPlanHdr *>(plan + (size_t)hs * rec_bytes);
            cc_sts32(sLH + (k - k0) * 4, (unsigned)hd.lo | (unsigned)hd.hi << 8 | (unsigned)hd.runs << 16);
        }
        __syncthreads();

        // descriptors + run list of segment c -> its slot of the ring (visible to the whole CTA one barrier later)
        auto fetch_desc = [&](const CcCursor &c) {
            if (c.k < kend) {
                const unsigned lh = cc_lds32(sLH + (c.k - k0) * 4);
                const int lo = lh & 0xff, hi = (lh >> 8) & 0xff, runs = lh >> 16;
                const unsigned char *rec = plan + ((size_t)c.h * nsx + c.sx) * rec_bytes + sizeof(CcPlanHdr);
                const unsigned ring = sDR + c.d * DSZ;
                const int t = lo + slot + SLOTS * gi;
                if (gi < RM && t < hi) cc_cp16(ring + my_desc, rec + (size_t)t * 16);
                if (tid * 4 < runs) cc_cp16(ring + RM * SLOTS * 16 + tid * 16, rec + (size_t)TP * 16 + tid * 16);
            }
        };
        // runs of segment c -> its stage; a thread fetches exactly the 16 bytes of every run it will add itself
        auto issue_runs = [&](const CcCursor &c) {
            if (c.k < kend) {
                const unsigned lh = cc_lds32(sLH + (c.k - k0) * 4);
                const int lo = lh & 0xff, hi = (lh >> 8) & 0xff;
                const int g0 = c.gz * CS_GC;
                if (g0 + gi < G) {
                    const unsigned ring = sDR + c.d * DSZ;
                    const unsigned rl = ring + RM * SLOTS * 16;
                    const unsigned st = sST + c.s * (CAP * 256) + my16;
                    const float4 *sp = src + ((size_t)c.h * W + c.sx * S - HM) * G + g0 + gi;      // tile pixel 0
                    unsigned dsc = ring + slot * 16 + 12;
#pragma unroll 1
                    for (int t = lo + slot; t < hi; t += SLOTS, dsc += SLOTS * 16) {
                        const unsigned pk = cc_lds32(dsc);
                        if ((int)pk < 0) continue;
                        const unsigned off = pk & 0xffff, more = (pk >> 16) & 0xff;
                        unsigned q = st + off * 256, r = rl + off * 4;
                        const unsigned qe = q + more * 256;
#pragma unroll 1
                        for (; q <= qe; q += 256, r += 4) cc_cp16(q, sp + (int)cc_lds32(r));
                    }
                }
            }
        };

        CcCursor cp_ = start, ci = start, cc = start;
        // prologue: descriptors of the first 2L+1 segments, then the runs of the first L
        for (int p = 0; p <= 2 * L; p++) { fetch_desc(cp_); advance(cp_); }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();
        for (int p = 0; p < L; p++) {
            issue_runs(ci);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            advance(ci);
        }
#pragma unroll 1
        for (; cc.k < kend; advance(cc), tb ^= 1) {
            fetch_desc(cp_);
            issue_runs(ci);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            advance(cp_);
            advance(ci);
            asm volatile("cp.async.wait_group %0;\n" ::"n"(L) : "memory");

            const unsigned lh = cc_lds32(sLH + (cc.k - k0) * 4);
            const int lo = lh & 0xff, hi = (lh >> 8) & 0xff;
            const int w0 = cc.sx * S, sv = min(S, W - w0), g0 = cc.gz * CS_GC;
            const unsigned T = sT + tb * (TP * 256) + my16, TA = sTA + tb * (TP * 4);
            {
                // column phase: out_k = (sum of the pixel's runs, in order) / |U|
                const unsigned st = sST + cc.s * (CAP * 256) + my16;
                unsigned dsc = sDR + cc.d * DSZ + slot * 16;
#pragma unroll 1
                for (int t = lo + slot; t < hi; t += SLOTS, dsc += SLOTS * 16) {
                    const uint4 d = cc_lds128u(dsc);
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((int)d.w >= 0) {
                        unsigned q = st + (d.w & 0xffff) * 256;
                        const unsigned qe = q + ((d.w >> 16) & 0xff) * 256;
                        cs_add(acc, cc_lds128(q));
#pragma unroll 1
                        for (q += 256; q <= qe; q += 256) cs_add(acc, cc_lds128(q));
                    } else if (g0 + gi < G) {
                        const float4 *cp = src + ((size_t)cc.h * W + w0 - HM + t) * G + g0 + gi;
                        cs_add(acc, cp[0]);
                        cc_walk(acc, cp, -vstride, 1, d.x & 0xff);
                        cc_walk(acc, cp, vstride, 1, (d.x >> 8) & 0xff);
                    }
                    cc_sts128(T + t * 256, cc_divide(acc, __uint_as_float(d.y), __uint_as_float(d.z)));
                    if (gi == 0) cc_sts32(TA + t * 4, d.x);
                }
            }
            __syncthreads();
            // row phase: Hs_{k+1} = sum of out_k along the horizontal arm, from shared memory
            if (g0 + gi < G) {
                float4 *out = dst + ((size_t)cc.h * W + w0) * G + g0 + gi;
#pragma unroll 1
                for (int px = slot; px < sv; px += SLOTS) {
                    const unsigned a = cc_lds32(TA + (HM + px) * 4);
                    const unsigned q0 = T + (HM + px) * 256;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    cs_add(acc, cc_lds128(q0));                                                     // w, w-1, .., w-left, w+1, .., w+right
                    const unsigned ql = q0 - ((a >> 16) & 0xff) * 256, qr = q0 + (a >> 24) * 256;
#pragma unroll 1
                    for (unsigned q = q0 - 256; q >= ql && q < q0; q -= 256) cs_add(acc, cc_lds128(q));
#pragma unroll 1
                    for (unsigned q = q0 + 256; q <= qr; q += 256) cs_add(acc, cc_lds128(q));
                    out[(size_t)px * G] = acc;
                }
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        start = cc;
    }
}

typedef CcShape<30, 3, 80, 2> CcDeep;      // 2 CTAs per SM, two segments in flight each (98 KB at arm limit 13)
typedef CcShape<30, 2, 72, 3> CcShallow;   // 3 CTAs per SM, one segment in flight each (71 KB)

// can the chained kernels run this problem?  (tile within the plan kernel's lanes, a pixel's runs within a stage,
// halo bounds in a byte, 32-bit run offsets)
template <class C>
static inline bool cc_supports(int G, int H, int W, int hm) {
    const long long nseg = (long long)((G + CS_GC - 1) / CS_GC) * ((W + C::S - 1) / C::S) * H;
    return C::S + 2 * hm <= C::TPMAX && 1 + 2 * hm <= C::CAP && C::CAP <= 255 && (long long)H * W * G < (1ll << 31) && nseg < (1ll << 31);
}

}  // namespace mccnn
