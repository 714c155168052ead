// Matching-cost volume (pf:78-113):  L[d,h,w] = -<fl[h,w,:], fr[h,w-d,:]>  (w >= d),
// R[d,h,w] = L[d,h,w+d]  (w < W-d), the invalid triangles filled by the 3-tap mean recurrence
// (pf:94-95, :105-106).  Volumes are written in the HWD layout.
//
// The stage is 8 + 512/D bytes and 128 flop per cell: at the float32-SIMT ridge of B200, so the dot
// products run on the tensor cores and the kernel is left with what it should be bound by -- writing
// the two volumes.  For one image row the scores are a banded slice of the GEMM
// S = FL[h] (W x 64) . FR[h]^T (64 x W): only x = w - d with 0 <= d < ndisp is needed.
//
// k_cost_volume_tc: persistent, one CTA per SM, 512 threads in two groups.
//   front end (warps 0-7): TMA (cp.async.bulk.tensor, 128B swizzle) brings the 128-pixel left tile and, 64 right pixels
//       at a time, the right rows it can match; the float32 operands are split in shared memory into hi = tf32(x) and
//       lo = x - hi (exact), and one thread issues tcgen05.mma kind::tf32 three times per K step -- hi.hi + hi.lo +
//       lo.hi, accumulated in float32 in TMEM -- which restores float32-level accuracy (error ~1e-6 of the volume's
//       scale; plain TF32 would break the 1e-4 gate).  S (lanes = left pixels w, columns = right
//       pixels x) sits in one of four TMEM accumulators, so the next chunks are loaded, split and multiplied while the
//       previous ones are written out; the negation (pf:111-112) is the descriptor's negate-A bit.
//   epilogue (warps 8-15, two per TMEM lane quarter: one writes R, one writes L, from the same accumulator).
//       Both volumes are [pixel][d], d = w - x, so a tile's cells are parallelograms: whatever the mapping, a pixel's
//       disparities come out of a (tile, chunk) pair as a contiguous run that starts at an arbitrary float.  B200's L2
//       takes stores at full speed only in whole 32-byte sectors (scripts/microbench/span_write.cu: 6.3 TB/s against
//       2.8-3.5 TB/s for 128-byte runs that start mid-sector, which is what round 1's straight-from-registers epilogue
//       issued), so both volumes leave through shared memory:
//       L: a lane IS a left pixel, and all of its disparities are produced by this warp in this tile, 32 per column
//          group, descending.  The lane scatters them into its own 64-float ring (bank = d mod 32: conflict free); after
//          every group exactly one aligned 32-float line per pixel is complete and the warp writes those 32 lines as
//          whole 128-byte lines (8 lanes x float4 per line).  Every L sector is written once, whole.
//       R: the 128 lanes of the four R warps are 128 consecutive disparities of each of the group's 32 right pixels:
//          the warps transpose the group through a shared stage ([pixel][d], conflict free), and each pixel's run
//          (512 bytes, clipped to the band) is copied out as float4 pieces from the first 16-byte boundary on, with one
//          predicated scalar instruction for the ends.  Only the two sectors at the ends of a run are partial (they
//          are completed by the neighbouring tile).
// k_cost_fill then overwrites the cells that have no correspondent.
#include "tc_common.cuh"

namespace mccnn {

constexpr int CV_C = 64;                   // feature channels (model.py:38)
constexpr int CV_BM = 128;                 // left pixels per tile
constexpr int CV_BN = 64;                  // right pixels per chunk
constexpr int CV_KA_BYTES = CV_BM * 128;   // one K block of the left tile: 128 rows x 32 floats, 128B-swizzled
constexpr int CV_KB_BYTES = CV_BN * 128;   // one K block of a right chunk
constexpr int CV_A_BYTES = 2 * CV_KA_BYTES, CV_B_BYTES = 2 * CV_KB_BYTES;
constexpr int CV_THREADS = 512;            // warp 0: MMA issue, warp 1: TMA, warps 2-7: operand split, 8-11: R, 12-15: L
constexpr int CV_NSPLIT = 192;             // splitter threads (warps 2-7)
constexpr int CV_NACC = 4;                 // TMEM accumulators (64 columns each)
constexpr int CV_TMEM_COLS = CV_NACC * CV_BN;
constexpr int CV_RING = 64;                // floats per pixel in an L warp's ring (two lines)
constexpr int CV_RPITCH = 132;             // floats per pixel row of the R stage: 128 + up to 3 of alignment shift

struct __align__(1024) CvSmem {
    unsigned char a_hi[CV_A_BYTES], a_lo[CV_A_BYTES];                  // left tile, split
    unsigned char a_raw[CV_A_BYTES];                                   // next left tile as loaded (prefetch)
    unsigned char b_hi[2][CV_B_BYTES], b_lo[2][CV_B_BYTES];            // right chunks, split, double buffered
    float l_ring[4][32 * CV_RING];                                     // per L warp: [pixel (lane)][d mod 64]
    float r_stage[2][32 * CV_RPITCH];                                  // R warps: [right pixel of the group][d - d0 + shift]
    unsigned long long bar_tma_a, bar_tma_b[2], bar_full[CV_NACC], bar_empty[CV_NACC];
    unsigned tmem_base;
};
static_assert(sizeof(CvSmem) <= 232448, "shared memory of k_cost_volume_tc");

// The clipping of an R run depends only on (column group of the tile, right pixel of the group): its first disparity is
// d0 = b0 - i with b0 = CV_BN * nchunks - CV_BM - 32 * group.  Made once on the host, read through the constant bank:
//   bits 0-7  column in the stage row of the first float4 piece     bits 8-13  float4 pieces (0 .. 32)
//   bits 14-23  disparity of the first piece                        bits 24-25 / 26-27  floats before / after the pieces
constexpr int CV_MAX_GROUPS = 2 * ((CV_BM - 1 + 512 + CV_BN - 1) / CV_BN);
struct CvTable { unsigned r[CV_MAX_GROUPS][32]; };
static unsigned cv_table_entry(int b0, int i, int D) {
    const int d0 = b0 - i, dlo = d0 > 0 ? d0 : 0, dhi = d0 + 128 < D ? d0 + 128 : D;
    const int len = dhi > dlo ? dhi - dlo : 0;
    int head = (4 - (dlo & 3)) & 3;
    if (head > len) head = len;
    const int nb = (len - head) >> 2, tail = len - head - 4 * nb, gd = dlo + head, col = (d0 & 3) + gd - d0;
    return (unsigned)col | (unsigned)nb << 8 | (unsigned)gd << 14 | (unsigned)head << 24 | (unsigned)tail << 26;
}

struct CvMaps { CUtensorMap fl, fr; };      // [H][W][64] float32, box {32 channels, 128 | 64 pixels, 1 row}, SWIZZLE_128B

__device__ __forceinline__ void cv_sts(unsigned a, unsigned v) { asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float cv_lds(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 cv_lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

__global__ void __launch_bounds__(CV_THREADS, 1)
k_cost_volume_tc(const __grid_constant__ CvMaps maps, const __grid_constant__ CvTable tab, float *__restrict__ L,
                 float *__restrict__ R, int H, int W, int D, int Dp, int nwt, int nchunks, int ntiles, int dbase) {
    // Disparity slab [dbase, dbase + D): the right pixel matched at local disparity d is x = w - dbase - d, so the
    // right chunks are fetched (and the R cells stored) dbase pixels to the left of where the local band sits;
    // chunks left of the image are zero-filled by the TMA unit and land in the triangle k_cost_fill overwrites.
    extern __shared__ __align__(1024) unsigned char cv_raw[];
    CvSmem &sm = *reinterpret_cast<CvSmem *>(cv_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(&sm.tmem_base)),
                     "r"(CV_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        tc_mbar_init(&sm.bar_tma_a, 1);
        tc_mbar_init(&sm.bar_tma_b[0], 1);
        tc_mbar_init(&sm.bar_tma_b[1], 1);
        for (int i = 0; i < CV_NACC; i++) {
            tc_mbar_init(&sm.bar_full[i], 1);
            tc_mbar_init(&sm.bar_empty[i], 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = sm.tmem_base;
    // instruction descriptor: D = F32, A = B = TF32, A negated (pf:111-112), both K-major, N = 64, M = 128
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((unsigned)(CV_BN >> 3) << 17) |
                           ((unsigned)(CV_BM >> 4) << 24);
    // chunk g of this CTA uses right-chunk stage g & 1 and accumulator g % CV_NACC; bar_full[g % CV_NACC] completes its
    // phase (g / CV_NACC) & 1 when the MMAs of chunk g are done (they have then also finished reading the stage)

    if (warp == 1) {
        // ================= TMA producer (one lane) =================
        if (lane == 0) {
            auto issue_b = [&](int tile, int c, unsigned g) {
                const int h = tile / nwt, w0 = (tile - h * nwt) * CV_BM;
                const int x0c = w0 + CV_BM - CV_BN * nchunks + CV_BN * c;
                unsigned char *dst = sm.b_hi[g & 1];
                tc_mbar_expect_tx(&sm.bar_tma_b[g & 1], CV_B_BYTES);
                tc_tma_load_3d(dst, &maps.fr, 0, x0c - dbase, h, &sm.bar_tma_b[g & 1]);
                tc_tma_load_3d(dst + CV_KB_BYTES, &maps.fr, 32, x0c - dbase, h, &sm.bar_tma_b[g & 1]);
            };
            auto issue_a = [&](int tile) {
                const int h = tile / nwt, w0 = (tile - h * nwt) * CV_BM;
                tc_mbar_expect_tx(&sm.bar_tma_a, CV_A_BYTES);
                tc_tma_load_3d(sm.a_raw, &maps.fl, 0, w0, h, &sm.bar_tma_a);
                tc_tma_load_3d(sm.a_raw + CV_KA_BYTES, &maps.fl, 32, w0, h, &sm.bar_tma_a);
            };
            unsigned g = 0;
            if ((int)blockIdx.x < ntiles) {
                issue_a(blockIdx.x);
                issue_b(blockIdx.x, 0, 0);
            }
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int c = 0; c < nchunks; c++, g++) {
                    // chunk g+1 lands in the stage chunk g-1 used: wait until the MMAs of g-1 are done with it.
                    if (g > 0) {
                        tc_mbar_wait_sleep(&sm.bar_full[(g - 1) % CV_NACC], ((g - 1) / CV_NACC) & 1);
                        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                        // chunk g-1 opened a tile <=> c == 1 (or nchunks == 1): its split has consumed a_raw
                        const bool opened = (nchunks == 1) ? true : (c == 1);
                        const int opened_tile = (nchunks == 1) ? tile - (int)gridDim.x : tile;
                        if (opened && opened_tile + (int)gridDim.x < ntiles) issue_a(opened_tile + gridDim.x);
                    }
                    if (c + 1 < nchunks) issue_b(tile, c + 1, g + 1);
                    else if (tile + (int)gridDim.x < ntiles) issue_b(tile + gridDim.x, 0, g + 1);
                }
            }
        }
    } else if (warp < 8) {
        // ================= operand split (warps 2-7) and MMA issue (warp 0, lane 0) =================
        unsigned g = 0, ta = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ta++) {
            for (int c = 0; c < nchunks; c++, g++) {
                const unsigned stage = g & 1, acc_i = g % CV_NACC;
                if (warp >= 2) {
                    if (c == 0) {
                        // the previous tile's MMAs no longer read a_hi / a_lo; the new tile was prefetched into a_raw
                        if (g > 0) tc_mbar_wait(&sm.bar_full[(g - 1) % CV_NACC], ((g - 1) / CV_NACC) & 1);
                        tc_mbar_wait(&sm.bar_tma_a, ta & 1);
                        tc_split(sm.a_raw, sm.a_hi, sm.a_lo, CV_A_BYTES, tid - 64, CV_NSPLIT);
                    }
                    tc_mbar_wait(&sm.bar_tma_b[stage], (g >> 1) & 1);
                    tc_split(sm.b_hi[stage], sm.b_hi[stage], sm.b_lo[stage], CV_B_BYTES, tid - 64, CV_NSPLIT);
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // operand accesses -> async proxy
                }
                tc_named_barrier(1, 32 + CV_NSPLIT);
                if (tid == 0) {
                    if (g >= CV_NACC) tc_mbar_wait(&sm.bar_empty[acc_i], ((g / CV_NACC) - 1) & 1);   // drained by the epilogue
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const unsigned d_s = tmem_base + acc_i * CV_BN;
                    const unsigned long long ah = tc_smem_desc(tc_smem_u32(sm.a_hi)), al = tc_smem_desc(tc_smem_u32(sm.a_lo));
                    const unsigned long long bh = tc_smem_desc(tc_smem_u32(sm.b_hi[stage])), bl = tc_smem_desc(tc_smem_u32(sm.b_lo[stage]));
                    unsigned acc = 0;
#pragma unroll
                    for (int kb = 0; kb < 2; kb++)
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const unsigned long long oa = (unsigned long long)((kb * CV_KA_BYTES + ks * 32) >> 4);
                            const unsigned long long ob = (unsigned long long)((kb * CV_KB_BYTES + ks * 32) >> 4);
                            // S = -FL.FR^T as hi.hi + hi.lo + lo.hi
                            tc_mma_tf32(d_s, ah + oa, bh + ob, idesc, acc);
                            acc = 1;
                            tc_mma_tf32(d_s, ah + oa, bl + ob, idesc, 1);
                            tc_mma_tf32(d_s, al + oa, bh + ob, idesc, 1);
                        }
                    tc_mma_commit(&sm.bar_full[acc_i]);
                }
            }
        }
    } else {
        // ================= epilogue: TMEM -> shared memory -> R and L =================
        const int q = warp & 3;                       // TMEM lane quarter this warp may read: left pixels w0 + 32q + lane
        const bool does_l = warp >= 12;               // warps 8-11 stage R, warps 12-15 stage and write L; all eight write R
        const int ew = warp - 8;
        const unsigned ring = tc_smem_u32(sm.l_ring[q]) + lane * (CV_RING * 4);      // this lane's (= left pixel's) ring
        const int sub = lane >> 3, l8 = lane & 7;     // flush of L: 4 pixels per instruction, 8 float4 pieces per pixel
        unsigned g = 0, rg = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int h = tile / nwt, w0 = (tile - h * nwt) * CV_BM;
            const int x_lo = w0 + CV_BM - CV_BN * nchunks;
            const size_t rowbase = (size_t)h * W;
            for (int c = 0; c < nchunks; c++, g++) {
                const unsigned acc_i = g % CV_NACC;
                tc_mbar_wait(&sm.bar_full[acc_i], (g / CV_NACC) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const unsigned taddr = tmem_base + ((unsigned)(32 * q) << 16) + acc_i * CV_BN;
#pragma unroll 1
                for (int j = 0; j < CV_BN / 32; j++) {
                    const int xg = x_lo + CV_BN * c + 32 * j;       // band coordinate of the group's first column
                    // column i is right pixel xg + i.  R: the 128 lanes of the tile hold d = b0 - i + t, t = 32q + lane.
                    // L: this warp's lane holds d = base + lane - i of left pixel w0 + 32q + lane.
                    const int b0 = w0 - xg, base = b0 + 32 * q;
                    const bool r_on = b0 + 127 >= 0 && b0 - 31 < D;                    // (uniform over the eight warps)
                    const bool l_on = base + 31 >= 0 && base - 31 < D;                 // (warp uniform; implies r_on)
                    if (!r_on) continue;
                    const unsigned st = tc_smem_u32(sm.r_stage[rg & 1]);
                    rg++;
                    if (does_l ? l_on : true) {
                        unsigned v[32];
                        tc_tmem_ld32(taddr + 32 * j, v);
                        if (does_l) {
                            __syncwarp();                                              // the previous flush has read the ring
                            const int dl4 = (base + lane) << 2;
#pragma unroll
                            for (int i = 0; i < 32; i++) cv_sts(ring | ((dl4 - 4 * i) & (CV_RING * 4 - 4)), v[i]);
                            __syncwarp();
                        } else {
                            // (b0 - i) & 3 only depends on i & 3: four base addresses, the rest is an immediate offset
                            const unsigned mine = st + ((32 * q + lane) << 2);
                            const unsigned o0 = mine + ((b0 & 3) << 2), o1 = mine + (((b0 - 1) & 3) << 2),
                                           o2 = mine + (((b0 - 2) & 3) << 2), o3 = mine + (((b0 - 3) & 3) << 2);
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                cv_sts(o0 + i * (CV_RPITCH * 4), v[i]);
                                cv_sts(o1 + (i + 1) * (CV_RPITCH * 4), v[i + 1]);
                                cv_sts(o2 + (i + 2) * (CV_RPITCH * 4), v[i + 2]);
                                cv_sts(o3 + (i + 3) * (CV_RPITCH * 4), v[i + 3]);
                            }
                        }
                    }
                    tc_named_barrier(2, 256);
                    // Every warp copies rows 4 ew .. 4 ew + 3 of the stage: the part of each right pixel's run that lies inside
                    // the band, as float4 pieces from the first 16-byte boundary on (lane = piece) plus ONE predicated scalar
                    // instruction for the ends of all four rows (lanes 8 r .. 8 r + 2: head of row r, 8 r + 4 .. 8 r + 6: tail).
                    {
                        const int gi = 2 * c + j, xp = xg - dbase + 4 * ew;             // right pixel of row 4 ew
                        float *row0 = R + ((ptrdiff_t)rowbase + xp) * (ptrdiff_t)Dp;   // (not dereferenced outside the image)
                        const unsigned srow = st + ((4 * ew * CV_RPITCH) << 2) + (lane << 4);
                        const bool inside = xp >= 0 && xp + 4 <= W;                     // (warp uniform: all four pixels exist)
                        float4 body[4];
                        int off[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const unsigned t = tab.r[gi][4 * ew + k];
                            const bool on = lane < (int)((t >> 8) & 63) && (inside || (xp + k >= 0 && xp + k < W));
                            off[k] = on ? k * Dp + (int)((t >> 14) & 1023) + 4 * lane : -1;
                            if (on) body[k] = cv_lds128(srow + ((k * CV_RPITCH + (t & 255)) << 2));
                        }
                        const int rk = lane >> 3, e = lane & 7;
                        const unsigned te = tab.r[gi][4 * ew + rk];
                        const int head = (te >> 24) & 3, tail = (te >> 26) & 3, nb = (te >> 8) & 63;
                        // float index relative to the first piece: -head .. -1 | 4 nb .. 4 nb + tail - 1
                        const int fe = e < 3 ? (e < head ? e - head : 1 << 20) : (e >= 4 && e - 4 < tail ? 4 * nb + e - 4 : 1 << 20);
                        const bool eon = fe != (1 << 20) && (inside || (xp + rk >= 0 && xp + rk < W));
                        float ev = 0.f;
                        if (eon) ev = cv_lds(st + (((4 * ew + rk) * CV_RPITCH + (int)(te & 255) + fe) << 2));
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (off[k] >= 0) *reinterpret_cast<float4 *>(row0 + off[k]) = body[k];
                        if (eon) row0[rk * Dp + (int)((te >> 14) & 1023) + fe] = ev;
                    }
                    if (does_l && l_on) {
                        // the line [32 m, 32 m + 32) with m = floor(d of column 0 / 32) is complete for every pixel now:
                        // 4 pixels per instruction (p = 4 k + sub), 8 float4 pieces per pixel
                        const unsigned rq = tc_smem_u32(sm.l_ring[q]) + ((sub * CV_RING + 4 * l8) << 2);
                        const int wq = w0 + 32 * q + sub;                               // left pixel of k = 0
                        float *lrow = L + (rowbase + wq) * (size_t)Dp + 4 * l8;
                        const int t0 = base + sub;
                        float4 o[8];
                        int off[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const int dm = (t0 + 4 * k) & ~31;                          // 32 m (negative = band over)
                            const bool on = dm >= 0 && dm + 4 * l8 < Dp && wq + 4 * k < W;
                            off[k] = on ? 4 * k * Dp + dm : -1;
                            if (on) o[k] = cv_lds128(rq + ((4 * k * CV_RING + (dm & 32)) << 2));
                        }
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            if (off[k] >= 0) *reinterpret_cast<float4 *>(lrow + off[k]) = o[k];
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                tc_mbar_arrive(&sm.bar_empty[acc_i]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(CV_TMEM_COLS) : "memory");
}

// Invalid triangles (pf:94-95 for L, pf:105-106 for R), in the already negated domain (negation
// commutes exactly with the mean).  One warp per (row, 32 disparities); lanes over d so that the
// cells written at each step are contiguous; each lane slides a 3-value window along w.
__global__ void k_cost_fill(float *__restrict__ L, float *__restrict__ R, int H, int W, int D, int Dp, int dbase) {
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.y;
    const int d0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (d0 >= D) return;
    const int dl = d0 + lane;                                   // disparity inside the slab (the volume's index)
    const int d = dbase + dl;                                   // disparity (the triangle's extent)
    const bool live = dl < D && d >= 1;
    const int dmax = dbase + min(d0 + 31, D - 1);
    float *Lrow = L + (size_t)h * W * Dp;
    float *Rrow = R + (size_t)h * W * Dp;
    // the three valid cells each recurrence starts from (W >= ndisp + 2 keeps them inside the row): all loads first
    float l1 = 0.f, l2 = 0.f, l3 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
    if (live) {
        l1 = Lrow[(size_t)d * Dp + dl];                         // columns d, d+1, d+2
        l2 = Lrow[(size_t)(d + 1) * Dp + dl];
        l3 = Lrow[(size_t)(d + 2) * Dp + dl];
        r1 = Rrow[(size_t)(W - d - 1) * Dp + dl];               // columns W-d-1, W-d-2, W-d-3
        r2 = Rrow[(size_t)(W - d - 2) * Dp + dl];
        r3 = Rrow[(size_t)(W - d - 3) * Dp + dl];
    }
    // L: columns d-1 .. 0, right to left (pf:94-95); lanes over d so that each step writes a contiguous run
    for (int c = dmax - 1; c >= 0; c--) {
        if (live && c <= d - 1) {
            const float v = ((l1 + l2) + l3) / 3.0f;
            Lrow[(size_t)c * Dp + dl] = v;
            l3 = l2; l2 = l1; l1 = v;
        }
    }
    // R: columns W-d .. W-1, left to right (pf:105-106)
    for (int c = W - dmax; c < W; c++) {
        if (live && c >= W - d) {
            const float v = ((r3 + r2) + r1) / 3.0f;
            Rrow[(size_t)c * Dp + dl] = v;
            r3 = r2; r2 = r1; r1 = v;
        }
    }
}

}  // namespace mccnn

using namespace mccnn;

extern "C" {

static int cost_volume_slab(const float *fl, const float *fr, float *L, float *R, int H, int W, int C, int Dtot, int dbase,
                            int D, void *stream) {
    MCCNN_REQUIRE(fl && fr && L && R, "cost_volume: null pointer");
    MCCNN_REQUIRE(C == CV_C, "cost_volume: %d feature channels unsupported (the network emits 64, model.py:38)", C);
    MCCNN_REQUIRE(H >= 1 && Dtot >= 1 && W >= Dtot + 2, "cost_volume: need W >= ndisp + 2 (pf:94-95), got W=%d ndisp=%d", W, Dtot);
    MCCNN_REQUIRE(dbase >= 0 && D >= 1 && dbase + D <= Dtot, "cost_volume: slab [%d, %d) outside [0, %d)", dbase, dbase + D, Dtot);
    MCCNN_REQUIRE(D <= 512, "cost_volume: ndisp %d too large (max 512)", D);
    MCCNN_REQUIRE(((uintptr_t)fl & 15) == 0 && ((uintptr_t)fr & 15) == 0, "cost_volume: features must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const int Dp = dpitch(D);
    CvMaps maps;
    int rc = tc_encode_map_3d(maps.fl, fl, CV_C, W, H, 32, CV_BM, true, "cost_volume");
    if (rc) return rc;
    rc = tc_encode_map_3d(maps.fr, fr, CV_C, W, H, 32, CV_BN, true, "cost_volume");
    if (rc) return rc;
    // per device, asked on every call (no process-wide caches: one process may drive several GPUs)
    int dev = 0, num_sms = 0;
    MCCNN_CUDA(cudaGetDevice(&dev));
    MCCNN_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    MCCNN_CUDA(cudaFuncSetAttribute(k_cost_volume_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CvSmem)));
    const int nwt = cdiv(W, CV_BM), nchunks = cdiv(CV_BM - 1 + D, CV_BN);
    const long long ntiles = (long long)nwt * H;
    MCCNN_REQUIRE(ntiles < (1ll << 31), "cost_volume: image too large");
    const int grid = ntiles < num_sms ? (int)ntiles : num_sms;
    CvTable tab;
    for (int gi = 0; gi < 2 * nchunks; gi++)
        for (int i = 0; i < 32; i++) tab.r[gi][i] = cv_table_entry(CV_BN * nchunks - CV_BM - 32 * gi, i, D);
    k_cost_volume_tc<<<grid, CV_THREADS, sizeof(CvSmem), s>>>(maps, tab, L, R, H, W, D, Dp, nwt, nchunks, (int)ntiles, dbase);
    MCCNN_LAUNCHED("cost_volume_tc");
    if (dbase + D > 1) {
        dim3 fgrid(cdiv(cdiv(D, 32), 4), H);
        k_cost_fill<<<fgrid, 128, 0, s>>>(L, R, H, W, D, Dp, dbase);
        MCCNN_LAUNCHED("cost_fill");
    }
    return MCCNN_OK;
}

int mccnn_cost_volume(const float *fl, const float *fr, float *L, float *R, int H, int W, int C, int D, void *stream) {
    return cost_volume_slab(fl, fr, L, R, H, W, C, D, 0, D, stream);
}

int mccnn_cost_volume_slab(const float *fl, const float *fr, float *L, float *R, int H, int W, int C, int D, int d_base,
                           int d_count, void *stream) {
    return cost_volume_slab(fl, fr, L, R, H, W, C, D, d_base, d_count, stream);
}

}  // extern "C"
