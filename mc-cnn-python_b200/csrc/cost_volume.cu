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
//       2.8-3.5 TB/s for 128-byte runs that start mid-sector).
//       R: a column of S is one right pixel x and the 32 lanes of a warp are consecutive left pixels w, i.e.
//          consecutive d = w - x: R[h][x][d..d+31] is one coalesced 128-byte store straight from tcgen05.ld registers
//          (it starts mid-sector 7 times out of 8: the slow kind).
//       L: a lane IS a left pixel, and all of its disparities are produced by this warp in this tile, 32 per column
//          group, descending.  The lane scatters them into its own 64-float ring in shared memory (bank = d mod 32:
//          conflict free); after every group exactly one aligned 32-float line per pixel is complete and the warp writes
//          those 32 lines as whole 128-byte lines (8 lanes x float4 per line).  Every L sector is written once, whole.
//       Round 1 formed S^T as well and wrote both volumes the R way (0.575 ms); staging R through shared memory too
//       (scripts/experiments/cost_volume_staged.cu) makes every store a whole sector but costs eight times the
//       instructions (0.72 ms): the epilogue becomes issue bound.  With L in whole sectors the stores stopped being the
//       bound: three right-chunk stages (a chunk is asked for two chunks ahead, so its flight from L2 overlaps the split
//       and the MMAs before it) took the kernel from 0.52 to 0.46 ms; what is left is the operand fetch itself (every
//       right pixel is read by 2.5 left tiles: 0.94 GB of L2 -> SM traffic per call, 0.21 ms with everything else knocked
//       out) plus MMAs and split that do not fully overlap it.  This mix measures 0.52 ms.
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
constexpr int CV_NSTAGE = 3;                // right-chunk stages in shared memory
constexpr int CV_NACC = 4;                 // TMEM accumulators (64 columns each)
constexpr int CV_TMEM_COLS = CV_NACC * CV_BN;
constexpr int CV_RING = 64;                // floats per pixel in an L warp's ring (two lines)

struct __align__(1024) CvSmem {
    unsigned char a_hi[CV_A_BYTES], a_lo[CV_A_BYTES];                  // left tile, split
    unsigned char a_raw[CV_A_BYTES];                                   // next left tile as loaded (prefetch)
    unsigned char b_hi[CV_NSTAGE][CV_B_BYTES], b_lo[CV_NSTAGE][CV_B_BYTES];   // right chunks (landed raw in b_hi, split in place)
    float l_ring[4][32 * CV_RING];                                     // per L warp: [pixel (lane)][d mod 64]
    unsigned long long bar_tma_a, bar_tma_b[CV_NSTAGE], bar_full[CV_NACC], bar_empty[CV_NACC];
    unsigned tmem_base;
};
static_assert(sizeof(CvSmem) <= 232448, "shared memory of k_cost_volume_tc");

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

// One warp writes its 32 TMEM lanes x 128 columns of an accumulator to a volume.  Column n is pixel pix0 + n;
// lane l holds disparity d0 + dl*l + dn*n with (dl, dn) = (+1, -1) for R (lanes = left pixels) and (-1, +1)
// for L (lanes = right pixels): the 32 lanes of one column are 32 consecutive disparities of one pixel,
// i.e. one coalesced 128-byte store.  lane_ok masks lanes whose own pixel lies outside the image.
template <int DL>
__device__ __forceinline__ void cv_store_quarter(float *__restrict__ vol, unsigned taddr, size_t rowbase, int pix0, int d0,
                                                 bool lane_ok, int W, int D, int Dp) {
    constexpr int DN = -DL;
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int grp = 0; grp < CV_BN / 32; grp++) {
        const int n0 = 32 * grp;
        const int dc = d0 + DN * n0;                            // d at (lane 0, column n0); the group spans dc - 31 .. dc + 31
        if (dc + 31 < 0 || dc - 31 >= D) continue;              // (warp uniform)
        if (pix0 + n0 >= W || pix0 + n0 + 31 < 0) continue;
        unsigned v[32];
        tc_tmem_ld32(taddr + n0, v);
        const int dl = dc + DL * lane;                          // this lane's d at column n0
        float *p = vol + ((ptrdiff_t)rowbase + pix0 + n0) * (ptrdiff_t)Dp + dl;
        const bool interior = dc - 31 >= 0 && dc + 31 < D && pix0 + n0 >= 0 && pix0 + n0 + 31 < W &&
                              __all_sync(0xffffffffu, lane_ok);
        if (interior) {
#pragma unroll
            for (int j = 0; j < 32; j++) p[(ptrdiff_t)j * (Dp + DN)] = __uint_as_float(v[j]);
        } else {
            // columns j whose cell exists: 0 <= dl + DN*j < D and 0 <= pix0 + n0 + j < W is one interval [ja, jb) per
            // lane; as a bit mask it costs one test per store instead of two index computations and two compares
            int ja = max(0, -(pix0 + n0)), jb = min(32, W - (pix0 + n0));
            if (DN > 0) { ja = max(ja, -dl); jb = min(jb, D - dl); }
            else        { ja = max(ja, dl - D + 1); jb = min(jb, dl + 1); }
            unsigned mask = 0;
            if (lane_ok && jb > ja) mask = (jb - ja >= 32 ? 0xffffffffu : ((1u << (jb - ja)) - 1u)) << ja;
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (mask & (1u << j)) p[(ptrdiff_t)j * (Dp + DN)] = __uint_as_float(v[j]);
        }
    }
}

__global__ void __launch_bounds__(CV_THREADS, 1)
k_cost_volume_tc(const __grid_constant__ CvMaps maps, float *__restrict__ L, float *__restrict__ R, int H, int W, int D,
                 int Dp, int nwt, int nchunks, int ntiles, int dbase) {
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
        for (int i = 0; i < CV_NSTAGE; i++) tc_mbar_init(&sm.bar_tma_b[i], 1);
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
    // chunk g of this CTA uses right-chunk stage g % CV_NSTAGE and accumulator g % CV_NACC; bar_full[g % CV_NACC] completes its
    // phase (g / CV_NACC) & 1 when the MMAs of chunk g are done (they have then also finished reading the stage)

    if (warp == 1) {
        // ================= TMA producer (one lane) =================
        if (lane == 0) {
            auto issue_b = [&](int tile, int c, unsigned g) {
                const int h = tile / nwt, w0 = (tile - h * nwt) * CV_BM;
                const int x0c = w0 + CV_BM - CV_BN * nchunks + CV_BN * c;
                const unsigned st = g % CV_NSTAGE;
                unsigned char *dst = sm.b_hi[st];
                tc_mbar_expect_tx(&sm.bar_tma_b[st], CV_B_BYTES);
                tc_tma_load_3d(dst, &maps.fr, 0, x0c - dbase, h, &sm.bar_tma_b[st]);
                tc_tma_load_3d(dst + CV_KB_BYTES, &maps.fr, 32, x0c - dbase, h, &sm.bar_tma_b[st]);
            };
            auto issue_a = [&](int tile) {
                const int h = tile / nwt, w0 = (tile - h * nwt) * CV_BM;
                tc_mbar_expect_tx(&sm.bar_tma_a, CV_A_BYTES);
                tc_tma_load_3d(sm.a_raw, &maps.fl, 0, w0, h, &sm.bar_tma_a);
                tc_tma_load_3d(sm.a_raw + CV_KA_BYTES, &maps.fl, 32, w0, h, &sm.bar_tma_a);
            };
            // the chunk to ask for next: CV_NSTAGE - 1 chunks ahead of the one whose MMAs are awaited, so that a chunk's
            // flight from L2 overlaps the split and the MMAs of the chunks before it (with two stages a chunk could only be
            // asked for when the splitters were already waiting for it: its whole latency was exposed, every chunk)
            int it = blockIdx.x, ic = 0;
            unsigned ik = 0;
            auto issue_next = [&]() {
                if (it >= ntiles) return;
                issue_b(it, ic, ik);
                ik++;
                if (++ic == nchunks) { ic = 0; it += gridDim.x; }
            };
            if ((int)blockIdx.x < ntiles) issue_a(blockIdx.x);
            for (int i = 0; i < CV_NSTAGE; i++) issue_next();
            unsigned g = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int c = 0; c < nchunks; c++, g++) {
                    if (g == 0) continue;
                    // the MMAs of chunk g-1 are done with its stage: chunk g-1 + CV_NSTAGE lands there
                    tc_mbar_wait_sleep(&sm.bar_full[(g - 1) % CV_NACC], ((g - 1) / CV_NACC) & 1);
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    // chunk g-1 opened a tile <=> c == 1 (or nchunks == 1): its split has consumed a_raw
                    const bool opened = (nchunks == 1) ? true : (c == 1);
                    const int opened_tile = (nchunks == 1) ? tile - (int)gridDim.x : tile;
                    if (opened && opened_tile + (int)gridDim.x < ntiles) issue_a(opened_tile + gridDim.x);
                    issue_next();
                }
            }
        }
    } else if (warp < 8) {
        // ================= operand split (warps 2-7) and MMA issue (warp 0, lane 0) =================
        unsigned g = 0, ta = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ta++) {
            for (int c = 0; c < nchunks; c++, g++) {
                const unsigned stage = g % CV_NSTAGE, acc_i = g % CV_NACC;
                if (warp >= 2) {
                    if (c == 0) {
                        // the previous tile's MMAs no longer read a_hi / a_lo; the new tile was prefetched into a_raw
                        if (g > 0) tc_mbar_wait(&sm.bar_full[(g - 1) % CV_NACC], ((g - 1) / CV_NACC) & 1);
                        tc_mbar_wait(&sm.bar_tma_a, ta & 1);
                        tc_split(sm.a_raw, sm.a_hi, sm.a_lo, CV_A_BYTES, tid - 64, CV_NSPLIT);
                    }
                    tc_mbar_wait(&sm.bar_tma_b[stage], (g / CV_NSTAGE) & 1);
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
        // ================= epilogue: R straight from registers, L through the per-lane rings =================
        const int q = warp & 3;                       // TMEM lane quarter this warp may read: left pixels w0 + 32q + lane
        const bool does_l = warp >= 12;               // warps 8-11 write R, warps 12-15 write L
        const unsigned ring = tc_smem_u32(sm.l_ring[q]) + lane * (CV_RING * 4);      // this lane's (= left pixel's) ring
        const int sub = lane >> 3, l8 = lane & 7;     // flush of L: 4 pixels per instruction, 8 float4 pieces per pixel
        unsigned g = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int h = tile / nwt, w0 = (tile - h * nwt) * CV_BM;
            const int x_lo = w0 + CV_BM - CV_BN * nchunks;
            const size_t rowbase = (size_t)h * W;
            for (int c = 0; c < nchunks; c++, g++) {
                const unsigned acc_i = g % CV_NACC;
                const int x0c = x_lo + CV_BN * c;
                tc_mbar_wait(&sm.bar_full[acc_i], (g / CV_NACC) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const unsigned taddr = tmem_base + ((unsigned)(32 * q) << 16) + acc_i * CV_BN;
                if (!does_l) {
                    // R[h][x][d]: lanes are left pixels w = w0 + 32q + lane, columns right pixels x = x0c + n, d = w - x
                    cv_store_quarter<+1>(R, taddr, rowbase, x0c - dbase, w0 + 32 * q - x0c, w0 + 32 * q + lane < W, W, D, Dp);
                } else {
#pragma unroll 1
                    for (int j = 0; j < CV_BN / 32; j++) {
                        const int base = w0 + 32 * q - (x0c + 32 * j);                 // d of (lane 0, column 0)
                        if (base + 31 < 0 || base - 31 >= D) continue;                 // (warp uniform)
                        unsigned v[32];
                        tc_tmem_ld32(taddr + 32 * j, v);
                        __syncwarp();                                                  // the previous flush has read the ring
                        const int dl4 = (base + lane) << 2;
#pragma unroll
                        for (int i = 0; i < 32; i++) cv_sts(ring | ((dl4 - 4 * i) & (CV_RING * 4 - 4)), v[i]);
                        __syncwarp();
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
    k_cost_volume_tc<<<grid, CV_THREADS, sizeof(CvSmem), s>>>(maps, L, R, H, W, D, Dp, nwt, nchunks, (int)ntiles, dbase);
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
