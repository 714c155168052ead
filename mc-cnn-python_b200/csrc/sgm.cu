// Semi-global matching (pf:476-568, pf:187-235).  Compiled with -fmad=false; bit-exact against the
// reference: only +, -, min and four pre-divided penalty constants are involved, evaluated in the
// reference's order ((C + min(...)) - m, pf:552/559/566).
//
// Mapping: one warp per scanline, lanes over disparity.  In the HWD layout the D-vector of a pixel
// is contiguous, so every step of every direction reads and writes one coalesced 4*Dp-byte run;
// the four directions are the same kernel with a different pixel stride.  Lane l owns the float4
// granules g = l + NLg*j (j < JP), the previous pixel's path costs stay in registers, neighbours
// d-1 / d+1 that cross a granule come from two warp shuffles per granule, and the minimum over D
// is one redux.sync on an order-preserving integer key.  The input cells of the next PF steps are
// prefetched into a register ring so that the HBM latency is off the recurrence's critical path.
//
// Image-adaptive penalties (pf:504-541) are not materialised as [D,H,W] arrays: the two gradient
// tests D1 >= tauD / D2 >= tauD are precomputed once per pass as bit maps of the two images
// (k_sgm_flags) and looked up per cell.
#include "common.cuh"
#include <math_constants.h>

namespace mccnn {

struct SgmJob {
    float *vol;
    const uint32_t *own_map;   // gradient-flag bit map of the volume's own image (eh or ev)
    const uint32_t *oth_map;   // ... of the other image
    int is_left;
};

// Where a pass stores its result when the volume is about to be re-partitioned (one big pair over several GPUs):
// instead of writing in place and packing / sending / unpacking afterwards, the pass writes every cell straight
// into the buffer of the rank that needs it next, over NVLink peer memory.
//   mode 1 (horizontal pass on a ROW slab): pixel (h, w) goes to the rank that owns column w, into its column
//           slab [H][w_count_r][Dp] at row h_base + h;
//   mode 2 (vertical pass on a COLUMN slab): granule g of pixel (h, w) goes to the rank that owns granule g, into
//           its disparity slab [H][W][4 * g_count_r] at column w_base + w.
constexpr int SGM_MAX_PARTS = 8;
struct SgmScatter {
    int mode, nparts;
    int lo[SGM_MAX_PARTS + 1];               // column (mode 1) or granule (mode 2) bounds of the parts
    float4 *base[2][SGM_MAX_PARTS];          // [job][part]: the destination buffers
    int h_base, w_full;                      // mode 1: global row of this slab's row 0; mode 2: image width
};

struct SgmParams {
    SgmJob job[2];
    SgmScatter sc;
    int D, G, NLg, H, W, WR, PADW;   // W = width of the volume (and of the images unless wbase/flag geometry say otherwise)
    int wbase;                       // image column of the volume's column 0 (column slabs of one big pair)
    int rh, rw;
    float P1, P2, P1q1, P2q1, P1q2, P2q2;
};

// Bit maps.  eh(h,x) = x>=1 && |I(h,x)-I(h,x-1)| >= tauD ; ev(h,x) = h>=1 && |I(h,x)-I(h-1,x)| >= tauD.
// Rows are padded with PADW zero words on both sides so that x = w -/+ d outside the image reads 0,
// which is the reference's "D2 = 0 when the correspondent is outside" (pf:516-520, :529-533).
// maps: [image(2)][kind(2: eh, ev)][H][WR] u32.
__global__ void k_sgm_flags(const float *__restrict__ img0, const float *__restrict__ img1, uint32_t *__restrict__ maps,
                            int H, int W, int WR, int PADW, float tauD) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;      // bit position within the padded row
    const int h = blockIdx.y;
    const float *img = blockIdx.z ? img1 : img0;
    const int x = pos - PADW * 32;
    bool fh = false, fv = false;
    if (x >= 0 && x < W) {
        float c = img[(size_t)h * W + x];
        if (x >= 1) fh = fabsf(c - img[(size_t)h * W + x - 1]) >= tauD;
        if (h >= 1) fv = fabsf(c - img[(size_t)(h - 1) * W + x]) >= tauD;
    }
    unsigned bh = __ballot_sync(0xffffffffu, fh), bv = __ballot_sync(0xffffffffu, fv);
    const int word = pos >> 5;
    if ((threadIdx.x & 31) == 0 && word < WR) {
        size_t base = ((size_t)blockIdx.z * 2) * H * WR;
        maps[base + (size_t)h * WR + word] = bh;
        maps[base + (size_t)H * WR + (size_t)h * WR + word] = bv;
    }
}

__device__ __forceinline__ int f2key(float x) {
    int i = __float_as_int(x);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

__device__ __forceinline__ void sgm_cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}

// JP granules per lane, PF = look-ahead (steps) of the flag register ring, CP = slots of the shared-memory cell ring
// (cells are fetched CP - 1 steps ahead with cp.async: the pass is bound by bytes in flight -- there are only H or
// W scanlines x 2 volumes = 2048 warps on the whole chip -- and a ring in shared memory buys look-ahead that
// registers cannot).
template <int JP, int PF, int CP, bool SC = false>
__global__ void __launch_bounds__(32, 16) k_sgm_pass(const __grid_constant__ SgmParams prm) {
    __shared__ float4 cring[CP][JP][32];
    constexpr int DIST = CP - 1;
    const SgmJob job = prm.job[blockIdx.y];
    const int lane = threadIdx.x;
    const int line = blockIdx.x;
    const int G = prm.G, NLg = prm.NLg, W = prm.W, H = prm.H, D = prm.D, WR = prm.WR;
    const bool horizontal = (prm.rh == 0);
    const int N = horizontal ? W : H;
    int h0, w0, dh, dw;
    if (horizontal) { h0 = line; dh = 0; dw = prm.rw; w0 = prm.rw > 0 ? 0 : W - 1; }
    else            { w0 = line; dw = 0; dh = prm.rh; h0 = prm.rh > 0 ? 0 : H - 1; }
    const int hoff = prm.rh < 0 ? 1 : 0, woff = prm.rw < 0 ? 1 : 0;
    const long long pstride = (long long)dh * W + dw;            // pixels per step
    const long long p0 = (long long)h0 * W + w0;
    float4 *const vol4 = reinterpret_cast<float4 *>(job.vol);
    const float INF = CUDART_INF_F;

    int g[JP];
    bool gv[JP];
#pragma unroll
    for (int j = 0; j < JP; j++) { g[j] = lane + NLg * j; gv[j] = (lane < NLg) && (g[j] < G); }
    const bool ragged = (D & 3) != 0;
    const int src_dn = (lane == 0) ? NLg - 1 : lane - 1;
    const int src_up = (lane >= NLg - 1) ? 0 : lane + 1;
    const bool first_lane = (lane == 0), last_lane = (lane == NLg - 1);

    // scatter destinations (SC): mode 2 is fixed per lane granule for the whole pass, mode 1 changes with the pixel
    float4 *sc_dst[JP];                      // mode 2: destination of granule g[j] at pixel 0 of its slab
    int sc_pitch[JP];                        //         granules per pixel of that slab
    if (SC && prm.sc.mode == 2) {
#pragma unroll
        for (int j = 0; j < JP; j++) {
            sc_dst[j] = nullptr; sc_pitch[j] = 0;
            if (gv[j]) {
                int r = 0;
                while (r + 1 < prm.sc.nparts && g[j] >= prm.sc.lo[r + 1]) r++;
                sc_pitch[j] = prm.sc.lo[r + 1] - prm.sc.lo[r];
                sc_dst[j] = prm.sc.base[blockIdx.y][r] + (g[j] - prm.sc.lo[r]);
            }
        }
    }
    // cell (pixel index p of this slab = h * W + w, lane granule j) -> where it is stored
    auto store_cell = [&](long long p, int t, int j, const float4 &v) {
        if (!SC) { vol4[p * G + g[j]] = v; return; }
        const int h = h0 + t * dh, w = w0 + t * dw;              // pixel of step t on this scanline
        if (prm.sc.mode == 1) {
            int r = 0;
            while (r + 1 < prm.sc.nparts && w >= prm.sc.lo[r + 1]) r++;
            const int wc = prm.sc.lo[r + 1] - prm.sc.lo[r];
            prm.sc.base[blockIdx.y][r][((size_t)(prm.sc.h_base + h) * wc + (w - prm.sc.lo[r])) * G + g[j]] = v;
        } else {
            sc_dst[j][((size_t)h * prm.sc.w_full + prm.wbase + w) * sc_pitch[j]] = v;
        }
    };

    auto fix_ragged = [&](float4 &v, int j) {
        if (ragged) {
            int d = g[j] << 2;
            if (d + 1 >= D) v.y = INF;
            if (d + 2 >= D) v.z = INF;
            if (d + 3 >= D) v.w = INF;
        }
    };
    // cells of step t -> ring slot t % CP (one cp.async group per step, empty past the end of the scanline).  The slot
    // and the cell addresses are carried from step to step (no division, no 64-bit multiply in the recurrence).
    const long long stepG = pstride * G;                             // granules between consecutive pixels of the scanline
    const float4 *pre[JP];                                           // this lane's granules of the next step to prefetch
    float4 *cellp[JP];                                               // ... of the step being computed (in-place store)
#pragma unroll
    for (int j = 0; j < JP; j++) {
        pre[j] = vol4 + (p0 + pstride) * G + g[j];                   // step 1
        cellp[j] = vol4 + (p0 + pstride) * G + g[j];
    }
    auto prefetch_cells = [&](int t, int slot) {                     // must be called for t = 1, 2, 3, .. in order
        if (t < N) {
#pragma unroll
            for (int j = 0; j < JP; j++)
                if (gv[j]) sgm_cp_async16(&cring[slot][j][lane], pre[j]);
        }
#pragma unroll
        for (int j = 0; j < JP; j++) pre[j] += stepG;
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    auto take_cells = [&](int slot, float4 (&dst)[JP]) {
#pragma unroll
        for (int j = 0; j < JP; j++) {
            float4 v = make_float4(INF, INF, INF, INF);
            if (gv[j]) {
                v = cring[slot][j][lane];
                fix_ragged(v, j);
            }
            dst[j] = v;
        }
    };
    // Penalty selectors of step t.  The ring keeps the RAW bit-map words (loaded PF steps ahead) and the 4-bit
    // selectors of this lane's granules (bit k <-> d = 4g + k) are extracted only when the step is executed, so
    // that the loads are never waited for at issue.
    struct FlagWords { uint32_t lo[JP], hi[JP], own; };
    // Bit positions and bit-map rows are carried from step to step: the load stream runs PF steps ahead of the
    // extract stream, both start at step 1 and advance by one step per call.
    int goff[JP];                                                    // bit offset of this lane's granule: x = wb -/+ d
#pragma unroll
    for (int j = 0; j < JP; j++) goff[j] = job.is_left ? -4 * g[j] - 3 : 4 * g[j];
    const int pos_step1 = prm.PADW * 32 + prm.wbase + w0 + dw + woff;               // pos1 of step 1
    const long long row_step = (long long)dh * WR;
    int ld_pos1 = pos_step1, ex_pos1 = pos_step1;
    const uint32_t *ld_own = job.own_map + (long long)(h0 + dh + hoff) * WR;        // bit-map rows of step 1
    const uint32_t *ld_oth = job.oth_map + (long long)(h0 + dh + hoff) * WR;
    auto load_flags = [&](bool valid, FlagWords &fw) {
        if (valid) {
            fw.own = ld_own[ld_pos1 >> 5];
#pragma unroll
            for (int j = 0; j < JP; j++) {
                fw.lo[j] = 0; fw.hi[j] = 0;
                if (gv[j]) {
                    const int pos = ld_pos1 + goff[j];
                    fw.lo[j] = ld_oth[pos >> 5];
                    fw.hi[j] = ld_oth[(pos >> 5) + 1];
                }
            }
        }
        ld_pos1 += dw; ld_own += row_step; ld_oth += row_step;
    };
    auto extract_flags = [&](const FlagWords &fw, uint32_t (&dst)[JP], uint32_t &own) {
        own = (fw.own >> (ex_pos1 & 31)) & 1u;
#pragma unroll
        for (int j = 0; j < JP; j++) {
            uint32_t f = __funnelshift_r(fw.lo[j], fw.hi[j], (ex_pos1 + goff[j]) & 31) & 15u;
            if (job.is_left) f = __brev(f) >> 28;
            dst[j] = f;
        }
        ex_pos1 += dw;
    };

    // step 0: the first pixel of the scanline is left unchanged (pf:485-501) and seeds the recurrence
    float4 prev[JP];
    {
#pragma unroll
        for (int j = 0; j < JP; j++) {
            float4 v = make_float4(INF, INF, INF, INF);
            if (gv[j]) {
                v = vol4[p0 * G + g[j]];
                fix_ragged(v, j);
            }
            prev[j] = v;
            if (SC && gv[j]) store_cell(p0, 0, j, v);               // (the seed pixel travels unchanged)
        }
    }
#pragma unroll 1
    for (int t = 1; t <= DIST; t++) prefetch_cells(t, t);           // (DIST < CP: slot = t)
    int slot_t = 1;                                                  // ring slot of the step being computed
    float m;
    {
        float lm = INF;
#pragma unroll
        for (int j = 0; j < JP; j++) lm = fminf(fminf(lm, fminf(prev[j].x, prev[j].y)), fminf(prev[j].z, prev[j].w));
        m = key2f(__reduce_min_sync(0xffffffffu, f2key(lm)));
    }

    FlagWords fring[PF];
#pragma unroll
    for (int u = 0; u < PF; u++)
        load_flags(1 + u < N, fring[u]);

    for (int t0 = 1; t0 < N; t0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int t = t0 + u;
            if (t < N) {                                           // warp-uniform
                float4 cur[JP];
                uint32_t fb[JP];
#pragma unroll
                uint32_t f1;
                extract_flags(fring[u], fb, f1);
                load_flags(t + PF < N, fring[u]);
                asm volatile("cp.async.wait_group %0;\n" ::"n"(DIST - 1) : "memory");     // the cells of step t have landed
                take_cells(slot_t, cur);
                prefetch_cells(t + DIST, slot_t == 0 ? CP - 1 : slot_t - 1);               // into the slot step t - 1 used
                slot_t = slot_t == CP - 1 ? 0 : slot_t + 1;

                // penalties (pf:535-541): f1 = (D1 >= tauD), per-cell bit = (D2 >= tauD)
                const float pa1 = f1 ? prm.P1q1 : prm.P1, pb1 = f1 ? prm.P1q2 : prm.P1q1;
                const float cA = m + (f1 ? prm.P2q1 : prm.P2), cB = m + (f1 ? prm.P2q2 : prm.P2q1);
                float rw_[JP], rx_[JP];
#pragma unroll
                for (int j = 0; j < JP; j++) {
                    rw_[j] = __shfl_sync(0xffffffffu, prev[j].w, src_dn);
                    rx_[j] = __shfl_sync(0xffffffffu, prev[j].x, src_up);
                }
                float lm = INF;
                const long long p = SC ? p0 + (long long)t * pstride : 0;
#pragma unroll
                for (int j = 0; j < JP; j++) {
                    const float lnb = first_lane ? (j > 0 ? rw_[j > 0 ? j - 1 : 0] : INF) : rw_[j];
                    const float rnb = last_lane ? (j < JP - 1 ? rx_[j < JP - 1 ? j + 1 : 0] : INF) : rx_[j];
                    const float4 q = prev[j];
                    const uint32_t b = fb[j];
                    float4 o;
                    o.x = (cur[j].x + fminf(q.x, fminf(fminf(lnb, q.y) + ((b & 1u) ? pb1 : pa1), (b & 1u) ? cB : cA))) - m;
                    o.y = (cur[j].y + fminf(q.y, fminf(fminf(q.x, q.z) + ((b & 2u) ? pb1 : pa1), (b & 2u) ? cB : cA))) - m;
                    o.z = (cur[j].z + fminf(q.z, fminf(fminf(q.y, q.w) + ((b & 4u) ? pb1 : pa1), (b & 4u) ? cB : cA))) - m;
                    o.w = (cur[j].w + fminf(q.w, fminf(fminf(q.z, rnb) + ((b & 8u) ? pb1 : pa1), (b & 8u) ? cB : cA))) - m;
                    if (gv[j]) {
                        if (SC) store_cell(p, t, j, o);
                        else *cellp[j] = o;
                    }
                    cellp[j] += stepG;
                    prev[j] = o;
                    lm = fminf(fminf(lm, fminf(o.x, o.y)), fminf(o.z, o.w));
                }
                m = key2f(__reduce_min_sync(0xffffffffu, f2key(lm)));
            }
        }
    }
}

static int launch_pass(const SgmParams &prm, int njobs, cudaStream_t s) {
    const int JP = cdiv(prm.G, 32);
    const bool horizontal = (prm.rh == 0);
    dim3 grid(horizontal ? prm.H : prm.W, njobs), block(32);
    if (prm.sc.mode != 0) {
        switch (JP) {
            case 1: k_sgm_pass<1, 4, 16, true><<<grid, block, 0, s>>>(prm); break;
            case 2: k_sgm_pass<2, 4, 12, true><<<grid, block, 0, s>>>(prm); break;
            case 3: k_sgm_pass<3, 3, 8, true><<<grid, block, 0, s>>>(prm); break;
            case 4: k_sgm_pass<4, 2, 8, true><<<grid, block, 0, s>>>(prm); break;
            default:
                set_error("sgm: ndisp %d too large (max 512)", prm.D);
                return MCCNN_ERR_UNSUPPORTED;
        }
        return check_launch("sgm_pass_scatter");
    }
    switch (JP) {
        case 1: k_sgm_pass<1, 4, 16><<<grid, block, 0, s>>>(prm); break;
        case 2: k_sgm_pass<2, 4, 12><<<grid, block, 0, s>>>(prm); break;
        case 3: k_sgm_pass<3, 3, 8><<<grid, block, 0, s>>>(prm); break;
        case 4: k_sgm_pass<4, 2, 8><<<grid, block, 0, s>>>(prm); break;
        default:
            set_error("sgm: ndisp %d too large (max 512)", prm.D);
            return MCCNN_ERR_UNSUPPORTED;
    }
    return check_launch("sgm_pass");
}

struct SgmGeom { int WR, PADW; size_t map_words; };
static SgmGeom sgm_geom(int H, int W, int D) {
    SgmGeom g;
    g.PADW = cdiv(dpitch(D), 32) + 1;
    g.WR = 2 * g.PADW + cdiv(W, 32) + 1;
    g.map_words = (size_t)H * g.WR;
    return g;
}

static int build_flags(const float *img_left, const float *img_right, uint32_t *maps, int H, int W, int D, float tauD,
                       cudaStream_t s) {
    SgmGeom gm = sgm_geom(H, W, D);
    dim3 block(256), grid(cdiv((long long)gm.WR * 32, 256), H, 2);
    k_sgm_flags<<<grid, block, 0, s>>>(img_left, img_right, maps, H, W, gm.WR, gm.PADW, tauD);
    return check_launch("sgm_flags");
}

static int check_dir(int rh, int rw) {
    return (rh == 0 && (rw == 1 || rw == -1)) || (rw == 0 && (rh == 1 || rh == -1));
}

static void fill_params(SgmParams &prm, int D, int H, int W, int rh, int rw, double P1, double P2, double Q1, double Q2) {
    SgmGeom gm = sgm_geom(H, W, D);
    prm.D = D; prm.G = dpitch(D) / 4; prm.H = H; prm.W = W; prm.WR = gm.WR; prm.PADW = gm.PADW; prm.wbase = 0;
    prm.sc.mode = 0; prm.sc.nparts = 0;
    prm.NLg = cdiv(prm.G, cdiv(prm.G, 32));
    prm.rh = rh; prm.rw = rw;
    // float32 rounding exactly as pf:504-505 (P*ones(float32)) and pf:538-541 (float32 array / scalar)
    const float P1a = (float)P1, P2a = (float)P2, q1 = (float)Q1, q2 = (float)Q2;
    prm.P1 = P1a; prm.P2 = P2a;
    prm.P1q1 = P1a / q1; prm.P2q1 = P2a / q1; prm.P1q2 = P1a / q2; prm.P2q2 = P2a / q2;
}

static void fill_job(SgmJob &job, float *vol, uint32_t *maps, int H, int W, int D, int rh, int is_left) {
    SgmGeom gm = sgm_geom(H, W, D);
    const int kind = (rh == 0) ? 0 : 1;                       // eh for horizontal passes, ev for vertical
    const uint32_t *left = maps + (size_t)(0 * 2 + kind) * gm.map_words;
    const uint32_t *right = maps + (size_t)(1 * 2 + kind) * gm.map_words;
    job.vol = vol;
    job.is_left = is_left;
    job.own_map = is_left ? left : right;
    job.oth_map = is_left ? right : left;
}

}  // namespace mccnn

using namespace mccnn;

extern "C" {

size_t mccnn_sgm_scratch_bytes(int H, int W, int D) {
    if (H < 1 || W < 1 || D < 1) return 0;
    return 4 * sgm_geom(H, W, D).map_words * sizeof(uint32_t);
}

int mccnn_sgm_pass(float *vol, const float *img_left, const float *img_right, void *flags_scratch, int D, int H, int W,
                   int rh, int rw, double P1, double P2, double Q1, double Q2, double tauD, int is_left, void *stream) {
    MCCNN_REQUIRE(vol && img_left && img_right && flags_scratch, "sgm_pass: null pointer");
    MCCNN_REQUIRE(D >= 2 && H >= 1 && W >= 1, "sgm_pass: need ndisp >= 2 (pf:547-566), got D=%d H=%d W=%d", D, H, W);
    MCCNN_REQUIRE(check_dir(rh, rw), "sgm_pass: direction (%d,%d) is not axis-aligned (pf:484)", rh, rw);
    MCCNN_REQUIRE(tauD > 0.0, "sgm_pass: sgm_D must be > 0");
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *maps = (uint32_t *)flags_scratch;
    int rc = build_flags(img_left, img_right, maps, H, W, D, (float)tauD, s);
    if (rc) return rc;
    SgmParams prm;
    fill_params(prm, D, H, W, rh, rw, P1, P2, Q1, Q2);
    fill_job(prm.job[0], vol, maps, H, W, D, rh, is_left);
    prm.job[1] = prm.job[0];
    return launch_pass(prm, 1, s);
}

// Both volumes of a pair, four chained passes each (pf:194-208 and :212-230).  The final
// (X+X+X+X)/4. of pf:210/:232 is the identity on finite float32 and is not executed.
// vol_right may be NULL (left volume only) and vice versa.
int mccnn_sgm_average_pair(float *vol_left, float *vol_right, const float *img_left, const float *img_right,
                           void *flags_scratch, int D, int H, int W, double P1, double P2, double Q1, double Q2,
                           double tauD, double V, void *stream) {
    MCCNN_REQUIRE((vol_left || vol_right) && img_left && img_right && flags_scratch, "sgm_average: null pointer");
    MCCNN_REQUIRE(D >= 2 && H >= 1 && W >= 1, "sgm_average: need ndisp >= 2, got D=%d H=%d W=%d", D, H, W);
    MCCNN_REQUIRE(tauD > 0.0 && V != 0.0, "sgm_average: sgm_D must be > 0 and sgm_V non-zero");
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *maps = (uint32_t *)flags_scratch;
    int rc = build_flags(img_left, img_right, maps, H, W, D, (float)tauD, s);
    if (rc) return rc;
    const int dirs[4][2] = {{0, 1}, {0, -1}, {-1, 0}, {1, 0}};      // pf:195, :198, :203, :207
    for (int i = 0; i < 4; i++) {
        const int rh = dirs[i][0], rw = dirs[i][1];
        SgmParams prm;
        // vertical passes use P1/V formed in float64 by the caller (pf:204), then rounded to float32
        fill_params(prm, D, H, W, rh, rw, rh == 0 ? P1 : P1 / V, P2, Q1, Q2);
        int n = 0;
        if (vol_left) fill_job(prm.job[n++], vol_left, maps, H, W, D, rh, 1);
        if (vol_right) fill_job(prm.job[n++], vol_right, maps, H, W, D, rh, 0);
        if (n == 1) prm.job[1] = prm.job[0];
        rc = launch_pass(prm, n, s);
        if (rc) return rc;
    }
    return MCCNN_OK;
}

// Two of the four chained passes on a slab of one big pair: which = 0 runs (0,1) then (0,-1) on a
// ROW slab (volumes [h_count][W][Dp], images passed from their row h_base: rows are independent for horizontal
// passes), which = 1 runs (-1,0) then (1,0) on a COLUMN slab (volumes [H][w_count][Dp] holding image columns
// [w_base, w_base + w_count), full images: the penalty tests look up the other image at column w -/+ d).
static int sgm_passes_slab(float *vol_left, float *vol_right, const float *img_left, const float *img_right,
                           void *flags_scratch, int D, int H, int W, int w_base, int w_count, int which, double P1,
                           double P2, double Q1, double Q2, double tauD, double V, int nparts, const int *bounds,
                           float *const *dst_left, float *const *dst_right, int h_base, void *stream) {
    MCCNN_REQUIRE((vol_left || vol_right) && img_left && img_right && flags_scratch, "sgm_passes_slab: null pointer");
    MCCNN_REQUIRE(D >= 2 && H >= 1 && W >= 1, "sgm_passes_slab: need ndisp >= 2, got D=%d H=%d W=%d", D, H, W);
    MCCNN_REQUIRE(w_base >= 0 && w_count >= 1 && w_base + w_count <= W, "sgm_passes_slab: columns [%d, %d) outside the image",
                  w_base, w_base + w_count);
    MCCNN_REQUIRE(which == 0 || which == 1, "sgm_passes_slab: which must be 0 (horizontal) or 1 (vertical)");
    MCCNN_REQUIRE(which == 1 || (w_base == 0 && w_count == W), "sgm_passes_slab: horizontal passes need whole rows");
    MCCNN_REQUIRE(tauD > 0.0 && V != 0.0, "sgm_passes_slab: sgm_D must be > 0 and sgm_V non-zero");
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *maps = (uint32_t *)flags_scratch;
    int rc = build_flags(img_left, img_right, maps, H, W, D, (float)tauD, s);
    if (rc) return rc;
    const int dirs[2][2][2] = {{{0, 1}, {0, -1}}, {{-1, 0}, {1, 0}}};      // pf:195, :198 | pf:203, :207
    for (int i = 0; i < 2; i++) {
        const int rh = dirs[which][i][0], rw = dirs[which][i][1];
        SgmParams prm;
        fill_params(prm, D, H, W, rh, rw, rh == 0 ? P1 : P1 / V, P2, Q1, Q2);     // flag geometry of the full image
        prm.W = w_count;
        prm.wbase = w_base;
        int n = 0;
        if (vol_left) fill_job(prm.job[n++], vol_left, maps, H, W, D, rh, 1);
        if (vol_right) fill_job(prm.job[n++], vol_right, maps, H, W, D, rh, 0);
        if (n == 1) prm.job[1] = prm.job[0];
        if (i == 1 && nparts > 0) {                                              // the pair's last pass stores remotely
            prm.sc.mode = which == 0 ? 1 : 2;
            prm.sc.nparts = nparts;
            for (int r = 0; r <= nparts; r++) prm.sc.lo[r] = bounds[r];
            for (int r = 0; r < nparts; r++) {
                int m = 0;
                if (vol_left) prm.sc.base[m++][r] = reinterpret_cast<float4 *>(dst_left[r]);
                if (vol_right) prm.sc.base[m++][r] = reinterpret_cast<float4 *>(dst_right[r]);
                if (m == 1) prm.sc.base[1][r] = prm.sc.base[0][r];
            }
            prm.sc.h_base = h_base;
            prm.sc.w_full = W;
        }
        rc = launch_pass(prm, n, s);
        if (rc) return rc;
    }
    return MCCNN_OK;
}

int mccnn_sgm_passes_slab(float *vol_left, float *vol_right, const float *img_left, const float *img_right,
                          void *flags_scratch, int D, int H, int W, int w_base, int w_count, int which, double P1,
                          double P2, double Q1, double Q2, double tauD, double V, void *stream) {
    return sgm_passes_slab(vol_left, vol_right, img_left, img_right, flags_scratch, D, H, W, w_base, w_count, which, P1, P2,
                           Q1, Q2, tauD, V, 0, nullptr, nullptr, nullptr, 0, stream);
}

// The same, with the second pass of the pair storing every cell straight into the buffer of the rank that owns it
// in the next layout (peer memory over NVLink; see SgmScatter): which = 0 sends columns [bounds[r], bounds[r+1]) to
// dst_*[r], a column slab [h_total][bounds[r+1] - bounds[r]][Dp] whose row h_base + h receives this slab's row h;
// which = 1 sends granules [bounds[r], bounds[r+1]) to dst_*[r], a disparity slab [H][W][4 * (bounds[r+1] - bounds[r])].
// bounds and the dst tables are host arrays.
int mccnn_sgm_passes_slab_to(float *vol_left, float *vol_right, const float *img_left, const float *img_right,
                             void *flags_scratch, int D, int H, int W, int w_base, int w_count, int which, double P1,
                             double P2, double Q1, double Q2, double tauD, double V, int nparts, const int *bounds,
                             float *const *dst_left, float *const *dst_right, int h_base, void *stream) {
    MCCNN_REQUIRE(nparts >= 1 && nparts <= SGM_MAX_PARTS && bounds, "sgm_passes_slab_to: 1 to %d parts", SGM_MAX_PARTS);
    MCCNN_REQUIRE((!vol_left || dst_left) && (!vol_right || dst_right), "sgm_passes_slab_to: destination table missing");
    const int extent = which == 0 ? W : dpitch(D) / 4;
    MCCNN_REQUIRE(bounds[0] == 0 && bounds[nparts] == extent, "sgm_passes_slab_to: bounds must tile [0, %d)", extent);
    for (int r = 0; r < nparts; r++) {
        MCCNN_REQUIRE(bounds[r] < bounds[r + 1], "sgm_passes_slab_to: empty part %d", r);
        MCCNN_REQUIRE((!vol_left || dst_left[r]) && (!vol_right || dst_right[r]), "sgm_passes_slab_to: null destination %d", r);
    }
    return sgm_passes_slab(vol_left, vol_right, img_left, img_right, flags_scratch, D, H, W, w_base, w_count, which, P1, P2,
                           Q1, Q2, tauD, V, nparts, bounds, dst_left, dst_right, h_base, stream);
}

int mccnn_sgm_average(float *vol, const float *img_left, const float *img_right, void *flags_scratch, int D, int H,
                      int W, double P1, double P2, double Q1, double Q2, double tauD, double V, int is_left,
                      void *stream) {
    return mccnn_sgm_average_pair(is_left ? vol : nullptr, is_left ? nullptr : vol, img_left, img_right, flags_scratch, D,
                                  H, W, P1, P2, Q1, Q2, tauD, V, stream);
}

}  // extern "C"
