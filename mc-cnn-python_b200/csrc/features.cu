// Siamese feature network, forward only (model.py:40-64, :90-125; pf:15-73):
// zero-pad the image by num_layers pixels once (pf:20-25), then num_layers 3x3 VALID
// cross-correlations with 64 maps (NHWC activations, HWIO weights, stride 1), bias, ReLU after all
// but the last (model.py:51-60), and x * rsqrt(max(sum_c x^2, 1e-12)) over channels (model.py:64).
//
// k_conv1 (1 -> 64) is a bandwidth-bound float32 SIMT kernel.  Layers 2..n (64 -> 64, 99.8 % of the
// flops) run on the tensor cores as implicit GEMMs; one TF32/BF16/FP16 product would break the 1e-4 cost-volume
// tolerance, so operands are split into hi + lo and three products are accumulated in float32:
// k_conv64_h (default: fp16 hi/lo, kind::f16, half the MMAs) and k_conv64_tc (tf32 hi/lo, MCCNN_CONV_TF32=1).
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>
#include "tc_common.cuh"

namespace mccnn {

constexpr int F = 64;              // feature maps (model.py:38)
constexpr float CONV_F16_MAX = 6.0e4f;   // operands beyond this go through the TF32 kernel (fp16 overflows at 65504)

// Layer 1: 1 -> 64.  A thread owns 4 output channels -- their 36 weights and 4 biases stay in registers -- and
// walks C1_PX pixels of a row, 16 apart (the 16 channel quads of a pixel are 16 consecutive lanes: every store of a
// warp is 512 contiguous bytes).  grid = (pixel groups of a row, rows).
constexpr int C1_PX = 8;
__global__ void __launch_bounds__(256) k_conv1(const float *__restrict__ img, const float *__restrict__ wgt,
                                               const float *__restrict__ bias, float *__restrict__ out, int H, int W,
                                               int pad, int OH, int OW, int relu, int *__restrict__ big) {
    const int oc = (threadIdx.x & 15) * 4, slot = threadIdx.x >> 4, y = blockIdx.y;
    float4 wv[9];
#pragma unroll
    for (int t = 0; t < 9; t++) wv[t] = *reinterpret_cast<const float4 *>(wgt + t * F + oc);
    const float4 bv = *reinterpret_cast<const float4 *>(bias + oc);
    float4 *orow = reinterpret_cast<float4 *>(out) + (size_t)y * OW * 16 + (oc >> 2);
    unsigned mxb = 0;
#pragma unroll 2
    for (int it = 0; it < C1_PX; it++) {
        const int x = (blockIdx.x * C1_PX + it) * 16 + slot;
        if (x >= OW) break;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                const int iy = y + ky - pad, ix = x + kx - pad;   // coordinates in the unpadded image
                const float v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? img[(size_t)iy * W + ix] : 0.0f;
                const float4 w4 = wv[ky * 3 + kx];
                sum.x = fmaf(v, w4.x, sum.x); sum.y = fmaf(v, w4.y, sum.y);
                sum.z = fmaf(v, w4.z, sum.z); sum.w = fmaf(v, w4.w, sum.w);
            }
        float4 acc = make_float4(bv.x + sum.x, bv.y + sum.y, bv.z + sum.z, bv.w + sum.w);
        if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
        // (the largest activation written, as bits: after the ReLU they are non-negative, so unsigned order = float order)
        mxb = max(max(mxb, __float_as_uint(acc.x)), max(__float_as_uint(acc.y), max(__float_as_uint(acc.z), __float_as_uint(acc.w))));
        orow[(size_t)x * 16] = acc;
    }
    // an activation beyond fp16's range sends the next layer to the TF32 kernel (see k_conv64_h); only asked for with the ReLU
    if (big && relu && mxb > __float_as_uint(CONV_F16_MAX)) *big = 1;
}

// ------------------------------------------------------------------------------------------
// Layers 2..n on the tensor cores.  A 3x3 VALID convolution with 64 maps in and out is the implicit GEMM
// out[p][oc] = sum_{tap, ic} in[p + tap][ic] * w[tap][ic][oc]  (M = pixels, N = 64, K = 9 * 64), dense enough
// for tcgen05; float32 accuracy is kept with the three-way TF32 split (hi.hi + hi.lo + lo.hi, float32
// accumulation in TMEM), as in the cost-volume kernel.
//
// k_conv64_tc: persistent, one CTA per SM.  A tile is 4 output rows x 126 output pixels; its four
// accumulators (128 lanes = pixels, 64 columns = output maps each) live in TMEM, double buffered across
// tiles so that the epilogue of a tile overlaps the next tile's MMAs.  The input channels are processed in
// two blocks of 32 (one 128-byte swizzled row per pixel): for a channel block the weights of all nine taps,
// pre-split into hi / lo (k_conv_prep_weights), are resident in shared memory (144 KB) and the six input
// rows the tile touches stream through a double-buffered 128-pixel row buffer that is split in place.  The
// three horizontal taps are the same rows read through descriptors whose start address is shifted by
// 0 / 1 / 2 pixels (the 128B swizzle is a function of the absolute shared-memory address, so a start that
// is not 1024-byte aligned needs no base offset -- checked on the device), which is why a tile yields 126
// pixels, not 128.  Tiles are taken in pairs, block 0 of both then block 1 of both, so the weights are
// reloaded once per tile, not twice, and the accumulation order is the same for every tile.  An input row feeds up to three output rows (one per kernel row); the
// weights of a kernel column are stacked by kernel row in shared memory and the accumulators of consecutive
// output rows are adjacent in TMEM, so ONE tcgen05.mma with N = 64 / 128 / 192 updates all of them -- small-N
// MMAs do not run proportionally faster, so this halves the MMA count and the time.  Accumulators are zeroed
// through tcgen05.st by the epilogue, so every MMA accumulates.
//   warp 0: MMA issue (one lane); warp 1: TMA producer (one lane); warps 2-7: operand split;
//   warps 8-15: epilogue -- bias, ReLU or (last layer) channel L2 normalisation, 256-bit stores.
// ------------------------------------------------------------------------------------------
constexpr int CT_ROWS = 4;                      // output rows per tile
constexpr int CT_PIX = 126;                     // output pixels per tile row (128 loaded - 2 for the horizontal taps)
constexpr int CT_ROW_BYTES = 128 * 128;         // one input row block: 128 pixels x 32 channels, 128B-swizzled
constexpr int CT_WBLK_BYTES = 64 * 128;         // one weight block: 64 output maps x 32 input channels
constexpr int CT_W_BYTES = 9 * 2 * CT_WBLK_BYTES;   // nine taps, hi and lo: 144 KB
constexpr int CT_THREADS = 512;
constexpr int CT_NSPLIT = 192;

struct __align__(1024) CtcSmem {
    unsigned char w[2][3][3][CT_WBLK_BYTES];     // [hi, lo][kx][2 - ky]: the three kernel rows of a column stacked, ky descending
    unsigned char a_hi[2][CT_ROW_BYTES], a_lo[2][CT_ROW_BYTES];
    unsigned long long bar_w, bar_row[2], bar_rowdone[2], bar_full[2], bar_empty[2], bar_zero;
    unsigned tmem_base;
};

struct CtcMaps { CUtensorMap in, whi, wlo; };   // in: [IH][IW][64], box {32, 128, 1}; w: [9][64 oc][64 ic], box {32, 64, 1}

// HWIO [3][3][ic][oc] -> [tap][oc][ic], split into hi = tf32(w) and lo = w - hi
__global__ void k_conv_prep_weights(const float *__restrict__ w, float *__restrict__ whi, float *__restrict__ wlo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;        // index into [tap][oc][ic]
    if (i >= 9 * F * F) return;
    const int ic = i % F, oc = (i / F) % F, tap = i / (F * F);
    const float x = w[((size_t)tap * F + ic) * F + oc];
    const float h = __uint_as_float(tc_tf32(x));
    whi[i] = h;
    wlo[i] = x - h;
}

template <bool LAST>
__global__ void __launch_bounds__(CT_THREADS, 1)
k_conv64_tc(const __grid_constant__ CtcMaps maps, const float *__restrict__ bias, float *__restrict__ out, int OH, int OW,
            int ntx, int ntiles, const int *__restrict__ in_big, const int *__restrict__ w_big, int run_if, int *__restrict__ big) {
    // *in_big | *w_big != 0: this layer's input (or its weights) does not fit fp16 -- the TF32 kernel (run_if = 1) does
    // the layer, the FP16 kernel (run_if = 0) leaves at once; and the other way round.  in_big == NULL: run.
    if (in_big && ((*in_big | *w_big) != 0) != (run_if != 0)) return;
    extern __shared__ __align__(1024) unsigned char ctc_raw[];
    CtcSmem &sm = *reinterpret_cast<CtcSmem *>(ctc_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(&sm.tmem_base)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        tc_mbar_init(&sm.bar_w, 1);
        tc_mbar_init(&sm.bar_zero, 1);
        for (int i = 0; i < 2; i++) {
            tc_mbar_init(&sm.bar_row[i], 1);
            tc_mbar_init(&sm.bar_rowdone[i], 1);
            tc_mbar_init(&sm.bar_full[i], 1);
            tc_mbar_init(&sm.bar_empty[i], 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = sm.tmem_base;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128; N (64, 128 or 192) is filled in per MMA
    const unsigned idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(128 >> 4) << 24);
    constexpr int NROW = CT_ROWS + 2;           // input rows per tile

    // A CTA takes its tiles two at a time (one per accumulator buffer) and visits  block 0 of both, then block 1 of both:
    // the weights are reloaded once per tile on average, and EVERY tile accumulates block 0 before block 1 -- a pixel's
    // features do not depend on where its tile falls in some CTA's sequence, so a row band of an image (slab.py) gives
    // the bits of the whole image.  Row steps are numbered globally in that order.
    if (warp == 1) {
        // ================= TMA producer (one lane) =================
        if (lane == 0) {
            unsigned rr = 0, wloads = 0;
            int resident = -1;
            for (int tile0 = blockIdx.x; tile0 < ntiles; tile0 += 2 * gridDim.x) {
              for (int kb = 0; kb < 2; kb++) {
                for (int mem = 0; mem < 2; mem++) {
                    const int tile = tile0 + mem * gridDim.x;
                    if (tile >= ntiles) break;
                    const int ty = tile / ntx, y0 = ty * CT_ROWS, x0 = (tile - ty * ntx) * CT_PIX;
                    if (kb != resident) {
                        // every MMA issued so far reads the resident weights: wait for the last row step
                        if (rr > 0) {
                            tc_mbar_wait_sleep(&sm.bar_rowdone[(rr - 1) & 1], ((rr - 1) >> 1) & 1);
                            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                        }
                        tc_mbar_expect_tx(&sm.bar_w, CT_W_BYTES);
                        for (int tap = 0; tap < 9; tap++) {
                            tc_tma_load_3d(sm.w[0][tap % 3][2 - tap / 3], &maps.whi, kb * 32, 0, tap, &sm.bar_w);
                            tc_tma_load_3d(sm.w[1][tap % 3][2 - tap / 3], &maps.wlo, kb * 32, 0, tap, &sm.bar_w);
                        }
                        resident = kb;
                        wloads++;
                    }
                    for (int r = 0; r < NROW; r++, rr++) {
                        // the row buffer was last used by step rr - 2
                        if (rr >= 2) {
                            tc_mbar_wait_sleep(&sm.bar_rowdone[rr & 1], ((rr - 2) >> 1) & 1);
                            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                        }
                        tc_mbar_expect_tx(&sm.bar_row[rr & 1], CT_ROW_BYTES);
                        tc_tma_load_3d(sm.a_hi[rr & 1], &maps.in, kb * 32, x0, y0 + r, &sm.bar_row[rr & 1]);
                    }
                }
              }
            }
        }
    } else if (warp < 8) {
        // ================= operand split (warps 2-7) and MMA issue (warp 0, lane 0) =================
        unsigned rr = 0, wloads = 0, pair = 0;
        int resident = -1;
        for (int tile0 = blockIdx.x; tile0 < ntiles; tile0 += 2 * gridDim.x, pair++) {
          for (int kb = 0; kb < 2; kb++) {
            for (int mem = 0; mem < 2; mem++) {
                if (tile0 + mem * (int)gridDim.x >= ntiles) break;
                const unsigned abuf = mem, tt = 2 * pair + mem;
                const int bi = kb;
                bool new_weights = false;
                if (kb != resident) { new_weights = true; resident = kb; }
                for (int r = 0; r < NROW; r++, rr++) {
                    const unsigned rb = rr & 1;
                    if (warp >= 2) {
                        tc_mbar_wait(&sm.bar_row[rb], (rr >> 1) & 1);
                        tc_split(sm.a_hi[rb], sm.a_hi[rb], sm.a_lo[rb], CT_ROW_BYTES, tid - 64, CT_NSPLIT);
                        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    }
                    tc_named_barrier(1, 32 + CT_NSPLIT);
                    if (tid == 0) {
                        if (new_weights && r == 0) { tc_mbar_wait(&sm.bar_w, wloads & 1); }
                        if (bi == 0 && r == 0 && tt >= 2) tc_mbar_wait(&sm.bar_empty[abuf], ((tt >> 1) - 1) & 1);
                        if (rr == 0) tc_mbar_wait(&sm.bar_zero, 0);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        const unsigned long long ah = tc_smem_desc(tc_smem_u32(sm.a_hi[rb])), al = tc_smem_desc(tc_smem_u32(sm.a_lo[rb]));
                        // Input row r feeds accumulator j = r - ky for every kernel row ky it can pair with: the weights
                        // of those kernel rows are stacked (ky descending = j ascending), so ONE MMA with N = 64, 128 or
                        // 192 updates the adjacent accumulators j = r - kymax .. r - kymin.  Accumulators start from the
                        // zeros the epilogue leaves in TMEM, so every MMA accumulates.
                        const int kymax = min(2, r), kymin = max(0, r - (CT_ROWS - 1));
                        const unsigned nacc = (unsigned)(kymax - kymin + 1);
                        const unsigned idesc_n = idesc_base | ((nacc * F >> 3) << 17);
                        const unsigned d_tmem = tmem_base + abuf * (CT_ROWS * F) + (unsigned)(r - kymax) * F;
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) {
                            const unsigned long long wh = tc_smem_desc(tc_smem_u32(sm.w[0][kx][2 - kymax]));
                            const unsigned long long wl = tc_smem_desc(tc_smem_u32(sm.w[1][kx][2 - kymax]));
#pragma unroll
                            for (int ks = 0; ks < 4; ks++) {
                                const unsigned long long aoff = (unsigned long long)((kx * 128 + ks * 32) >> 4);
                                const unsigned long long woff = (unsigned long long)((ks * 32) >> 4);
                                tc_mma_tf32(d_tmem, ah + aoff, wh + woff, idesc_n, 1);
                                tc_mma_tf32(d_tmem, ah + aoff, wl + woff, idesc_n, 1);
                                tc_mma_tf32(d_tmem, al + aoff, wh + woff, idesc_n, 1);
                            }
                        }
                        tc_mma_commit(&sm.bar_rowdone[rb]);
                        if (bi == 1 && r == NROW - 1) tc_mma_commit(&sm.bar_full[abuf]);
                    }
                }
                if (new_weights) wloads++;
            }
          }
        }
    } else {
        // ================= epilogue (warps 8-15): two warps per TMEM lane quarter, two output rows each =================
        const int q = warp & 3, half = (warp - 8) >> 2;
        const int m = 32 * q + lane;
        const float4 *b4 = reinterpret_cast<const float4 *>(bias);
        // MMAs only ever accumulate: this warp's share of both accumulator buffers (its lane quarter, its two rows)
        // starts as zeros and is zeroed again after every read
        auto zero_acc = [&](unsigned abuf, int j) {
            const unsigned taddr = tmem_base + ((unsigned)(32 * q) << 16) + abuf * (CT_ROWS * F) + j * F;
            tc_tmem_st32_zero(taddr);
            tc_tmem_st32_zero(taddr + 32);
        };
        for (unsigned ab = 0; ab < 2; ab++)
            for (int jj = 0; jj < CT_ROWS / 2; jj++) zero_acc(ab, half * (CT_ROWS / 2) + jj);
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        tc_named_barrier(2, 256);                       // all eight epilogue warps have zeroed their part ...
        if (tid == 256) tc_mbar_arrive(&sm.bar_zero);    // ... tell the MMA issuer
        unsigned tt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tt++) {
            const unsigned abuf = tt & 1;
            const int ty = tile / ntx, y0 = ty * CT_ROWS, x0 = (tile - ty * ntx) * CT_PIX;
            tc_mbar_wait(&sm.bar_full[abuf], (tt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
            for (int jj = 0; jj < CT_ROWS / 2; jj++) {
                const int j = half * (CT_ROWS / 2) + jj;
                const int y = y0 + j, x = x0 + m;
                unsigned v0[32], v1[32];
                const unsigned taddr = tmem_base + ((unsigned)(32 * q) << 16) + abuf * (CT_ROWS * F) + j * F;
                tc_tmem_ld32(taddr, v0);
                tc_tmem_ld32(taddr + 32, v1);
                zero_acc(abuf, j);
                float ss = 0.f;
#pragma unroll
                for (int o = 0; o < 32; o += 4) {
                    const float4 ba = __ldg(b4 + (o >> 2)), bb = __ldg(b4 + 8 + (o >> 2));
                    float t;
                    t = __uint_as_float(v0[o]) + ba.x;     v0[o] = __float_as_uint(t);     ss = fmaf(t, t, ss);
                    t = __uint_as_float(v0[o + 1]) + ba.y; v0[o + 1] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v0[o + 2]) + ba.z; v0[o + 2] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v0[o + 3]) + ba.w; v0[o + 3] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v1[o]) + bb.x;     v1[o] = __float_as_uint(t);     ss = fmaf(t, t, ss);
                    t = __uint_as_float(v1[o + 1]) + bb.y; v1[o + 1] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v1[o + 2]) + bb.z; v1[o + 2] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v1[o + 3]) + bb.w; v1[o + 3] = __float_as_uint(t); ss = fmaf(t, t, ss);
                }
                const float inv = LAST ? 1.0f / sqrtf(fmaxf(ss, 1e-12f)) : 1.0f;   // model.py:64
                if (y < OH && m < CT_PIX && x < OW) {
                    unsigned mxb = 0;                                               // largest activation written, as bits
                    float *dst = out + ((size_t)y * OW + x) * F;
#pragma unroll
                    for (int o = 0; o < 32; o += 8) {
                        float e[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const float t = __uint_as_float(v0[o + k]);
                            e[k] = LAST ? t * inv : fmaxf(t, 0.f);                  // model.py:120-123
                            if (!LAST) mxb = max(mxb, __float_as_uint(e[k]));   // (non-negative after the ReLU)
                        }
                        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + o), "f"(e[0]), "f"(e[1]),
                                     "f"(e[2]), "f"(e[3]), "f"(e[4]), "f"(e[5]), "f"(e[6]), "f"(e[7])
                                     : "memory");
                    }
#pragma unroll
                    for (int o = 0; o < 32; o += 8) {
                        float e[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const float t = __uint_as_float(v1[o + k]);
                            e[k] = LAST ? t * inv : fmaxf(t, 0.f);
                            if (!LAST) mxb = max(mxb, __float_as_uint(e[k]));   // (non-negative after the ReLU)
                        }
                        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 32 + o), "f"(e[0]), "f"(e[1]),
                                     "f"(e[2]), "f"(e[3]), "f"(e[4]), "f"(e[5]), "f"(e[6]), "f"(e[7])
                                     : "memory");
                    }
                    if (!LAST && big && mxb > __float_as_uint(CONV_F16_MAX)) *big = 1;   // the next layer must not use fp16 operands
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            tc_mbar_arrive(&sm.bar_empty[abuf]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------
// The same kernel with FP16 operands (k_conv64_h).  hi = fp16(x) and lo = fp16(x - hi) carry the same 11 + 11
// significant bits as the TF32 split, so hi.hi + hi.lo + lo.hi in float32 has the same accuracy, but kind::f16 MMAs
// take K = 16 per instruction: half as many MMAs for the same contraction.  A 32-channel block is a 64-byte row
// (SWIZZLE_64B); the weights of BOTH channel blocks fit in shared memory at once (144 KB) and are loaded once per CTA
// instead of once per tile and block.  Range: fp16 overflows at 65504 -- activations of a trained MC-CNN are O(1..10);
// mccnn_features keeps the TF32 kernel behind MCCNN_CONV_TF32=1 for networks outside that range.  Values below 6e-5
// lose relative (not absolute) precision in the lo part: an absolute error of 3e-8 per operand, far below the 2e-5 gate.
// ------------------------------------------------------------------------------------------
constexpr int CH_TILE_BYTES = 128 * 64;         // fp16 operand tile of one input row: 128 pixels x 32 channels, 64B-swizzled
constexpr int CH_WBLK_BYTES = 64 * 64;          // one weight block: 64 output maps x 32 input channels, fp16
constexpr int CH_W_BYTES = 2 * 9 * 2 * CH_WBLK_BYTES;   // both channel blocks, nine taps, hi and lo: 144 KB
constexpr int CH_NRAW = 3;                      // float32 row buffers

struct __align__(1024) ChSmem {
    unsigned char w[2][2][3][3][CH_WBLK_BYTES];  // [channel block][hi, lo][kx][2 - ky]
    unsigned char raw[CH_NRAW][CT_ROW_BYTES];    // input rows as loaded (float32, 128B-swizzled): a ring of their own, so that
                                                 // a row can be asked for before the MMAs two steps back have finished
    unsigned char a_hi[2][CH_TILE_BYTES], a_lo[2][CH_TILE_BYTES];
    unsigned long long bar_w, bar_row[CH_NRAW], bar_rawfree[CH_NRAW], bar_rowdone[2], bar_full[2], bar_empty[2], bar_zero;
    unsigned tmem_base;
};

// K-major, 64B-swizzled operand: rows of 64 bytes, 8-row groups 512 bytes apart
__device__ __forceinline__ unsigned long long ch_smem_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
    d |= (unsigned long long)1 << 16;
    d |= (unsigned long long)(512 >> 4) << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)4 << 61;                     // SWIZZLE_64B
    return d;
}
__device__ __forceinline__ void ch_mma_f16(unsigned tmem_d, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc,
                                           unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// raw: [128 pixels][32 channels] float32 with the 128B swizzle of the TMA load (16-byte chunk c of pixel p sits at chunk
// c ^ (p & 7)); hi / lo: [128 pixels][32 channels] fp16 with the 64B swizzle (chunk c at c ^ ((p >> 1) & 3))
__device__ __forceinline__ void ch_split(const unsigned char *raw, unsigned char *hi, unsigned char *lo, int ftid, int nthr) {
    const float4 *r4 = reinterpret_cast<const float4 *>(raw);
#pragma unroll 2
    for (int i = ftid; i < CT_ROW_BYTES / 16; i += nthr) {
        const int p = i >> 3, cg = (i & 7) ^ (p & 7);            // pixel, channel group (4 channels)
        const float4 x = r4[i];
        const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2half2_rn(x.z - f23.x, x.w - f23.y);
        const int off = p * 64 + ((((cg >> 1) ^ ((p >> 1) & 3))) << 4) + ((cg & 1) << 3);
        *reinterpret_cast<uint2 *>(hi + off) = make_uint2(*reinterpret_cast<const unsigned *>(&h01), *reinterpret_cast<const unsigned *>(&h23));
        *reinterpret_cast<uint2 *>(lo + off) = make_uint2(*reinterpret_cast<const unsigned *>(&l01), *reinterpret_cast<const unsigned *>(&l23));
    }
}

// HWIO [3][3][ic][oc] float32 -> [tap][oc][ic] fp16, hi = fp16(w) and lo = fp16(w - hi)
__global__ void k_conv_prep_weights_h(const float *__restrict__ w, __half *__restrict__ whi, __half *__restrict__ wlo,
                                      int *__restrict__ big) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;        // index into [tap][oc][ic]
    if (i >= 9 * F * F) return;
    const int ic = i % F, oc = (i / F) % F, tap = i / (F * F);
    const float x = w[((size_t)tap * F + ic) * F + oc];
    if (!(fabsf(x) <= CONV_F16_MAX)) *big = 1;                  // this layer's weights do not fit fp16
    const __half h = __float2half_rn(x);
    whi[i] = h;
    wlo[i] = __float2half_rn(x - __half2float(h));
}

template <bool LAST>
__global__ void __launch_bounds__(CT_THREADS, 1)
k_conv64_h(const __grid_constant__ CtcMaps maps, const float *__restrict__ bias, float *__restrict__ out, int OH, int OW,
            int ntx, int ntiles, const int *__restrict__ in_big, const int *__restrict__ w_big, int run_if, int *__restrict__ big) {
    // *in_big | *w_big != 0: this layer's input (or its weights) does not fit fp16 -- the TF32 kernel (run_if = 1) does
    // the layer, the FP16 kernel (run_if = 0) leaves at once; and the other way round.  in_big == NULL: run.
    if (in_big && ((*in_big | *w_big) != 0) != (run_if != 0)) return;
    extern __shared__ __align__(1024) unsigned char ctc_raw[];
    ChSmem &sm = *reinterpret_cast<ChSmem *>(ctc_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(&sm.tmem_base)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        tc_mbar_init(&sm.bar_w, 1);
        tc_mbar_init(&sm.bar_zero, 1);
        for (int i = 0; i < CH_NRAW; i++) {
            tc_mbar_init(&sm.bar_row[i], 1);
            tc_mbar_init(&sm.bar_rawfree[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            tc_mbar_init(&sm.bar_rowdone[i], 1);
            tc_mbar_init(&sm.bar_full[i], 1);
            tc_mbar_init(&sm.bar_empty[i], 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = sm.tmem_base;
    // instruction descriptor: D = F32, A = B = F16, both K-major, M = 128; N (64, 128 or 192) is filled in per MMA
    const unsigned idesc_base = (1u << 4) | ((unsigned)(128 >> 4) << 24);
    constexpr int NROW = CT_ROWS + 2;           // input rows per tile

    // A CTA takes its tiles two at a time (one per accumulator buffer) and visits  block 0 of both, then block 1 of both:
    // the weights are reloaded once per tile on average, and EVERY tile accumulates block 0 before block 1 -- a pixel's
    // features do not depend on where its tile falls in some CTA's sequence, so a row band of an image (slab.py) gives
    // the bits of the whole image.  Row steps are numbered globally in that order.
    if (warp == 1) {
        // ================= TMA producer (one lane) =================
        if (lane == 0) {
            unsigned rr = 0;
            // the weights of both channel blocks, all nine taps, hi and lo (144 KB as fp16) stay for the whole kernel
            tc_mbar_expect_tx(&sm.bar_w, CH_W_BYTES);
            for (int kb = 0; kb < 2; kb++)
                for (int tap = 0; tap < 9; tap++) {
                    tc_tma_load_3d(sm.w[kb][0][tap % 3][2 - tap / 3], &maps.whi, kb * 32, 0, tap, &sm.bar_w);
                    tc_tma_load_3d(sm.w[kb][1][tap % 3][2 - tap / 3], &maps.wlo, kb * 32, 0, tap, &sm.bar_w);
                }
            for (int tile0 = blockIdx.x; tile0 < ntiles; tile0 += 2 * gridDim.x) {
              for (int kb = 0; kb < 2; kb++) {
                for (int mem = 0; mem < 2; mem++) {
                    const int tile = tile0 + mem * gridDim.x;
                    if (tile >= ntiles) break;
                    const int ty = tile / ntx, y0 = ty * CT_ROWS, x0 = (tile - ty * ntx) * CT_PIX;
                    for (int r = 0; r < NROW; r++, rr++) {
                        // the raw buffer was last used by step rr - CH_NRAW: free once that step's split has read it
                        const unsigned rs = rr % CH_NRAW;
                        if (rr >= CH_NRAW) {
                            tc_mbar_wait_sleep(&sm.bar_rawfree[rs], ((rr - CH_NRAW) / CH_NRAW) & 1);
                            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                        }
                        tc_mbar_expect_tx(&sm.bar_row[rs], CT_ROW_BYTES);
                        tc_tma_load_3d(sm.raw[rs], &maps.in, kb * 32, x0, y0 + r, &sm.bar_row[rs]);
                    }
                }
              }
            }
        }
    } else if (warp < 8) {
        // ================= operand split (warps 2-7) and MMA issue (warp 0, lane 0) =================
        unsigned rr = 0, pair = 0;
        for (int tile0 = blockIdx.x; tile0 < ntiles; tile0 += 2 * gridDim.x, pair++) {
          for (int kb = 0; kb < 2; kb++) {
            for (int mem = 0; mem < 2; mem++) {
                if (tile0 + mem * (int)gridDim.x >= ntiles) break;
                const unsigned abuf = mem, tt = 2 * pair + mem;
                const int bi = kb;
                for (int r = 0; r < NROW; r++, rr++) {
                    const unsigned rb = rr & 1;
                    const unsigned rs = rr % CH_NRAW;
                    if (warp >= 2) {
                        // the fp16 tiles of this step were last read by the MMAs of step rr - 2
                        if (rr >= 2) tc_mbar_wait(&sm.bar_rowdone[rb], ((rr - 2) >> 1) & 1);
                        tc_mbar_wait(&sm.bar_row[rs], (rr / CH_NRAW) & 1);
                        ch_split(sm.raw[rs], sm.a_hi[rb], sm.a_lo[rb], tid - 64, CT_NSPLIT);
                        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    }
                    tc_named_barrier(1, 32 + CT_NSPLIT);
                    if (tid == 0) {
                        tc_mbar_arrive(&sm.bar_rawfree[rs]);                       // every splitter is done with the raw row
                        if (rr == 0) tc_mbar_wait(&sm.bar_w, 0);
                        if (bi == 0 && r == 0 && tt >= 2) tc_mbar_wait(&sm.bar_empty[abuf], ((tt >> 1) - 1) & 1);
                        if (rr == 0) tc_mbar_wait(&sm.bar_zero, 0);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        const unsigned long long ah = ch_smem_desc(tc_smem_u32(sm.a_hi[rb])), al = ch_smem_desc(tc_smem_u32(sm.a_lo[rb]));
                        // Input row r feeds accumulator j = r - ky for every kernel row ky it can pair with: the weights
                        // of those kernel rows are stacked (ky descending = j ascending), so ONE MMA with N = 64, 128 or
                        // 192 updates the adjacent accumulators j = r - kymax .. r - kymin.  Accumulators start from the
                        // zeros the epilogue leaves in TMEM, so every MMA accumulates.
                        const int kymax = min(2, r), kymin = max(0, r - (CT_ROWS - 1));
                        const unsigned nacc = (unsigned)(kymax - kymin + 1);
                        const unsigned idesc_n = idesc_base | ((nacc * F >> 3) << 17);
                        const unsigned d_tmem = tmem_base + abuf * (CT_ROWS * F) + (unsigned)(r - kymax) * F;
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) {
                            const unsigned long long wh = ch_smem_desc(tc_smem_u32(sm.w[bi][0][kx][2 - kymax]));
                            const unsigned long long wl = ch_smem_desc(tc_smem_u32(sm.w[bi][1][kx][2 - kymax]));
#pragma unroll
                            for (int ks = 0; ks < 2; ks++) {                       // K = 16 halves = 32 bytes per MMA
                                const unsigned long long aoff = (unsigned long long)((kx * 64 + ks * 32) >> 4);
                                const unsigned long long woff = (unsigned long long)((ks * 32) >> 4);
                                ch_mma_f16(d_tmem, ah + aoff, wh + woff, idesc_n, 1);
                                ch_mma_f16(d_tmem, ah + aoff, wl + woff, idesc_n, 1);
                                ch_mma_f16(d_tmem, al + aoff, wh + woff, idesc_n, 1);
                            }
                        }
                        tc_mma_commit(&sm.bar_rowdone[rb]);
                        if (bi == 1 && r == NROW - 1) tc_mma_commit(&sm.bar_full[abuf]);
                    }
                }
            }
          }
        }
    } else {
        // ================= epilogue (warps 8-15): two warps per TMEM lane quarter, two output rows each =================
        const int q = warp & 3, half = (warp - 8) >> 2;
        const int m = 32 * q + lane;
        const float4 *b4 = reinterpret_cast<const float4 *>(bias);
        // MMAs only ever accumulate: this warp's share of both accumulator buffers (its lane quarter, its two rows)
        // starts as zeros and is zeroed again after every read
        auto zero_acc = [&](unsigned abuf, int j) {
            const unsigned taddr = tmem_base + ((unsigned)(32 * q) << 16) + abuf * (CT_ROWS * F) + j * F;
            tc_tmem_st32_zero(taddr);
            tc_tmem_st32_zero(taddr + 32);
        };
        for (unsigned ab = 0; ab < 2; ab++)
            for (int jj = 0; jj < CT_ROWS / 2; jj++) zero_acc(ab, half * (CT_ROWS / 2) + jj);
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        tc_named_barrier(2, 256);                       // all eight epilogue warps have zeroed their part ...
        if (tid == 256) tc_mbar_arrive(&sm.bar_zero);    // ... tell the MMA issuer
        unsigned tt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tt++) {
            const unsigned abuf = tt & 1;
            const int ty = tile / ntx, y0 = ty * CT_ROWS, x0 = (tile - ty * ntx) * CT_PIX;
            tc_mbar_wait(&sm.bar_full[abuf], (tt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
            for (int jj = 0; jj < CT_ROWS / 2; jj++) {
                const int j = half * (CT_ROWS / 2) + jj;
                const int y = y0 + j, x = x0 + m;
                unsigned v0[32], v1[32];
                const unsigned taddr = tmem_base + ((unsigned)(32 * q) << 16) + abuf * (CT_ROWS * F) + j * F;
                tc_tmem_ld32(taddr, v0);
                tc_tmem_ld32(taddr + 32, v1);
                zero_acc(abuf, j);
                float ss = 0.f;
#pragma unroll
                for (int o = 0; o < 32; o += 4) {
                    const float4 ba = __ldg(b4 + (o >> 2)), bb = __ldg(b4 + 8 + (o >> 2));
                    float t;
                    t = __uint_as_float(v0[o]) + ba.x;     v0[o] = __float_as_uint(t);     ss = fmaf(t, t, ss);
                    t = __uint_as_float(v0[o + 1]) + ba.y; v0[o + 1] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v0[o + 2]) + ba.z; v0[o + 2] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v0[o + 3]) + ba.w; v0[o + 3] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v1[o]) + bb.x;     v1[o] = __float_as_uint(t);     ss = fmaf(t, t, ss);
                    t = __uint_as_float(v1[o + 1]) + bb.y; v1[o + 1] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v1[o + 2]) + bb.z; v1[o + 2] = __float_as_uint(t); ss = fmaf(t, t, ss);
                    t = __uint_as_float(v1[o + 3]) + bb.w; v1[o + 3] = __float_as_uint(t); ss = fmaf(t, t, ss);
                }
                const float inv = LAST ? 1.0f / sqrtf(fmaxf(ss, 1e-12f)) : 1.0f;   // model.py:64
                if (y < OH && m < CT_PIX && x < OW) {
                    unsigned mxb = 0;                                               // largest activation written, as bits
                    float *dst = out + ((size_t)y * OW + x) * F;
#pragma unroll
                    for (int o = 0; o < 32; o += 8) {
                        float e[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const float t = __uint_as_float(v0[o + k]);
                            e[k] = LAST ? t * inv : fmaxf(t, 0.f);                  // model.py:120-123
                            if (!LAST) mxb = max(mxb, __float_as_uint(e[k]));   // (non-negative after the ReLU)
                        }
                        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + o), "f"(e[0]), "f"(e[1]),
                                     "f"(e[2]), "f"(e[3]), "f"(e[4]), "f"(e[5]), "f"(e[6]), "f"(e[7])
                                     : "memory");
                    }
#pragma unroll
                    for (int o = 0; o < 32; o += 8) {
                        float e[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const float t = __uint_as_float(v1[o + k]);
                            e[k] = LAST ? t * inv : fmaxf(t, 0.f);
                            if (!LAST) mxb = max(mxb, __float_as_uint(e[k]));   // (non-negative after the ReLU)
                        }
                        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 32 + o), "f"(e[0]), "f"(e[1]),
                                     "f"(e[2]), "f"(e[3]), "f"(e[4]), "f"(e[5]), "f"(e[6]), "f"(e[7])
                                     : "memory");
                    }
                    if (!LAST && big && mxb > __float_as_uint(CONV_F16_MAX)) *big = 1;   // the next layer must not use fp16 operands
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            tc_mbar_arrive(&sm.bar_empty[abuf]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

// Single-layer network (num_layers == 1): normalise the conv1 output in place.
__global__ void k_l2norm64(float *__restrict__ x, long long P) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long px = t >> 4;
    int c4 = (int)(t & 15);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px < P) v = reinterpret_cast<float4 *>(x)[px * 16 + c4];
    float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    ss += __shfl_xor_sync(0xffffffffu, ss, 8);
    const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
    if (px < P) reinterpret_cast<float4 *>(x)[px * 16 + c4] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
}

}  // namespace mccnn

using namespace mccnn;

extern "C" {

// Layers 2..n run with FP16 hi/lo operands (k_conv64_h) unless the layer's input or weights leave fp16's range, in which
// case the TF32 kernel (k_conv64_tc) does that layer: both kernels are launched for every layer and exactly one of them
// works, decided on the device by two flags -- gate[0]: the producer of the layer's input saw an activation beyond
// CONV_F16_MAX (set by its epilogue), gate[1]: the weight split did.  No host read-back.  MCCNN_CONV_TF32=1 forces the
// TF32 kernel everywhere (only it is launched).
static bool conv_tf32() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("MCCNN_CONV_TF32"); v = (e && atoi(e) != 0) ? 1 : 0; }
    return v == 1;
}

// prepared weights of layer l >= 1 (floats): [tf32 hi | tf32 lo | fp16 hi, fp16 lo (halves) | flag word + padding]
constexpr size_t CONV_PREP_FLOATS = 3 * 9 * F * F + 16;
static size_t conv_weights_floats(int num_layers) { return (size_t)(num_layers > 1 ? num_layers - 1 : 0) * CONV_PREP_FLOATS; }
constexpr size_t CONV_FLAG_INTS = 32;          // per call: flag of every layer's input (in the activation scratch)

size_t mccnn_features_scratch_bytes(int H, int W, int pad, int num_layers) {
    if (H < 1 || W < 1 || num_layers < 1 || pad < 0 || H + 2 * pad < 3 || W + 2 * pad < 3) return 0;
    size_t oh = (size_t)H + 2 * pad - 2, ow = (size_t)W + 2 * pad - 2;
    // two ping-pong activation maps + the per-layer range flags + the pre-split weights of layers 2..n
    return (2 * oh * ow * F + CONV_FLAG_INTS + conv_weights_floats(num_layers)) * sizeof(float);
}

size_t mccnn_features_weights_bytes(int num_layers) {
    return num_layers >= 1 ? conv_weights_floats(num_layers) * sizeof(float) : 0;
}

static int prep_layer(const float *w, float *base, cudaStream_t s) {
    float *whi = base, *wlo = base + 9 * F * F;
    __half *hhi = reinterpret_cast<__half *>(base + 2 * 9 * F * F);
    int *wflag = reinterpret_cast<int *>(base + 3 * 9 * F * F);
    MCCNN_CUDA(cudaMemsetAsync(wflag, 0, 16 * sizeof(float), s));
    k_conv_prep_weights<<<cdiv(9 * F * F, 256), 256, 0, s>>>(w, whi, wlo);
    MCCNN_LAUNCHED("conv_prep_weights");
    k_conv_prep_weights_h<<<cdiv(9 * F * F, 256), 256, 0, s>>>(w, hhi, hhi + 9 * F * F, wflag);
    MCCNN_LAUNCHED("conv_prep_weights_h");
    return MCCNN_OK;
}

int mccnn_features_prepare(int num_layers, const float *const *weights_host, void *prepared, void *stream) {
    MCCNN_REQUIRE(weights_host && num_layers >= 1 && num_layers <= 16, "features_prepare: bad arguments");
    MCCNN_REQUIRE(num_layers == 1 || (prepared && ((uintptr_t)prepared & 31) == 0), "features_prepare: prepared must be 32-byte aligned");
    for (int l = 1; l < num_layers; l++) {
        int rc = prep_layer(weights_host[l], (float *)prepared + (size_t)(l - 1) * CONV_PREP_FLOATS, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return MCCNN_OK;
}

static int features_impl(const float *img, int H, int W, int pad, int num_layers, const float *const *weights_host,
                         const float *const *biases_host, const float *prepared, float *out, void *scratch, void *stream) {
    MCCNN_REQUIRE(img && weights_host && biases_host && out, "features: null pointer");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && pad >= 0 && num_layers >= 1 && num_layers <= 16,
                  "features: bad shape H=%d W=%d pad=%d layers=%d", H, W, pad, num_layers);
    MCCNN_REQUIRE(H + 2 * pad - 2 * num_layers >= 1 && W + 2 * pad - 2 * num_layers >= 1,
                  "features: image %dx%d (pad %d) too small for %d VALID 3x3 layers", H, W, pad, num_layers);
    MCCNN_REQUIRE(num_layers == 1 || scratch, "features: scratch required");
    MCCNN_REQUIRE(((uintptr_t)out & 31) == 0 && ((uintptr_t)scratch & 31) == 0, "features: out and scratch must be 32-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    int oh = H + 2 * pad - 2, ow = W + 2 * pad - 2;
    float *buf[2];
    buf[0] = (float *)scratch;
    buf[1] = buf[0] ? buf[0] + (size_t)oh * ow * F : nullptr;
    int *flags = buf[0] ? reinterpret_cast<int *>(buf[1] + (size_t)oh * ow * F) : nullptr;   // flags[l]: input of layer l too big
    // the pre-split weights of layers 2..n: the caller's (mccnn_features_prepare, once per network), else made here
    float *wsplit = prepared ? const_cast<float *>(prepared) : (buf[0] ? reinterpret_cast<float *>(flags + CONV_FLAG_INTS) : nullptr);
    float *dst = (num_layers == 1) ? out : buf[0];
    if (flags) MCCNN_CUDA(cudaMemsetAsync(flags, 0, CONV_FLAG_INTS * sizeof(int), s));
    {
        MCCNN_REQUIRE(oh <= 65535, "features: image too tall (%d rows)", oh);
        k_conv1<<<dim3(cdiv(ow, 16 * C1_PX), oh), 256, 0, s>>>(img, weights_host[0], biases_host[0], dst, H, W, pad, oh, ow,
                                                                        num_layers > 1, num_layers > 1 ? flags + 1 : nullptr);
        MCCNN_LAUNCHED("conv1");
    }
    if (num_layers == 1) {
        long long threads = (long long)oh * ow * 16;
        k_l2norm64<<<cdiv(threads, 256), 256, 0, s>>>(out, (long long)oh * ow);
        MCCNN_LAUNCHED("l2norm64");
        return MCCNN_OK;
    }
    // per device, asked on every call: the opt-in to > 48 KB of dynamic shared memory belongs to the current device's
    // context and the persistent grid to its SM count (no process-wide caches: one process may drive several GPUs)
    int dev = 0, num_sms = 0;
    MCCNN_CUDA(cudaGetDevice(&dev));
    MCCNN_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    MCCNN_CUDA(cudaFuncSetAttribute(k_conv64_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtcSmem)));
    MCCNN_CUDA(cudaFuncSetAttribute(k_conv64_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtcSmem)));
    MCCNN_CUDA(cudaFuncSetAttribute(k_conv64_h<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChSmem)));
    MCCNN_CUDA(cudaFuncSetAttribute(k_conv64_h<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChSmem)));
    const bool force_tf32 = conv_tf32();
    const float *src = dst;
    int ih = oh, iw = ow;
    for (int l = 1; l < num_layers; l++) {
        oh = ih - 2; ow = iw - 2;
        const bool last = (l == num_layers - 1);
        float *d = last ? out : buf[l & 1];
        float *base = wsplit + (size_t)(l - 1) * CONV_PREP_FLOATS;
        float *whi = base, *wlo = base + 9 * F * F;
        __half *hhi = reinterpret_cast<__half *>(base + 2 * 9 * F * F), *hlo = hhi + 9 * F * F;
        int *wflag = reinterpret_cast<int *>(base + 3 * 9 * F * F);
        if (!prepared) {
            int rc = prep_layer(weights_host[l], base, s);
            if (rc) return rc;
        }
        const int *in_big = flags + l;                            // set by the producer of this layer's input
        int *big = last ? nullptr : flags + l + 1;
        const int ntx = cdiv(ow, CT_PIX), nty = cdiv(oh, CT_ROWS);
        const int ntiles = ntx * nty;
        const int grid = ntiles < num_sms ? ntiles : num_sms;
        CtcMaps maps;
        int rc = tc_encode_map_3d(maps.in, src, F, iw, ih, 32, 128, true, "features");
        if (rc) return rc;
        if (!force_tf32) {
            rc = tc_encode_map_3d_f16_sw64(maps.whi, hhi, F, F, 9, 32, F, "features");
            if (rc) return rc;
            rc = tc_encode_map_3d_f16_sw64(maps.wlo, hlo, F, F, 9, 32, F, "features");
            if (rc) return rc;
            if (last) k_conv64_h<true><<<grid, CT_THREADS, sizeof(ChSmem), s>>>(maps, biases_host[l], d, oh, ow, ntx, ntiles, in_big, wflag, 0, big);
            else k_conv64_h<false><<<grid, CT_THREADS, sizeof(ChSmem), s>>>(maps, biases_host[l], d, oh, ow, ntx, ntiles, in_big, wflag, 0, big);
            MCCNN_LAUNCHED("conv64_h");
        }
        rc = tc_encode_map_3d(maps.whi, whi, F, F, 9, 32, F, true, "features");
        if (rc) return rc;
        rc = tc_encode_map_3d(maps.wlo, wlo, F, F, 9, 32, F, true, "features");
        if (rc) return rc;
        const int *tgate = force_tf32 ? nullptr : in_big;
        if (last) k_conv64_tc<true><<<grid, CT_THREADS, sizeof(CtcSmem), s>>>(maps, biases_host[l], d, oh, ow, ntx, ntiles, tgate, wflag, 1, big);
        else k_conv64_tc<false><<<grid, CT_THREADS, sizeof(CtcSmem), s>>>(maps, biases_host[l], d, oh, ow, ntx, ntiles, tgate, wflag, 1, big);
        MCCNN_LAUNCHED("conv64_tc");
        src = d; ih = oh; iw = ow;
    }
    return MCCNN_OK;
}

int mccnn_features(const float *img, int H, int W, int pad, int num_layers, const float *const *weights_host,
                   const float *const *biases_host, float *out, void *scratch, void *stream) {
    return features_impl(img, H, W, pad, num_layers, weights_host, biases_host, nullptr, out, scratch, stream);
}

int mccnn_features_prepared(const float *img, int H, int W, int pad, int num_layers, const float *const *weights_host,
                            const float *const *biases_host, const void *prepared, float *out, void *scratch, void *stream) {
    MCCNN_REQUIRE(num_layers == 1 || (prepared && ((uintptr_t)prepared & 31) == 0), "features_prepared: prepared weights required");
    return features_impl(img, H, W, pad, num_layers, weights_host, biases_host, (const float *)prepared, out, scratch, stream);
}

}  // extern "C"
