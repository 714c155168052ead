// Separable cross-based aggregation round (fast mode of mccnn_cbca).
//
// The support region of pf:571-657 is "for every spine row h' in [h-up, h+down]: the horizontal
// arm of (h', w)", so one round is exactly  out(h,w) = ( sum_{h'} Hs(h',w) ) / |U(h,w)|  with
// Hs(h',w) = sum_{w' in arm(h',w)} in(h',w').  The flat kernel walks all |U| <= 729 cells per output;
// this one computes each row sum once per tile and re-uses it for every pixel of the column:
// <= 27 + 27 additions per output instead of <= 729.  Only the association of the float32 sum
// differs from the reference ((row sums) summed, instead of one running sum): ~1e-7 relative.
//
// One CTA = a SEP_TH x SEP_TW pixel tile x SEP_GC disparity granules (SEP_GC*4 disparities).
//   phase A: row sums of the tile's rows plus the halo rows actually needed (the largest up/down
//            arm inside the tile, found first -- natural images need 0-2 halo rows, flat ones 13)
//            from global memory (neighbouring pixels share L1 lines) into shared memory;
//   phase B: column sums over shared memory, division by |U|, one coalesced 16 B store per item.
#pragma once
#include "common.cuh"

namespace mccnn {

constexpr int SEP_TH = 16, SEP_TW = 8, SEP_GC = 8, SEP_THREADS = 256;
constexpr int SEP_MAXROWS = SEP_TH + 2 * 13;                 // distance_threshold <= 14 in this mode

__global__ void __launch_bounds__(SEP_THREADS) k_cbca_round_sep(const float4 *__restrict__ in, float4 *__restrict__ out,
                                                                const uchar4 *__restrict__ arms,
                                                                const int32_t *__restrict__ count, int G, int H, int W) {
    __shared__ float4 Hs[SEP_MAXROWS][SEP_TW][SEP_GC];       // 42 * 8 * 8 * 16 B = 43 KB
    __shared__ int s_up, s_down;
    const int tid = threadIdx.x;
    const int w0 = blockIdx.x * SEP_TW, h0 = blockIdx.y * SEP_TH, g0 = blockIdx.z * SEP_GC;
    if (tid == 0) { s_up = 0; s_down = 0; }
    __syncthreads();
    // largest vertical arms in the tile -> halo rows needed
    if (tid < SEP_TH * SEP_TW) {
        const int h = h0 + tid / SEP_TW, w = w0 + tid % SEP_TW;
        if (h < H && w < W) {
            const uchar4 a = arms[(size_t)h * W + w];
            const int need_up = (int)a.x - (h - h0);                     // rows above the tile
            const int need_dn = (int)a.y - (h0 + SEP_TH - 1 - h);        // rows below the tile
            if (need_up > 0) atomicMax(&s_up, need_up);
            if (need_dn > 0) atomicMax(&s_down, need_dn);
        }
    }
    __syncthreads();
    const int rbeg = max(h0 - s_up, 0);
    const int rend = min(h0 + SEP_TH + s_down, H);                       // exclusive
    const int nrows = rend - rbeg;
    const int gc = tid % SEP_GC, px = (tid / SEP_GC) % SEP_TW, rsub = tid / (SEP_GC * SEP_TW);
    constexpr int RSTEP = SEP_THREADS / (SEP_GC * SEP_TW);               // rows per sweep (4)
    const int g = g0 + gc, w = w0 + px;
    const bool live = (g < G) && (w < W);

    // phase A: horizontal arm sums, order w, w-1, .., w-left, w+1, .., w+right (pf:645-650)
    for (int r = rsub; r < nrows; r += RSTEP) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            const int hh = rbeg + r;
            const size_t rowp = (size_t)hh * W;
            const uchar4 s = arms[rowp + w];
            const float4 *src = in + (rowp + w) * G + g;
            for (int j = 0; j <= s.z; j++) {
                const float4 v = src[-(ptrdiff_t)j * G];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            for (int j = 1; j <= s.w; j++) {
                const float4 v = src[(ptrdiff_t)j * G];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        Hs[r][px][gc] = acc;
    }
    __syncthreads();

    // phase B: vertical sums in spine order h, h-1, .., h-up, h+1, .., h+down (pf:640-644), then / |U|
    for (int r = rsub; r < SEP_TH; r += RSTEP) {
        const int h = h0 + r;
        if (!live || h >= H) continue;
        const size_t p = (size_t)h * W + w;
        const uchar4 a = arms[p];
        const int base = h - rbeg;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k <= a.x; k++) {
            const float4 v = Hs[base - k][px][gc];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        for (int k = 1; k <= a.y; k++) {
            const float4 v = Hs[base + k][px][gc];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const float n = (float)count[p];
        out[p * G + g] = make_float4(acc.x / n, acc.y / n, acc.z / n, acc.w / n);
    }
}

}  // namespace mccnn
