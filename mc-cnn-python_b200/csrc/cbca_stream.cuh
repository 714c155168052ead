// Separable cross-based aggregation as two streaming passes (default fast mode of mccnn_cbca).
//
// One round is  out(h,w) = ( sum_{h' in spine(h,w)} Hs(h',w) ) / |U(h,w)|,  Hs(h',w) = sum_{w' in arm(h',w)} in(h',w')
// (pf:640-650): k_cbca_rows writes Hs, k_cbca_cols adds it along the spine and divides.  Both are pure
// gathers with no shared memory and no barriers: the lanes of a warp are consecutive disparity granules of
// one pixel (two pixels when a warp straddles a pixel boundary), so every load and store is a contiguous run
// and an arm walk is warp uniform; neighbouring pixels are served by L1/L2 (a CTA covers an 8x4 pixel patch
// per 64 disparities).  HBM traffic is 16 B per cell per round -- twice the fused minimum -- but the passes
// run near copy speed, which the fused shared-memory kernel (cbca_tile.cuh) does not: its short
// data-dependent phases are dominated by barrier waits (profiles/r1_cbca_round_tile.md).
// Summation order inside a row and along the spine is the reference's; only the association
// (row sums first) differs: ~1e-7 relative.
#pragma once
#include "common.cuh"

namespace mccnn {

constexpr int CS_PH = 8, CS_PW = 4, CS_GC = 16, CS_THREADS = 128;   // patch 8x4 pixels x 16 granules per CTA (best of a sweep)

__device__ __forceinline__ void cs_add(float4 &acc, const float4 v) {
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
}

constexpr int CS_ITEMS = (CS_PH * CS_PW * CS_GC) / CS_THREADS;    // (pixel, granule) items per thread: 4

__device__ __forceinline__ void cs_cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}

// The centre cell and the two nearest neighbours of every item travel global -> shared with cp.async: bytes in
// flight that cost no registers (these kernels are bound by memory-level parallelism, and most arms are 0 or 1 long,
// so the common walk needs nothing else).  A thread reads back only its own slots: no barrier.
__global__ void __launch_bounds__(CS_THREADS) k_cbca_rows(const float4 *__restrict__ in, float4 *__restrict__ hs,
                                                          const uchar4 *__restrict__ arms, int G, int H, int W) {
    __shared__ float4 stage[3][CS_ITEMS][CS_THREADS];            // [centre, left, right]: 24 KB
    const int gi = threadIdx.x % CS_GC, g = blockIdx.z * CS_GC + gi;
    if (g >= G) return;
    int p[CS_ITEMS];
    uchar4 a[CS_ITEMS];
#pragma unroll
    for (int s = 0; s < CS_ITEMS; s++) {
        const int pi = s * (CS_THREADS / CS_GC) + threadIdx.x / CS_GC;
        const int h = blockIdx.y * CS_PH + pi / CS_PW, w = blockIdx.x * CS_PW + pi % CS_PW;
        p[s] = (h < H && w < W) ? h * W + w : -1;
        if (p[s] >= 0) {
            const float4 *c = in + (size_t)p[s] * G + g;
            cs_cp_async16(&stage[0][s][threadIdx.x], c);
            if (w > 0) cs_cp_async16(&stage[1][s][threadIdx.x], c - G);
            if (w + 1 < W) cs_cp_async16(&stage[2][s][threadIdx.x], c + G);
            a[s] = arms[p[s]];
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#pragma unroll
    for (int s = 0; s < CS_ITEMS; s++) {
        if (p[s] < 0) continue;
        const float4 *c = in + (size_t)p[s] * G + g;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        cs_add(acc, stage[0][s][threadIdx.x]);                                   // w, w-1, .., w-left (pf:645-650)
        if (a[s].z >= 1) cs_add(acc, stage[1][s][threadIdx.x]);
        for (int j = 2; j <= a[s].z; j++) cs_add(acc, c[-(ptrdiff_t)j * G]);
        if (a[s].w >= 1) cs_add(acc, stage[2][s][threadIdx.x]);                  // w+1, .., w+right
        for (int j = 2; j <= a[s].w; j++) cs_add(acc, c[(ptrdiff_t)j * G]);
        hs[(size_t)p[s] * G + g] = acc;
    }
}

// a / n for n = |U| (an integer <= 729), y = RN(1/n): q = RN(a*y); r = a - n*q (exact, FMA); RN(q + r*y):
// correctly rounded (Markstein), i.e. the IEEE division of pf:161; odd magnitudes take the plain division.
__device__ __forceinline__ float cs_div1(float a, float n, float y) {
    const float q = a * y;
    const float r = fmaf(-n, q, a);
    return fmaf(r, y, q);
}

__global__ void __launch_bounds__(CS_THREADS) k_cbca_cols(const float4 *__restrict__ hs, float4 *__restrict__ out,
                                                          const uchar4 *__restrict__ arms, const int32_t *__restrict__ count,
                                                          int G, int H, int W) {
    const int gi = threadIdx.x % CS_GC, g = blockIdx.z * CS_GC + gi;
    if (g >= G) return;
    const ptrdiff_t rs = (ptrdiff_t)W * G;
    size_t p[CS_ITEMS];
    bool ok[CS_ITEMS];
    uchar4 a[CS_ITEMS];
    float4 c0[CS_ITEMS];
    float n[CS_ITEMS];
#pragma unroll
    for (int s = 0; s < CS_ITEMS; s++) {
        const int pi = s * (CS_THREADS / CS_GC) + threadIdx.x / CS_GC;
        const int h = blockIdx.y * CS_PH + pi / CS_PW, w = blockIdx.x * CS_PW + pi % CS_PW;
        ok[s] = h < H && w < W;
        p[s] = ok[s] ? (size_t)h * W + w : 0;
        a[s] = arms[p[s]];
        n[s] = (float)count[p[s]];
        c0[s] = hs[p[s] * G + g];
    }
#pragma unroll
    for (int s = 0; s < CS_ITEMS; s++) {
        if (!ok[s]) continue;
        const float4 *c = hs + p[s] * G + g;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        cs_add(acc, c0[s]);                                                      // h, h-1, .., h-up (pf:640-644)
        for (int k = 1; k <= a[s].x; k++) cs_add(acc, c[-k * rs]);
        for (int k = 1; k <= a[s].y; k++) cs_add(acc, c[k * rs]);                // h+1, .., h+down
        const float hi = fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)));
        const float lo = fminf(fminf(fabsf(acc.x), fabsf(acc.y)), fminf(fabsf(acc.z), fabsf(acc.w)));
        float4 r;
        if (hi < 1e30f && lo > 1e-30f) {
            const float y = 1.0f / n[s];
            r = make_float4(cs_div1(acc.x, n[s], y), cs_div1(acc.y, n[s], y), cs_div1(acc.z, n[s], y), cs_div1(acc.w, n[s], y));
        } else {
            r = make_float4(acc.x / n[s], acc.y / n[s], acc.z / n[s], acc.w / n[s]);   // pf:161
        }
        out[p[s] * G + g] = r;
    }
}

}  // namespace mccnn
