// Separable cross-based aggregation round with the row sums kept in L2 (mode MCCNN_CBCA_SEPARABLE_L2).
//
// Same two passes, same arithmetic and order as cbca_stream.cuh (so the results are bit-identical), but ONE
// persistent kernel per round runs both, software-pipelined over bands of 8 image rows: the row-sum pass of band
// b + LAG is scheduled together with the column pass of band b, and the row sums live in a ring of 64 rows
// (50 MB at 1024 x 1024 x 192, the first rows of the caller's scratch volume) that is overwritten in place every
// 8 bands -- its lines stay dirty in the 126 MB L2 and are (mostly) never written back or fetched.  HBM traffic
// drops from 16 to about 8 B per cell per round; the column pass reads L2, not HBM.
//
// (Ring rows are reused: a row pass first checks that the column patches which read the rows it overwrites are done.)
// Scheduling: CTAs take tickets from a global counter; a ticket is 8 consecutive 8x4-pixel x 16-granule patches of one pass.
// Tickets are ordered R0 .. R(LAG-1), R(LAG), C0, R(LAG+1), C1, ...: a column patch of band b needs the row sums
// of bands b-2 .. b+2 (arms reach 13 rows), which were handed out at least two bands of tickets earlier, so its
// wait on their completion counters is almost always already satisfied, and never deadlocks (a waiting CTA only
// waits for tickets held by CTAs that are running and wait for nothing later than themselves).
// Row sums are read with ld.global.cg: the ring's addresses are reused, L1 could hold a previous band.
#pragma once
#include "common.cuh"
#include "cbca_stream.cuh"

namespace mccnn {

constexpr int CF_RING_ROWS = 64;              // power of two, >= 8 * (LAG + 3)
constexpr int CF_LAG = 4;                     // bands between a band's row pass and its column pass
constexpr int CF_REACH = 2;                   // bands a 13-row arm can reach into (8-row bands)

struct CfSched { int nbx, nz, nb, lag, items_band, chunk, chunks_band, total; };   // chunk patches per ticket; total = tickets per round

__device__ __forceinline__ unsigned cf_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(CS_THREADS, 8) k_cbca_fused(const float4 *__restrict__ in, float4 *hs,
                                                           float4 *__restrict__ out, const uchar4 *__restrict__ arms,
                                                           const int32_t *__restrict__ count, int G, int H, int W,
                                                           const CfSched sc, unsigned *__restrict__ counters) {
    __shared__ float4 stage[3][CS_ITEMS][CS_THREADS];            // [centre, left, right] of the row pass: 24 KB
    __shared__ int s_ticket;
    unsigned *ticket = counters, *rows_done = counters + 1, *cols_done = counters + 1 + sc.nb;
    const int gi = threadIdx.x % CS_GC;
    const ptrdiff_t rs = (ptrdiff_t)W * G;                       // one image row, in granules
    int next = 0;
    if (threadIdx.x == 0) next = (int)atomicAdd(ticket, 1u);
    for (;;) {
        if (threadIdx.x == 0) s_ticket = next;
        __syncthreads();
        const int t = s_ticket;
        __syncthreads();
        if (t >= sc.total) break;
        if (threadIdx.x == 0) next = (int)atomicAdd(ticket, 1u);     // the next ticket travels while this one is worked on
        // ---- decode the ticket: blocks of items_band tickets, R0..R(lag-1), then (R, C) pairs, then the last C's
        const int q = t / sc.chunks_band, chunk = t - q * sc.chunks_band;
        int band;
        bool rows_pass;
        if (q < sc.lag) { rows_pass = true; band = q; }
        else if (q < sc.lag + 2 * (sc.nb - sc.lag)) {
            const int u = (q - sc.lag) >> 1;
            rows_pass = ((q - sc.lag) & 1) == 0;
            band = rows_pass ? sc.lag + u : u;
        } else { rows_pass = false; band = (sc.nb - sc.lag) + (q - sc.lag - 2 * (sc.nb - sc.lag)); }
        const int idx0 = chunk * sc.chunk, idx1 = min(idx0 + sc.chunk, sc.items_band);

        if (rows_pass) {
            // this band's ring rows were band - 8's: every column patch that reads them must be finished (it was
            // handed out six blocks of tickets ago: the wait is a formality, but nothing else bounds a slow CTA)
            if (threadIdx.x == 0 && band >= CF_RING_ROWS / CS_PH) {
                const int old = band - CF_RING_ROWS / CS_PH;
                const int lo = max(0, old - CF_REACH), hi = min(sc.nb - 1, old + CF_REACH);
                for (int b = lo; b <= hi; b++)
                    while (cf_ld_acquire(cols_done + b) < (unsigned)sc.chunks_band) __nanosleep(200);
            }
            __syncthreads();
            for (int idx = idx0; idx < idx1; idx++) {
            const int bx = idx % sc.nbx, bz = idx / sc.nbx;
            const int g = bz * CS_GC + gi;
            const bool g_ok = g < G;
            int p[CS_ITEMS];
            uchar4 a[CS_ITEMS];
#pragma unroll
            for (int s = 0; s < CS_ITEMS; s++) {
                const int pi = s * (CS_THREADS / CS_GC) + threadIdx.x / CS_GC;
                const int h = band * CS_PH + pi / CS_PW, w = bx * CS_PW + pi % CS_PW;
                p[s] = (g_ok && h < H && w < W) ? h * W + w : -1;
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
                cs_add(acc, stage[0][s][threadIdx.x]);                               // w, w-1, .., w-left (pf:645-650)
                if (a[s].z >= 1) cs_add(acc, stage[1][s][threadIdx.x]);
                for (int j = 2; j <= a[s].z; j++) cs_add(acc, c[-(ptrdiff_t)j * G]);
                if (a[s].w >= 1) cs_add(acc, stage[2][s][threadIdx.x]);              // w+1, .., w+right
                for (int j = 2; j <= a[s].w; j++) cs_add(acc, c[(ptrdiff_t)j * G]);
                const int h = p[s] / W, w = p[s] - h * W;
                hs[(size_t)(h & (CF_RING_ROWS - 1)) * rs + (size_t)w * G + g] = acc;
            }
            }
            __threadfence();                                                         // row sums visible before the count
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(rows_done + band, 1u);
        } else {
            if (threadIdx.x == 0) {
                const int lo = max(0, band - CF_REACH), hi = min(sc.nb - 1, band + CF_REACH);
                for (int b = lo; b <= hi; b++)
                    while (cf_ld_acquire(rows_done + b) < (unsigned)sc.chunks_band) __nanosleep(200);
            }
            __syncthreads();
            for (int idx = idx0; idx < idx1; idx++) {
            const int bx = idx % sc.nbx, bz = idx / sc.nbx;
            const int g = bz * CS_GC + gi;
            const bool g_ok = g < G;
            size_t p[CS_ITEMS];
            int hh[CS_ITEMS];
            bool ok[CS_ITEMS];
            uchar4 a[CS_ITEMS];
            float4 c0[CS_ITEMS];
            float n[CS_ITEMS];
#pragma unroll
            for (int s = 0; s < CS_ITEMS; s++) {
                const int pi = s * (CS_THREADS / CS_GC) + threadIdx.x / CS_GC;
                const int h = band * CS_PH + pi / CS_PW, w = bx * CS_PW + pi % CS_PW;
                ok[s] = g_ok && h < H && w < W;
                hh[s] = ok[s] ? h : 0;
                p[s] = ok[s] ? (size_t)h * W + w : 0;
                a[s] = arms[p[s]];
                n[s] = (float)count[p[s]];
                c0[s] = ok[s] ? __ldcg(hs + (size_t)(hh[s] & (CF_RING_ROWS - 1)) * rs + (p[s] - (size_t)hh[s] * W) * G + g)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int s = 0; s < CS_ITEMS; s++) {
                if (!ok[s]) continue;
                const float4 *col = hs + (p[s] - (size_t)hh[s] * W) * G + g;           // this column of the ring
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                cs_add(acc, c0[s]);                                                  // h, h-1, .., h-up (pf:640-644)
                for (int k = 1; k <= a[s].x; k++) cs_add(acc, __ldcg(col + (size_t)((hh[s] - k) & (CF_RING_ROWS - 1)) * rs));
                for (int k = 1; k <= a[s].y; k++) cs_add(acc, __ldcg(col + (size_t)((hh[s] + k) & (CF_RING_ROWS - 1)) * rs));
                const float hi = fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)));
                const float lo = fminf(fminf(fabsf(acc.x), fabsf(acc.y)), fminf(fabsf(acc.z), fabsf(acc.w)));
                float4 r;
                if (hi < 1e30f && lo > 1e-30f) {
                    const float y = 1.0f / n[s];
                    r = make_float4(cs_div1(acc.x, n[s], y), cs_div1(acc.y, n[s], y), cs_div1(acc.z, n[s], y), cs_div1(acc.w, n[s], y));
                } else {
                    r = make_float4(acc.x / n[s], acc.y / n[s], acc.z / n[s], acc.w / n[s]);   // pf:161
                }
                __stcs(out + p[s] * G + g, r);
            }
            }
            __syncthreads();                                                         // (all ring reads of this patch done)
            if (threadIdx.x == 0) atomicAdd(cols_done + band, 1u);
        }
    }
}

}  // namespace mccnn
