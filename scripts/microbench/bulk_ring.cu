// Microbenchmark: how fast can ONE producer warp per CTA feed a shared-memory ring with the access pattern of the
// chained CBCA kernel (per image-row segment: for each of ~33 pixels, 256-byte runs [16 granules] of rows h and h-1 of
// an [H][W][G] float4 volume)?  Three producers: cp.async.bulk (TMA, one 256 B op per lane), cp.async 16 B
// (LDGSTS, one entry per half warp), and bulk copies of a whole pixel (G*16 B per op).  Consumers only recycle stages.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_ring bulk_ring.cu && ./bulk_ring
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src)); }
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long *b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }

constexpr int NS = 3, S = 32, PX = 33, ROWS = 2;

// MODE 0: bulk 256 B per (pixel,row); MODE 1: LDGSTS 16 B x 16 lanes per (pixel,row); MODE 2: bulk G*16 B per (pixel,row), all granules
template <int MODE>
__global__ void __launch_bounds__(64) k_ring(const float4 *__restrict__ vol, int G, int H, int W, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long full[NS], empty[NS];
    const int nz = MODE == 2 ? 1 : G / 16, gc = MODE == 2 ? G : 16;
    const int entry_bytes = gc * 16, stage_bytes = PX * ROWS * entry_bytes;
    const int nsx = W / S, nseg = H * nsx * nz;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; i++) { mbar_init(&full[i], MODE == 1 ? 32 : 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int it = 0;
    if (warp == 0) {
        for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x, it++) {
            const int st = it % NS, ph = (it / NS) & 1;
            const int gz = seg % nz, sx = (seg / nz) % nsx, h = seg / (nz * nsx);
            unsigned char *stage = smem + (size_t)st * stage_bytes;
            mbar_wait(&empty[st], ph ^ 1);
            if (MODE != 1) {
                if (lane == 0) mbar_expect_tx(&full[st], stage_bytes);
                __syncwarp();
                for (int p = lane; p < PX; p += 32) {
                    const int x = min(sx * S + p, W - 1);
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        const int hh = max(h - r, 0);
                        bulk_g2s(stage + (size_t)(p * ROWS + r) * entry_bytes, vol + ((size_t)hh * W + x) * G + gz * 16, entry_bytes, &full[st]);
                    }
                }
            } else {
                const int gi = lane % 16, half = lane / 16;
                for (int e = half; e < PX * ROWS; e += 2) {
                    const int p = e / ROWS, r = e % ROWS;
                    const int x = min(sx * S + p, W - 1), hh = max(h - r, 0);
                    cp_async16(stage + (size_t)e * 256 + gi * 16, vol + ((size_t)hh * W + x) * G + gz * 16 + gi);
                }
                cp_async_arrive_noinc(&full[st]);
            }
        }
    } else if (lane == 0) {
        unsigned long long acc = 0;
        for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x, it++) {
            const int st = it % NS, ph = (it / NS) & 1;
            mbar_wait(&full[st], ph);
            acc += *reinterpret_cast<volatile unsigned *>(smem + (size_t)st * stage_bytes);
            mbar_arrive(&empty[st]);
        }
        if (acc == 0x1234567) *sink = acc;
    }
}

template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a); for (int i = 0; i < reps; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main() {
    const int H = 1024, W = 1024, G = 48;
    size_t n4 = (size_t)H * W * G;
    float4 *vol; unsigned long long *sink;
    cudaMalloc(&vol, n4 * 16); cudaMalloc(&sink, 8); cudaMemset(vol, 0, n4 * 16);
    const double gb_unique = n4 * 16 / 1e9;
    for (int per_sm = 1; per_sm <= 4; per_sm++) {
        const int grid = 148 * per_sm;
        {
            const int smem = NS * PX * ROWS * 256;
            cudaFuncSetAttribute(k_ring<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            float ms = timeit([&] { k_ring<0><<<grid, 64, smem>>>(vol, G, H, W, sink); });
            printf("bulk 256 B/op      %d CTA/SM: %.3f ms  unique %.0f GB/s  staged %.0f GB/s  %.1f Mops/s/SM\n", per_sm, ms, gb_unique / ms * 1e3,
                   gb_unique * ROWS * PX / S / ms * 1e3, (double)H * (W / S) * (G / 16) * PX * ROWS / ms / 1e3 / 148);
            cudaFuncSetAttribute(k_ring<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            ms = timeit([&] { k_ring<1><<<grid, 64, smem>>>(vol, G, H, W, sink); });
            printf("ldgsts 16 B x 16   %d CTA/SM: %.3f ms  unique %.0f GB/s  staged %.0f GB/s\n", per_sm, ms, gb_unique / ms * 1e3,
                   gb_unique * ROWS * PX / S / ms * 1e3);
        }
        if (per_sm <= 1) {
            const int smem = NS * PX * ROWS * G * 16;
            cudaFuncSetAttribute(k_ring<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            float ms = timeit([&] { k_ring<2><<<grid, 64, smem>>>(vol, G, H, W, sink); });
            printf("bulk %d B/op       %d CTA/SM: %.3f ms  unique %.0f GB/s  staged %.0f GB/s\n", G * 16, per_sm, ms, gb_unique / ms * 1e3,
                   gb_unique * ROWS * PX / S / ms * 1e3);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    return 0;
}
