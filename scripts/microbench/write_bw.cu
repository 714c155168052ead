// Microbenchmark: write-only and write-heavy HBM bandwidth on B200 (what bounds the cost-volume kernel, which writes
// 1.49 GB and reads 0.55 GB).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_bw write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_write4(float4 *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (; i < n; i += stride) out[i] = v;
}
// each warp writes 128-byte runs of 32 floats, one run per "column", at a stride of Dp floats, starting misaligned by `mis`
__global__ void k_write_runs(float *out, size_t npix, int Dp, int mis) {
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int runs = Dp / 32;          // runs per pixel
    for (size_t r = warp; r < npix * runs; r += nwarps) {
        size_t p = r / runs; int k = (int)(r % runs);
        size_t idx = p * Dp + k * 32 + lane + mis;
        if (idx < npix * Dp) out[idx] = 1.0f;
    }
}
__global__ void k_read4(const float4 *in, float4 *sink, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    float4 acc = make_float4(0, 0, 0, 0);
    for (; i < n; i += stride) { float4 v = in[i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    if (acc.x == 12345.f) sink[0] = acc;
}
template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a); for (int i = 0; i < reps; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main() {
    const size_t npix = 1024 * 1024; const int Dp = 192;
    size_t n4 = npix * Dp / 4;
    float4 *a, *b; cudaMalloc(&a, n4 * 16); cudaMalloc(&b, n4 * 16); cudaMemset(a, 0, n4 * 16);
    double gb = n4 * 16 / 1e9; float ms;
    ms = timeit([&] { k_write4<<<148 * 16, 256>>>(a, n4); });
    printf("write-only float4 linear      %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { cudaMemsetAsync(a, 0, n4 * 16); });
    printf("cudaMemset                    %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { k_read4<<<148 * 16, 256>>>(a, b, n4); });
    printf("read-only float4 linear       %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    for (int mis = 0; mis <= 16; mis += 8) {
        ms = timeit([&] { k_write_runs<<<148 * 16, 256>>>((float *)a, npix, Dp, mis); });
        printf("write 128B runs, misalign %2d   %.3f ms  %.0f GB/s\n", mis, ms, gb / ms * 1e3);
    }
    // two volumes written concurrently (like L and R)
    ms = timeit([&] { k_write4<<<148 * 8, 256>>>(a, n4); k_write4<<<148 * 8, 256>>>(b, n4); });
    printf("two write-only kernels         %.3f ms  %.0f GB/s\n", ms, 2 * gb / ms * 1e3);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
