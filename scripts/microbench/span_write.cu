// Microbenchmark: how B200 takes misaligned contiguous spans written by different warps (the cost-volume epilogue's pattern:
// every (tile, chunk) pair produces, per pixel, a contiguous run of disparities that starts at an arbitrary float).
// The array is covered exactly once by consecutive spans of SP floats shifted by `mis` floats; span r is written by one warp
// instruction group: scalar (one float per lane) or float4 body + scalar head / tail.  No divisions in the loops.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o span_write span_write.cu
#include <cstdio>
#include <cuda_runtime.h>

// SP = 32: one scalar store per span (lane = float)
__global__ void k_scalar32(float *out, size_t nspans, int mis) {
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = warp; r + 1 < nspans; r += nwarps) out[r * 32 + mis + lane] = 1.0f;
}
// SP = 128 floats: float4 body (lanes over 16-byte pieces) + one scalar instruction for head and tail
__global__ void k_span128(float *out, size_t nspans, int mis) {
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int a = mis & 3, head = (4 - a) & 3, nbody = (128 - head) >> 2, tail = 128 - head - 4 * nbody;
    for (size_t r = warp; r + 1 < nspans; r += nwarps) {
        float *s = out + r * 128 + mis;
        if (lane < nbody) *reinterpret_cast<float4 *>(s + head + 4 * lane) = make_float4(1.f, 2.f, 3.f, 4.f);
        if (lane < head) s[lane] = 5.f;
        else if (lane >= 4 && lane < 4 + tail) s[head + 4 * nbody + lane - 4] = 6.f;
    }
}
// SP = 32 floats, four spans per warp instruction (8 lanes x 16 B each) + one scalar instruction for the heads and tails
__global__ void k_span32(float *out, size_t nspans, int mis) {
    const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int a = mis & 3, head = (4 - a) & 3, nbody = (32 - head) >> 2, tail = 32 - head - 4 * nbody;
    for (size_t r = warp * 4; r + 4 < nspans; r += nwarps * 4) {
        float *s = out + (r + sub) * 32 + mis;
        if (l8 < nbody) *reinterpret_cast<float4 *>(s + head + 4 * l8) = make_float4(1.f, 2.f, 3.f, 4.f);
        if (l8 < head) s[l8] = 5.f;
        else if (l8 >= 4 && l8 < 4 + tail) s[head + 4 * nbody + l8 - 4] = 6.f;
    }
}
__global__ void k_write4(float4 *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a); for (int i = 0; i < reps; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main() {
    const size_t nfl = (size_t)1024 * 1024 * 192 * 2;          // two C3 volumes
    float *a; cudaMalloc(&a, nfl * 4 + 4096); cudaMemset(a, 0, nfl * 4);
    const double gb = nfl * 4 / 1e9; float ms;
    for (int blocks : {148 * 2, 148 * 8}) {
        printf("grid %d x 256\n", blocks);
        ms = timeit([&] { k_write4<<<blocks, 256>>>((float4 *)a, nfl / 4); });
        printf("  float4 linear                         %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
        for (int mis : {0, 8, 4, 1, 9}) {
            ms = timeit([&] { k_scalar32<<<blocks, 256>>>(a, nfl / 32, mis); });
            printf("  scalar 128-B spans, shifted %2d floats  %.3f ms  %.0f GB/s\n", mis, ms, gb / ms * 1e3);
        }
        for (int mis : {0, 4, 1, 6}) {
            ms = timeit([&] { k_span32<<<blocks, 256>>>(a, nfl / 32, mis); });
            printf("  float4+ends 128-B spans, shifted %2d    %.3f ms  %.0f GB/s\n", mis, ms, gb / ms * 1e3);
        }
        for (int mis : {0, 4, 1, 6}) {
            ms = timeit([&] { k_span128<<<blocks, 256>>>(a, nfl / 128, mis); });
            printf("  float4+ends 512-B spans, shifted %2d    %.3f ms  %.0f GB/s\n", mis, ms, gb / ms * 1e3);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
