// Microbenchmark: how fast can B200 copy an [H][W][Dp] float volume when each CTA moves a
// 16x16-pixel tile x SLAB bytes per pixel (the CBCA access pattern), vs a plain linear copy?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pattern_copy pattern_copy.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_linear(const float4 *in, float4 *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

// grid (slabs, tilesX, tilesY); slab = GC float4 per pixel
template <int GC>
__global__ void k_tiled(const float4 *in, float4 *out, int G, int H, int W) {
    const int w0 = blockIdx.y * 16, h0 = blockIdx.z * 16, g0 = blockIdx.x * GC;
    for (int i = threadIdx.x; i < 256 * GC; i += blockDim.x) {
        int gc = i % GC, px = (i / GC) % 16, r = i / (GC * 16);
        size_t a = ((size_t)(h0 + r) * W + w0 + px) * G + g0 + gc;
        out[a] = in[a];
    }
}

// same but tile = 1 row x 256 px (long contiguous-ish runs along w)
template <int GC>
__global__ void k_rowtile(const float4 *in, float4 *out, int G, int H, int W) {
    const int w0 = blockIdx.y * 256, h = blockIdx.z, g0 = blockIdx.x * GC;
    for (int i = threadIdx.x; i < 256 * GC; i += blockDim.x) {
        int gc = i % GC, px = i / GC;
        size_t a = ((size_t)h * W + w0 + px) * G + g0 + gc;
        out[a] = in[a];
    }
}

template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const int H = 1024, W = 1024, G = 48;
    size_t n = (size_t)H * W * G;
    float4 *in, *out;
    cudaMalloc(&in, n * 16); cudaMalloc(&out, n * 16);
    cudaMemset(in, 0, n * 16);
    double gb = 2.0 * n * 16 / 1e9;
    float ms;
    ms = timeit([&] { k_linear<<<148 * 16, 256>>>(in, out, n); });
    printf("linear copy            %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { cudaMemcpyAsync(out, in, n * 16, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D         %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
#define T(GC, TH) ms = timeit([&] { k_tiled<GC><<<dim3(G / GC, W / 16, H / 16), TH>>>(in, out, G, H, W); }); \
    printf("tiled 16x16 x %3d B, %3d thr  %.3f ms  %.0f GB/s\n", GC * 16, TH, ms, gb / ms * 1e3);
    T(1, 256) T(2, 256) T(4, 128) T(4, 256) T(8, 256) T(16, 256) T(48, 256)
#define R(GC) ms = timeit([&] { k_rowtile<GC><<<dim3(G / GC, W / 256, H), 256>>>(in, out, G, H, W); }); \
    printf("rowtile 1x256 x %3d B         %.3f ms  %.0f GB/s\n", GC * 16, ms, gb / ms * 1e3);
    R(4) R(8) R(48)
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
