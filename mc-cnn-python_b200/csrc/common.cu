// Error string, launch accounting and the two layout-conversion kernels.
#include "common.cuh"

#include <atomic>
#include <string.h>

namespace mccnn {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void clear_error() { g_err[0] = 0; }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int check_launch(const char *what) {
    count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return MCCNN_ERR_CUDA;
    }
    return MCCNN_OK;
}

// [D][H][W] -> [H][W][Dp]: a 32x32 (pixel x disparity) tile through shared memory so that both
// the reads (contiguous in w) and the writes (contiguous in d) are coalesced.
__global__ void k_dhw_to_hwd(const float *__restrict__ src, float *__restrict__ dst, int D, int Dp, long long P) {
    __shared__ float tile[32][33];
    long long p0 = (long long)blockIdx.x * 32;
    int d0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int d = d0 + j;
        long long p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (d < D && p < P) ? src[(long long)d * P + p] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        long long p = p0 + j;
        int d = d0 + threadIdx.x;
        if (p < P && d < Dp) dst[p * Dp + d] = tile[threadIdx.x][j];
    }
}

__global__ void k_hwd_to_dhw(const float *__restrict__ src, float *__restrict__ dst, int D, int Dp, long long P) {
    __shared__ float tile[32][33];
    long long p0 = (long long)blockIdx.x * 32;
    int d0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        long long p = p0 + j;
        int d = d0 + threadIdx.x;
        tile[j][threadIdx.x] = (p < P && d < D) ? src[p * Dp + d] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int d = d0 + j;
        long long p = p0 + threadIdx.x;
        if (d < D && p < P) dst[(long long)d * P + p] = tile[threadIdx.x][j];
    }
}

// Strided block copy [n0][n1][n2 granules of 16 bytes] (re-partitioning of HWD volumes between slab layouts):
// consecutive threads take consecutive granules of the innermost run, so both sides move whole 16-byte words of
// contiguous runs (n2 granules = one pixel's disparity sub-range or a whole pixel).
__global__ void __launch_bounds__(256) k_copy3d(const float4 *__restrict__ src, float4 *__restrict__ dst, long long n0,
                                                long long n1, int n2, long long ss0, long long ss1, long long ds0,
                                                long long ds1) {
    const long long total = n0 * n1 * n2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % n2);
        const long long r = i / n2;
        const long long i1 = r % n1, i0 = r / n1;
        dst[i0 * ds0 + i1 * ds1 + g] = __ldcs(src + i0 * ss0 + i1 * ss1 + g);
    }
}

}  // namespace mccnn

using namespace mccnn;

extern "C" {

const char *mccnn_last_error(void) { return g_err; }
int mccnn_abi_version(void) { return 2; }
int mccnn_dpitch(int D) { return dpitch(D); }
unsigned long long mccnn_launch_count(void) { return g_launches.load(); }

int mccnn_copy3d(const float *src, float *dst, long long n0, long long n1, int n2_granules, long long src_stride0,
                 long long src_stride1, long long dst_stride0, long long dst_stride1, void *stream) {
    MCCNN_REQUIRE(src && dst && n0 >= 0 && n1 >= 0 && n2_granules >= 0, "copy3d: bad arguments");
    MCCNN_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "copy3d: buffers must be 16-byte aligned");
    const long long total = n0 * n1 * n2_granules;
    if (total == 0) return MCCNN_OK;
    const long long blocks = (total + 255) / 256;
    const int grid = (int)(blocks < 148 * 32 ? blocks : 148 * 32);
    k_copy3d<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst), n0, n1,
                                                     n2_granules, src_stride0, src_stride1, dst_stride0, dst_stride1);
    MCCNN_LAUNCHED("copy3d");
    return MCCNN_OK;
}

int mccnn_dhw_to_hwd(const float *dhw, float *hwd, int D, int H, int W, void *stream) {
    MCCNN_REQUIRE(dhw && hwd && D >= 1 && H >= 1 && W >= 1, "dhw_to_hwd: bad arguments");
    long long P = (long long)H * W;
    int Dp = dpitch(D);
    dim3 grid(cdiv(P, 32), cdiv(Dp, 32)), block(32, 8);
    k_dhw_to_hwd<<<grid, block, 0, (cudaStream_t)stream>>>(dhw, hwd, D, Dp, P);
    MCCNN_LAUNCHED("dhw_to_hwd");
    return MCCNN_OK;
}

int mccnn_hwd_to_dhw(const float *hwd, float *dhw, int D, int H, int W, void *stream) {
    MCCNN_REQUIRE(dhw && hwd && D >= 1 && H >= 1 && W >= 1, "hwd_to_dhw: bad arguments");
    long long P = (long long)H * W;
    int Dp = dpitch(D);
    dim3 grid(cdiv(P, 32), cdiv(D, 32)), block(32, 8);
    k_hwd_to_dhw<<<grid, block, 0, (cudaStream_t)stream>>>(hwd, dhw, D, Dp, P);
    MCCNN_LAUNCHED("hwd_to_dhw");
    return MCCNN_OK;
}

}  // extern "C"
