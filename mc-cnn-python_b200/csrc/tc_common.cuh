// Blackwell (sm_100a) building blocks shared by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 MMA / commit / TMEM load, shared-memory matrix descriptors, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace mccnn {

__device__ __forceinline__ unsigned tc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
        "@p bra TC_DONE_%=;\n"
        "bra TC_WAIT_%=;\n"
        "TC_DONE_%=:\n"
        "}\n" ::"r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
}
// same, for a lone waiting thread: back off between probes instead of competing for issue slots
__device__ __forceinline__ void tc_mbar_wait_sleep(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_SWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
        "@p bra TC_SDONE_%=;\n"
        "nanosleep.u32 200;\n"
        "bra TC_SWAIT_%=;\n"
        "TC_SDONE_%=:\n"
        "}\n" ::"r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2,
                                               unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
            "r"(tc_smem_u32(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2),
        "r"(tc_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// K-major, 128B-swizzled operand: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ unsigned long long tc_smem_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
    d |= (unsigned long long)1 << 16;                     // leading byte offset (unused for swizzled K-major)
    d |= (unsigned long long)(1024 >> 4) << 32;           // stride byte offset between 8-row groups
    d |= (unsigned long long)1 << 46;                     // descriptor version (sm_100)
    d |= (unsigned long long)2 << 61;                     // SWIZZLE_128B
    return d;
}
// D[tmem] (+)= -A[smem] . B[smem]^T, M = 128, N = 128, K = 8 (tf32)
__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, unsigned long long a_desc, unsigned long long b_desc,
                                            unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(tc_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}


// zero 32 consecutive TMEM columns of this warp's 32 lanes (completion: tcgen05.wait::st)
__device__ __forceinline__ void tc_tmem_st32_zero(unsigned taddr) {
    const unsigned z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr), "r"(z)
        : "memory");
}

__device__ __forceinline__ unsigned tc_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// hi = tf32(x), lo = x - hi (exact), both at the offset x has in `raw` (a swizzle permutes 16-byte chunks only)
__device__ __forceinline__ void tc_split(const unsigned char *raw, unsigned char *hi, unsigned char *lo, int nbytes, int ftid,
                                         int nthr) {
    const float4 *r4 = reinterpret_cast<const float4 *>(raw);
    float4 *h4 = reinterpret_cast<float4 *>(hi);
    float4 *l4 = reinterpret_cast<float4 *>(lo);
#pragma unroll 4
    for (int i = ftid; i < nbytes / 16; i += nthr) {
        const float4 x = r4[i];
        const float4 hv = make_float4(__uint_as_float(tc_tf32(x.x)), __uint_as_float(tc_tf32(x.y)),
                                      __uint_as_float(tc_tf32(x.z)), __uint_as_float(tc_tf32(x.w)));
        h4[i] = hv;
        l4[i] = make_float4(x.x - hv.x, x.y - hv.y, x.z - hv.z, x.w - hv.w);
    }
}

// fp16 tensor [d2][d1][d0] (d0 contiguous), box {b0, b1, 1}, 64-byte swizzle (b0 * 2 must be 64)
static inline int tc_encode_map_3d_f16_sw64(CUtensorMap &map, const void *base, unsigned long long d0, unsigned long long d1,
                                            unsigned long long d2, unsigned b0, unsigned b1, const char *what);

typedef CUresult (*TcEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// float32 tensor [d2][d1][d0] (d0 contiguous), box {b0, b1, b2}; swizzle128 = 128-byte swizzle (b0 * 4 must be 128)
static inline int tc_encode_map_3d(CUtensorMap &map, const float *base, unsigned long long d0, unsigned long long d1,
                                   unsigned long long d2, unsigned b0, unsigned b1, bool swizzle128, const char *what,
                                   unsigned b2 = 1) {
    static TcEncodeTiledFn enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            set_error("%s: cuTensorMapEncodeTiled is not available from this driver", what);
            return MCCNN_ERR_CUDA;
        }
        enc = (TcEncodeTiledFn)p;
    }
    const cuuint64_t gdim[3] = {d0, d1, d2};
    const cuuint64_t gstr[2] = {d0 * 4, d1 * d0 * 4};
    const cuuint32_t box[3] = {b0, b1, b2};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
        return MCCNN_ERR_CUDA;
    }
    return MCCNN_OK;
}

static inline int tc_encode_map_3d_f16_sw64(CUtensorMap &map, const void *base, unsigned long long d0, unsigned long long d1,
                                            unsigned long long d2, unsigned b0, unsigned b1, const char *what) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
        set_error("%s: cuTensorMapEncodeTiled is not available from this driver", what);
        return MCCNN_ERR_CUDA;
    }
    const cuuint64_t gdim[3] = {d0, d1, d2};
    const cuuint64_t gstr[2] = {d0 * 2, d1 * d0 * 2};
    const cuuint32_t box[3] = {b0, b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((TcEncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), gdim, gstr, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled (fp16) failed (%d)", what, (int)r);
        return MCCNN_ERR_CUDA;
    }
    return MCCNN_OK;
}

}  // namespace mccnn
