// Siamese feature network, forward only (model.py:40-64, :90-125; pf:15-73):
// zero-pad the image by num_layers pixels once (pf:20-25), then num_layers 3x3 VALID
// cross-correlations with 64 maps (NHWC activations, HWIO weights, stride 1), bias, ReLU after all
// but the last (model.py:51-60), and x * rsqrt(max(sum_c x^2, 1e-12)) over channels (model.py:64).
//
// Direct-convolution fp32 kernels (the 1e-4 cost-volume tolerance rules out plain TF32/BF16
// tensor-core math).  k_conv1 is bandwidth bound (1 input channel).  k_conv64 is an
// implicit GEMM with M = 128 output pixels (8 x 16 tile), N = 64, K = 576 walked in chunks of 8
// input channels staged in shared memory; each thread owns 8 pixels x 8 channels and reuses the 10
// activations of a patch row across the three horizontal taps (9 shared loads per 192 FMAs).
#include "common.cuh"

namespace mccnn {

constexpr int F = 64;              // feature maps (model.py:38)

// Layer 1: 1 -> 64.  One thread per (output pixel, 4 output channels).
__global__ void __launch_bounds__(256) k_conv1(const float *__restrict__ img, const float *__restrict__ wgt,
                                               const float *__restrict__ bias, float *__restrict__ out, int H, int W,
                                               int pad, int OH, int OW, int relu) {
    __shared__ float ws[9 * F];
    __shared__ float bs[F];
    for (int i = threadIdx.x; i < 9 * F; i += blockDim.x) ws[i] = wgt[i];
    for (int i = threadIdx.x; i < F; i += blockDim.x) bs[i] = bias[i];
    __syncthreads();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long px = t >> 4;
    int oc = (int)(t & 15) * 4;
    if (px >= (long long)OH * OW) return;
    int y = (int)(px / OW), x = (int)(px % OW);
    float4 acc = make_float4(bs[oc], bs[oc + 1], bs[oc + 2], bs[oc + 3]);
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            int iy = y + ky - pad, ix = x + kx - pad;             // coordinates in the unpadded image
            float v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? img[(size_t)iy * W + ix] : 0.0f;
            const float *wp = ws + (ky * 3 + kx) * F + oc;
            sum.x = fmaf(v, wp[0], sum.x); sum.y = fmaf(v, wp[1], sum.y);
            sum.z = fmaf(v, wp[2], sum.z); sum.w = fmaf(v, wp[3], sum.w);
        }
    acc.x += sum.x; acc.y += sum.y; acc.z += sum.z; acc.w += sum.w;
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    reinterpret_cast<float4 *>(out)[px * 16 + (oc >> 2)] = acc;
}

// Layers 2..n: 64 -> 64.
constexpr int CT_H = 8, CT_W = 16;             // output tile
constexpr int ICC = 8;                         // input channels per shared-memory chunk
constexpr int PR = CT_H + 2, PC = CT_W + 2;    // patch rows / cols
constexpr int PCP = 20;                        // padded patch row pitch (floats)

template <bool LAST>
__global__ void __launch_bounds__(128) k_conv64(const float *__restrict__ in, const float *__restrict__ wgt,
                                                const float *__restrict__ bias, float *__restrict__ out, int IH, int IW,
                                                int OH, int OW) {
    __shared__ __align__(16) float patch[ICC][PR][PCP];
    __shared__ __align__(16) float wsm[9][ICC][F];
    const int tid = threadIdx.x;
    const int ocg = tid & 7, pxg = tid >> 3;
    const int row = pxg >> 1, c0 = (pxg & 1) * 8;
    const int y0 = blockIdx.y * CT_H, x0 = blockIdx.x * CT_W;

    float acc[8][8];
#pragma unroll
    for (int p = 0; p < 8; p++)
#pragma unroll
        for (int o = 0; o < 8; o++) acc[p][o] = 0.f;

    for (int ic0 = 0; ic0 < F; ic0 += ICC) {
        __syncthreads();
        // activations: PR x PC pixels x 8 channels (two float4 per pixel), transposed to [ic][row][col]
        for (int i = tid; i < PR * PC * 2; i += 128) {
            int half = i & 1, pix = i >> 1;
            int r = pix / PC, c = pix % PC;
            int iy = y0 + r, ix = x0 + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy < IH && ix < IW)
                v = *reinterpret_cast<const float4 *>(in + ((size_t)iy * IW + ix) * F + ic0 + half * 4);
            patch[half * 4 + 0][r][c] = v.x;
            patch[half * 4 + 1][r][c] = v.y;
            patch[half * 4 + 2][r][c] = v.z;
            patch[half * 4 + 3][r][c] = v.w;
        }
        // weights: [tap][ic0..ic0+7][64]
        for (int i = tid; i < 9 * ICC * (F / 4); i += 128) {
            int tap = i / (ICC * 16), rem = i % (ICC * 16);
            int ic = rem >> 4, o4 = rem & 15;
            reinterpret_cast<float4 *>(&wsm[tap][ic][0])[o4] =
                *reinterpret_cast<const float4 *>(wgt + ((size_t)(tap * F + ic0 + ic)) * F + o4 * 4);
        }
        __syncthreads();
#pragma unroll 1
        for (int ic = 0; ic < ICC; ic++) {
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                float a[10];
                const float *pr = &patch[ic][row + ky][c0];
                float4 a0 = *reinterpret_cast<const float4 *>(pr);
                float4 a1 = *reinterpret_cast<const float4 *>(pr + 4);
                float2 a2 = *reinterpret_cast<const float2 *>(pr + 8);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                a[8] = a2.x; a[9] = a2.y;
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const float4 *wp = reinterpret_cast<const float4 *>(&wsm[ky * 3 + kx][ic][ocg * 8]);
                    float4 w0 = wp[0], w1 = wp[1];
                    float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int p = 0; p < 8; p++)
#pragma unroll
                        for (int o = 0; o < 8; o++) acc[p][o] = fmaf(a[p + kx], w[o], acc[p][o]);
                }
            }
        }
    }

    // epilogue: bias, then ReLU (model.py:120-123) or, on the last layer, channel L2-normalisation
    float b[8];
#pragma unroll
    for (int o = 0; o < 8; o++) b[o] = bias[ocg * 8 + o];
    const int y = y0 + row;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        float v[8];
#pragma unroll
        for (int o = 0; o < 8; o++) v[o] = acc[p][o] + b[o];
        if (LAST) {
            float ss = 0.f;
#pragma unroll
            for (int o = 0; o < 8; o++) ss = fmaf(v[o], v[o], ss);
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            ss += __shfl_xor_sync(0xffffffffu, ss, 4);
            const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));          // model.py:64
#pragma unroll
            for (int o = 0; o < 8; o++) v[o] *= inv;
        } else {
#pragma unroll
            for (int o = 0; o < 8; o++) v[o] = fmaxf(v[o], 0.f);
        }
        const int x = x0 + c0 + p;
        if (y < OH && x < OW) {
            float4 *dst = reinterpret_cast<float4 *>(out + ((size_t)y * OW + x) * F + ocg * 8);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
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

size_t mccnn_features_scratch_bytes(int H, int W, int pad, int num_layers) {
    if (H < 1 || W < 1 || num_layers < 1 || pad < 0 || H + 2 * pad < 3 || W + 2 * pad < 3) return 0;
    size_t oh = (size_t)H + 2 * pad - 2, ow = (size_t)W + 2 * pad - 2;
    return 2 * oh * ow * F * sizeof(float);
}

int mccnn_features(const float *img, int H, int W, int pad, int num_layers, const float *const *weights_host,
                   const float *const *biases_host, float *out, void *scratch, void *stream) {
    MCCNN_REQUIRE(img && weights_host && biases_host && out, "features: null pointer");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && pad >= 0 && num_layers >= 1 && num_layers <= 16,
                  "features: bad shape H=%d W=%d pad=%d layers=%d", H, W, pad, num_layers);
    MCCNN_REQUIRE(H + 2 * pad - 2 * num_layers >= 1 && W + 2 * pad - 2 * num_layers >= 1,
                  "features: image %dx%d (pad %d) too small for %d VALID 3x3 layers", H, W, pad, num_layers);
    MCCNN_REQUIRE(num_layers == 1 || scratch, "features: scratch required");
    cudaStream_t s = (cudaStream_t)stream;
    int oh = H + 2 * pad - 2, ow = W + 2 * pad - 2;
    float *buf[2];
    buf[0] = (float *)scratch;
    buf[1] = buf[0] ? buf[0] + (size_t)oh * ow * F : nullptr;
    float *dst = (num_layers == 1) ? out : buf[0];
    {
        long long threads = (long long)oh * ow * 16;
        k_conv1<<<cdiv(threads, 256), 256, 0, s>>>(img, weights_host[0], biases_host[0], dst, H, W, pad, oh, ow,
                                                   num_layers > 1);
        MCCNN_LAUNCHED("conv1");
    }
    if (num_layers == 1) {
        long long threads = (long long)oh * ow * 16;
        k_l2norm64<<<cdiv(threads, 256), 256, 0, s>>>(out, (long long)oh * ow);
        MCCNN_LAUNCHED("l2norm64");
        return MCCNN_OK;
    }
    const float *src = dst;
    int ih = oh, iw = ow;
    for (int l = 1; l < num_layers; l++) {
        oh = ih - 2; ow = iw - 2;
        const bool last = (l == num_layers - 1);
        float *d = last ? out : buf[l & 1];
        dim3 grid(cdiv(ow, CT_W), cdiv(oh, CT_H));
        if (last) k_conv64<true><<<grid, 128, 0, s>>>(src, weights_host[l], biases_host[l], d, ih, iw, oh, ow);
        else k_conv64<false><<<grid, 128, 0, s>>>(src, weights_host[l], biases_host[l], d, ih, iw, oh, ow);
        MCCNN_LAUNCHED("conv64");
        src = d; ih = oh; iw = ow;
    }
    return MCCNN_OK;
}

}  // extern "C"
