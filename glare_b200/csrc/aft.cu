// Elementwise glue of the Adaptive Feature Transformation decoder (MultiScaleDecoder2.forward, deformableDecoder_arch.py:553-567):
//   Mix.forward (:587-590)      h <- enc * m + h * (1 - m),  m = sigmoid(w)
//   mean-ratio residual (:567)  h <- h + x_vq * (mean(h) / mean(x_vq))            (per sample, DESIGN.md "batch coupling")
// Both are  out = a * alpha[n] + b * beta[n]  over one sample's activations, evaluated with the reference's rounding sequence
// (each product rounded, then the sum: three ATen kernels and two temporaries in the reference, 7 / 5 tensor passes instead of 3).
// HBM-bound streaming kernel, 128-bit accesses.
#include "common.cuh"

namespace glare {

__global__ void __launch_bounds__(256) aft_axpby_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float* __restrict__ alpha,
                                                        const float* __restrict__ beta, int alpha_stride, int beta_stride, long long n4_per_sample,
                                                        float4* __restrict__ out) {
    const int n = blockIdx.y;
    const float al = __ldg(alpha + (long long)n * alpha_stride), be = __ldg(beta + (long long)n * beta_stride);
    const long long base = (long long)n * n4_per_sample;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_per_sample; i += 2 * stride) {
        const long long j = i + stride;
        const bool has = j < n4_per_sample;
        const float4 a0 = __ldg(a + base + i), b0 = __ldg(b + base + i);
        const float4 a1 = has ? __ldg(a + base + j) : a0, b1 = has ? __ldg(b + base + j) : b0;
        out[base + i] = make_float4(__fadd_rn(__fmul_rn(a0.x, al), __fmul_rn(b0.x, be)), __fadd_rn(__fmul_rn(a0.y, al), __fmul_rn(b0.y, be)),
                                    __fadd_rn(__fmul_rn(a0.z, al), __fmul_rn(b0.z, be)), __fadd_rn(__fmul_rn(a0.w, al), __fmul_rn(b0.w, be)));
        if (has)
            out[base + j] = make_float4(__fadd_rn(__fmul_rn(a1.x, al), __fmul_rn(b1.x, be)), __fadd_rn(__fmul_rn(a1.y, al), __fmul_rn(b1.y, be)),
                                        __fadd_rn(__fmul_rn(a1.z, al), __fmul_rn(b1.z, be)), __fadd_rn(__fmul_rn(a1.w, al), __fmul_rn(b1.w, be)));
    }
}

// WarpBlock.forward's torch.cat([x_vq, h], dim=1) (deformableDecoder_arch.py:286) fused with the operand conversion of the conv that
// consumes it: a [P][Ca], b [P][Cb] fp32 (NHWC pixels) -> the bf16x3 operand [P][2 * (Ca + Cb)] of the concatenated tensor, written directly
// (no fp32 cat tensor: saves its write and its read by the conversion pass)
__global__ void __launch_bounds__(256) aft_cat_operand_kernel(const float4* __restrict__ a, const float4* __restrict__ b, long long P, int ca4, int cb4,
                                                              __nv_bfloat16* __restrict__ out) {
    const int c4n = ca4 + cb4;
    const long long n4 = P * c4n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / c4n;
        const int c4 = (int)(i - p * c4n);
        const float4 v = c4 < ca4 ? __ldg(a + p * ca4 + c4) : __ldg(b + p * cb4 + (c4 - ca4));
        store_b3_4(out, i * 4, v.x, v.y, v.z, v.w);            // element index of the concatenated tensor: p * (Ca + Cb) + c
    }
}

}  // namespace glare

using namespace glare;

// a NHWC [P][Ca], b NHWC [P][Cb] fp32 -> bf16x3 operand of cat([a, b], channels): out [P][2 * (Ca + Cb)] bf16 (Ca, Cb multiples of 32)
GLARE_API int glare_aft_cat_operand(const float* a, const float* b, long long P, int Ca, int Cb, void* out, cudaStream_t stream) {
    if (P < 0 || Ca <= 0 || Cb <= 0 || (Ca & 31) || (Cb & 31)) return GLARE_ERR_BAD_ARG;
    if (P == 0) return GLARE_OK;
    if (!a || !b || !out || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15)) return GLARE_ERR_BAD_ARG;
    const long long n4 = P * ((Ca + Cb) / 4);
    const int grid = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    aft_cat_operand_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), P, Ca / 4, Cb / 4,
                                                     reinterpret_cast<__nv_bfloat16*>(out));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// out[n][i] = a[n][i] * alpha[n * alpha_stride] + b[n][i] * beta[n * beta_stride], i < n_per_sample (multiple of 4), n < B; strides 0 or 1;
// out may alias a or b
GLARE_API int glare_aft_axpby_f32(const float* a, const float* b, const float* alpha, const float* beta, int alpha_stride, int beta_stride, int B,
                                  long long n_per_sample, float* out, cudaStream_t stream) {
    if (B < 0 || n_per_sample < 0 || (n_per_sample & 3) || (alpha_stride & ~1) || (beta_stride & ~1)) return GLARE_ERR_BAD_ARG;
    if (B == 0 || n_per_sample == 0) return GLARE_OK;
    if (!a || !b || !alpha || !beta || !out || B > 65535) return GLARE_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) return GLARE_ERR_BAD_ARG;
    const long long n4 = n_per_sample / 4;
    long long blocks = (n4 + 511) / 512;
    const long long cap = (148LL * 16 + B - 1) / B;
    if (blocks > cap) blocks = cap;
    aft_axpby_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, stream>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                                              alpha, beta, alpha_stride, beta_stride, n4,
                                                                              reinterpret_cast<float4*>(out));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
