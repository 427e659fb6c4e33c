// GroupNorm(32, eps=1e-6) + swish for NHWC activations, producing the tensor-core operand of the next conv.
//
// Replaces reference encoder_decoder.py:34-35 (Normalize = GroupNorm(num_groups=32, eps=1e-6, affine)) and
// :29-31 (nonlinearity x*sigmoid(x)) as used by ResnetBlock.forward (:117-137), AttnBlock.forward (:168-171,
// no swish) and the norm_out heads (:435-442, :546-550).  HBM-bound: one read for the statistics, one read +
// one operand write for the apply pass (the conv that follows consumes the operand through TMA).
//   stats : per (sample, group) sum and sum of squares, accumulated in fp64 (no E[x^2]-E[x]^2 cancellation issue)
//   apply : y = (x - mean) * rstd * gamma + beta ; y *= sigmoid(y) ; emitted as bf16 | fp32 | tf32 hi+lo split
#include "common.cuh"

namespace glare {

constexpr int GN_THREADS = 256;

// x [B][HW][C]; grid (chunks, B); stats [B][G][2] (pre-zeroed).
// Shifted-data sums: every value is taken relative to c_g = x[b, pixel 0, first channel of the group] before it is squared, so the fp32
// partials stay at the scale of the group's spread even when |mean| >> std (sum of x^2 in fp32 loses the variance at mean / std ~ 1000:
// tests/test_edge_cases_gpu.py); each CTA converts its shifted sums back to true sums in fp64 (exact: c_g is an fp32 value).
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const float* __restrict__ x, long long HW, int C, int G,
                                                              double* __restrict__ stats) {
    __shared__ double s_sum[64], s_sq[64];
    const int b = blockIdx.y;
    const int c4n = C >> 2;                          // float4 columns per pixel
    const int rows = GN_THREADS / c4n;               // pixels processed per CTA pass (C <= 1024)
    const int my_c4 = threadIdx.x % c4n, my_row = threadIdx.x / c4n;
    if (threadIdx.x < 64) { s_sum[threadIdx.x] = 0.0; s_sq[threadIdx.x] = 0.0; }
    __syncthreads();
    const int cpg = C / G;
    const long long per = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long)blockIdx.x * per, p1 = (p0 + per < HW) ? p0 + per : HW;
    const float* xs = x + (long long)b * HW * C;
    double s = 0.0, q = 0.0;
    if (my_row < rows) {
        const int g = (my_c4 * 4) / cpg;
        const float c = __ldg(xs + g * cpg);         // the group's shift: its first channel at pixel 0 of the sample
        const float4* xb = reinterpret_cast<const float4*>(xs);
        // four independent 128-bit loads in flight per thread (a single one leaves HBM latency exposed)
        long long p = p0 + my_row;
        for (; p + 3LL * rows < p1; p += 4LL * rows) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(xb + (p + (long long)u * rows) * c4n + my_c4);
            // 16 shifted values in fp32 (relative error 1e-7 of a 16-term sum), then into the fp64 running sums
            float s32 = 0.f, q32 = 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float d0 = v[u].x - c, d1 = v[u].y - c, d2 = v[u].z - c, d3 = v[u].w - c;
                s32 += (d0 + d1) + (d2 + d3);
                q32 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, q32))));
            }
            s += (double)s32;
            q += (double)q32;
        }
        for (; p < p1; p += rows) {
            const float4 v = __ldg(xb + p * c4n + my_c4);
            const double d0 = (double)v.x - c, d1 = (double)v.y - c, d2 = (double)v.z - c, d3 = (double)v.w - c;
            s += d0 + d1 + d2 + d3;
            q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        atomicAdd(&s_sum[g], s);
        atomicAdd(&s_sq[g], q);
    }
    __syncthreads();
    if (threadIdx.x < G && p1 > p0) {
        // shifted -> true sums over this CTA's n = (p1 - p0) * cpg values: sum x = S1 + n c, sum x^2 = S2 + 2 c S1 + n c^2
        const double c = (double)__ldg(xs + threadIdx.x * cpg), n = (double)(p1 - p0) * cpg;
        const double S1 = s_sum[threadIdx.x], S2 = s_sq[threadIdx.x];
        atomicAdd(&stats[((long long)b * G + threadIdx.x) * 2], S1 + n * c);
        atomicAdd(&stats[((long long)b * G + threadIdx.x) * 2 + 1], S2 + 2.0 * c * S1 + n * c * c);
    }
}

__device__ __forceinline__ float tf32_hi_n(float x) { return tf32_round(x); }

// OUT: 0 = bf16, 1 = fp32, 2 = tf32 hi + lo
template <int OUT>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, int swish, long long HW, int C, int G,
                                                              void* __restrict__ out_hi, float* __restrict__ out_lo) {
    extern __shared__ __align__(16) float s_ab[];    // [C] rstd*gamma, [C] beta, [C] mean for this sample
    const int b = blockIdx.y;
    const int cpg = C / G;
    const double cnt = (double)HW * cpg;
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
        const int g = c / cpg;
        const double mean = stats[((long long)b * G + g) * 2] / cnt;
        double var = stats[((long long)b * G + g) * 2 + 1] / cnt - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        s_ab[c] = gamma[c] * rstd;
        s_ab[C + c] = beta[c];
        s_ab[2 * C + c] = (float)mean;
    }
    __syncthreads();
    const int c4n = C >> 2;
    const int cmask = (c4n & (c4n - 1)) == 0 ? c4n - 1 : 0;
    const long long n4 = HW * c4n;
    const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * HW * C);
    const long long base4 = (long long)b * n4;
    const long long stride = (long long)gridDim.x * GN_THREADS;
    for (long long i0 = (long long)blockIdx.x * GN_THREADS + threadIdx.x; i0 < n4; i0 += 2 * stride) {
        // two independent 128-bit loads in flight per thread
        const long long i1 = i0 + stride;
        const bool has1 = i1 < n4;
        const float4 va = __ldg(xb + i0);
        const float4 vb = has1 ? __ldg(xb + i1) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !has1) break;
            const long long i = u == 0 ? i0 : i1;
            const float4 v = u == 0 ? va : vb;
            const int c = (cmask ? ((int)i & cmask) : (int)(i % c4n)) * 4;       // C / 4 is a power of two for every GLARE layer
            const float4 ga = *reinterpret_cast<const float4*>(s_ab + c), be = *reinterpret_cast<const float4*>(s_ab + C + c);
            const float4 mu = *reinterpret_cast<const float4*>(s_ab + 2 * C + c);
            float y[4] = {fmaf(v.x - mu.x, ga.x, be.x), fmaf(v.y - mu.y, ga.y, be.y), fmaf(v.z - mu.z, ga.z, be.z), fmaf(v.w - mu.w, ga.w, be.w)};
            if (swish) {
                if (OUT == 0 || OUT == 4) {
                    // operands of 8 / 16 significant bits: ex2.approx + fast divide (~1e-6 relative) is below their resolution and keeps
                    // this pass HBM-bound instead of ALU-bound
#pragma unroll
                    for (int k = 0; k < 4; ++k) y[k] = __fdividef(y[k], 1.0f + ex2_approx(-1.4426950408889634f * y[k]));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) y[k] = y[k] / (1.0f + expf(-y[k]));
                }
            }
            if (OUT == 0) {
                __nv_bfloat162 a0 = __floats2bfloat162_rn(y[0], y[1]), a1 = __floats2bfloat162_rn(y[2], y[3]);
                uint2 o;
                o.x = *reinterpret_cast<uint32_t*>(&a0);
                o.y = *reinterpret_cast<uint32_t*>(&a1);
                reinterpret_cast<uint2*>(out_hi)[base4 + i] = o;
            } else if (OUT == 1) {
                reinterpret_cast<float4*>(out_hi)[base4 + i] = make_float4(y[0], y[1], y[2], y[3]);
            } else if (OUT == 2) {
                const float4 h = make_float4(tf32_hi_n(y[0]), tf32_hi_n(y[1]), tf32_hi_n(y[2]), tf32_hi_n(y[3]));
                reinterpret_cast<float4*>(out_hi)[base4 + i] = h;
                reinterpret_cast<float4*>(out_lo)[base4 + i] = make_float4(y[0] - h.x, y[1] - h.y, y[2] - h.z, y[3] - h.w);
            } else if (OUT == 3) {
                const float4 h = make_float4(tf32_hi_n(y[0]), tf32_hi_n(y[1]), tf32_hi_n(y[2]), tf32_hi_n(y[3]));
                reinterpret_cast<float4*>(out_hi)[base4 + i] = h;
                store_x4(reinterpret_cast<__nv_bfloat16*>(out_lo), (base4 + i) * 4, y[0], y[1], y[2], y[3], h.x, h.y, h.z, h.w);
            } else {
                store_b3_4(reinterpret_cast<__nv_bfloat16*>(out_hi), (base4 + i) * 4, y[0], y[1], y[2], y[3]);
            }
        }
    }
}

}  // namespace glare

using namespace glare;

// x NHWC [B,HW,C] fp32 -> stats [B][G][2] fp64 (sum, sum of squares); stats is overwritten.
GLARE_API int glare_gn_stats_nhwc_f32(const float* x, int B, long long HW, int C, int G, double* stats, cudaStream_t stream) {
    if (B < 0 || HW < 0 || C <= 0 || G <= 0 || G > 64 || C % G != 0 || (C & 3) || C > 1024 || ((C / G) & 3)) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !stats || B > 65535) return GLARE_ERR_BAD_ARG;
    GLARE_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)B * G, stream));
    if (HW == 0) return GLARE_OK;
    const int rows = GN_THREADS / (C / 4);
    long long chunks = (HW + (long long)rows * 16 - 1) / ((long long)rows * 16);      // >= 16 pixels per thread row
    const long long cap = (148 * 8 + B - 1) / B;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    gn_stats_kernel<<<dim3((unsigned)chunks, (unsigned)B), GN_THREADS, 0, stream>>>(x, HW, C, G, stats);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// out_mode 0: bf16 -> out_hi; 1: fp32 -> out_hi; 2: tf32 hi -> out_hi, lo -> out_lo.  swish != 0 applies x*sigmoid(x).
GLARE_API int glare_gn_apply_nhwc(int out_mode, const float* x, const double* stats, const float* gamma, const float* beta,
                                  float eps, int swish, int B, long long HW, int C, int G, void* out_hi, void* out_lo,
                                  cudaStream_t stream) {
    if (out_mode < 0 || out_mode > 4 || B < 0 || HW < 0 || C <= 0 || G <= 0 || C % G != 0 || (C & 3) || C > 1024) return GLARE_ERR_BAD_ARG;
    if (out_mode >= 3 && (C & 31)) return GLARE_ERR_UNSUPPORTED;
    if (B == 0 || HW == 0) return GLARE_OK;
    if (!x || !stats || !gamma || !beta || !out_hi || ((out_mode == 2 || out_mode == 3) && !out_lo) || B > 65535) return GLARE_ERR_BAD_ARG;
    const long long n4 = HW * (C / 4);
    long long blocks = (n4 + GN_THREADS * 4 - 1) / (GN_THREADS * 4);
    const long long cap = (148 * 16 + B - 1) / B;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const dim3 grid((unsigned)blocks, (unsigned)B);
    const size_t smem = 3 * (size_t)C * sizeof(float);
    float* lo = reinterpret_cast<float*>(out_lo);
    if (out_mode == 0) gn_apply_kernel<0><<<grid, GN_THREADS, smem, stream>>>(x, stats, gamma, beta, eps, swish, HW, C, G, out_hi, lo);
    else if (out_mode == 1) gn_apply_kernel<1><<<grid, GN_THREADS, smem, stream>>>(x, stats, gamma, beta, eps, swish, HW, C, G, out_hi, lo);
    else if (out_mode == 2) gn_apply_kernel<2><<<grid, GN_THREADS, smem, stream>>>(x, stats, gamma, beta, eps, swish, HW, C, G, out_hi, lo);
    else if (out_mode == 3) gn_apply_kernel<3><<<grid, GN_THREADS, smem, stream>>>(x, stats, gamma, beta, eps, swish, HW, C, G, out_hi, lo);
    else gn_apply_kernel<4><<<grid, GN_THREADS, smem, stream>>>(x, stats, gamma, beta, eps, swish, HW, C, G, out_hi, lo);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
