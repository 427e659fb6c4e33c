// Backward-pass kernels of the taming-style blocks of the condition encoder (stage-2 training, BASELINE config 4): the reference runs these
// through torch autograd of encoder_decoder.py:29-35 (Normalize + nonlinearity), :117-137 (ResnetBlock), :168-192 (AttnBlock), :68-72 (Downsample).
// The dense parts of the backward (data gradients of convolutions, the attention GEMMs) reuse the tensor-core conv path with transposed / flipped
// operands and the split-K GEMM of dcn_bwd.cu; this file holds the memory-bound pieces in between (glare_b200/encoder_train.py is the host side):
//   gn_bwd_stats / gn_bwd_apply : GroupNorm(32) (+ swish) backward -- per (sample, group) sums of dxhat and dxhat * xhat, dgamma / dbeta, then dx
//   im2col_nhwc                 : [B,H,W,C] -> [B*Ho*Wo][k*k*C] (tap-major) for any 1x1 / 3x3, stride 1 / 2, low-side padding 0 / 1 conv: the
//                                 operand of a weight gradient  dW[tap*C + c][co] = sum_p col[p][tap*C + c] * dY[p][co]
//   attn_softmax_bwd            : dS = scale * P o (dP - rowsum(dP o P))
// STATUS: verified on the CPU (the same source through tests/cuda_emu) and on B200 (tests/flow_train_gpu_check.py, profiles/r70_train_check.log).
#ifdef GLARE_CUDA_EMU
#include "cuda_emu.h"
#define TE_LAUNCH(kern, grid, block, stream, ...) glare_emu::launch(kern, grid, dim3(block), __VA_ARGS__)
#else
#include "common.cuh"
#define TE_LAUNCH(kern, grid, block, stream, ...) kern<<<grid, block, 0, stream>>>(__VA_ARGS__)
#endif

namespace glare {

__device__ __forceinline__ float te_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// dn = dL/d(normalised, affine value n = xhat * gamma + beta) given dy = dL/d(output), output = swish ? n * sigmoid(n) : n
__device__ __forceinline__ float te_dn(float n, float dy, int swish) {
    if (!swish) return dy;
    const float s = te_sigmoid(n);
    return dy * (s * (1.0f + n * (1.0f - s)));
}

// stats [B][G][2] = (sum x, sum x^2) in fp64 as written by glare_gn_stats_nhwc_f32 -> mean, rstd of the group
__device__ __forceinline__ void te_mean_rstd(const double* stats, int b, int G, int g, double cnt, float eps, float& mean, float& rstd) {
    const double m = stats[((long long)b * G + g) * 2] / cnt;
    double var = stats[((long long)b * G + g) * 2 + 1] / cnt - m * m;
    var = var < 0.0 ? 0.0 : var;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// grid (chunks, B), 256 threads; x, gy NHWC [B][HW][C], C <= 1024, C % 4 == 0 handled per element (thread = one channel column walker)
// sums [B][G][2] += (sum dxhat, sum dxhat * xhat) ; dgamma[C] += sum dn * xhat ; dbeta[C] += sum dn
__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ gy, const double* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int swish,
                                                           long long HW, int C, int G, double* __restrict__ sums, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
    const int b = blockIdx.y;
    const int cpg = C / G;
    const double cnt = (double)HW * cpg;
    const long long per = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long)blockIdx.x * per, p1 = (p0 + per < HW) ? p0 + per : HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float mean, rstd;
        te_mean_rstd(stats, b, G, g, cnt, eps, mean, rstd);
        const float ga = __ldg(gamma + c), be = __ldg(beta + c);
        float s1 = 0.f, s2 = 0.f, sg = 0.f, sb = 0.f;
        for (long long p = p0; p < p1; ++p) {
            const long long i = ((long long)b * HW + p) * C + c;
            const float xh = (__ldg(x + i) - mean) * rstd;
            const float dn = te_dn(fmaf(xh, ga, be), __ldg(gy + i), swish);
            const float dxh = dn * ga;
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
            sg = fmaf(dn, xh, sg);
            sb += dn;
        }
        atomicAdd(sums + ((long long)b * G + g) * 2, (double)s1);
        atomicAdd(sums + ((long long)b * G + g) * 2 + 1, (double)s2);
        atomicAdd(dgamma + c, sg);
        atomicAdd(dbeta + c, sb);
    }
}

// gx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat * xhat))
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ gy, const double* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int swish,
                                                           long long HW, int C, int G, const double* __restrict__ sums, float* __restrict__ gx) {
    const int b = blockIdx.y;
    const int cpg = C / G;
    const double cnt = (double)HW * cpg;
    const long long n = HW * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C), g = c / cpg;
        float mean, rstd;
        te_mean_rstd(stats, b, G, g, cnt, eps, mean, rstd);
        const float m1 = (float)(sums[((long long)b * G + g) * 2] / cnt), m2 = (float)(sums[((long long)b * G + g) * 2 + 1] / cnt);
        const long long i = (long long)b * n + e;
        const float ga = __ldg(gamma + c);
        const float xh = (__ldg(x + i) - mean) * rstd;
        const float dxh = te_dn(fmaf(xh, ga, __ldg(beta + c)), __ldg(gy + i), swish) * ga;
        gx[i] = rstd * (dxh - m1 - xh * m2);
    }
}

// ---- fast path: C / 4 a power of two <= 256 and (C / G) % 4 == 0 (GLARE: 128 / 256 / 512 channels, 32 groups) --------------------------------
// thread = (walker, channel quad): float4 loads over the channels, 256 / (C / 4) pixel walkers per CTA so that every thread works whatever C
// is, two pixels in flight per walker.  Per CTA: one shared-memory reduction over the walkers, a shuffle reduction over the quads of a group,
// then ONE fp64 atomic pair per group and one fp32 atomic pair per channel.  (The per-channel column walkers above left half of the CTA idle
// at C = 128 and ran at 1 TB/s: 3.8 ms of a 58 ms stage-3 step.)
__device__ __forceinline__ void te_quad(const float4 xv, const float4 gv, const float4 ga, const float4 be, float mean, float rstd, int swish,
                                        float& s1, float& s2, float4& sg, float4& sb) {
    const float xh0 = (xv.x - mean) * rstd, xh1 = (xv.y - mean) * rstd, xh2 = (xv.z - mean) * rstd, xh3 = (xv.w - mean) * rstd;
    const float d0 = te_dn(fmaf(xh0, ga.x, be.x), gv.x, swish), d1 = te_dn(fmaf(xh1, ga.y, be.y), gv.y, swish);
    const float d2 = te_dn(fmaf(xh2, ga.z, be.z), gv.z, swish), d3 = te_dn(fmaf(xh3, ga.w, be.w), gv.w, swish);
    const float e0 = d0 * ga.x, e1 = d1 * ga.y, e2 = d2 * ga.z, e3 = d3 * ga.w;
    s1 += (e0 + e1) + (e2 + e3);
    s2 = fmaf(e0, xh0, fmaf(e1, xh1, fmaf(e2, xh2, fmaf(e3, xh3, s2))));
    sg.x = fmaf(d0, xh0, sg.x);
    sg.y = fmaf(d1, xh1, sg.y);
    sg.z = fmaf(d2, xh2, sg.z);
    sg.w = fmaf(d3, xh3, sg.w);
    sb.x += d0;
    sb.y += d1;
    sb.z += d2;
    sb.w += d3;
}

__global__ void __launch_bounds__(256) gn_bwd_stats_quad_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps, int swish, long long HW, int C, int G,
                                                                double* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float4 red_g[256], red_b[256];
    __shared__ float red_1[256], red_2[256];
    const int b = blockIdx.y;
    const int C4 = C >> 2, cpg = C / G;
    const int q = threadIdx.x % C4, w = threadIdx.x / C4, nwalk = 256 / C4;
    const int c = q * 4, g = c / cpg;
    const double cnt = (double)HW * cpg;
    float mean, rstd;
    te_mean_rstd(stats, b, G, g, cnt, eps, mean, rstd);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    const long long per = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long)blockIdx.x * per, p1 = (p0 + per < HW) ? p0 + per : HW;
    const float4* xs = reinterpret_cast<const float4*>(x + (long long)b * HW * C) + q;
    const float4* gs = reinterpret_cast<const float4*>(gy + (long long)b * HW * C) + q;
    float s1 = 0.f, s2 = 0.f;
    float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sb = make_float4(0.f, 0.f, 0.f, 0.f);
    long long p = p0 + w;
    for (; p + nwalk < p1; p += 2 * nwalk) {
        const float4 xa = xs[p * C4], gA = gs[p * C4], xb = xs[(p + nwalk) * C4], gB = gs[(p + nwalk) * C4];
        te_quad(xa, gA, ga, be, mean, rstd, swish, s1, s2, sg, sb);
        te_quad(xb, gB, ga, be, mean, rstd, swish, s1, s2, sg, sb);
    }
    if (p < p1) te_quad(xs[p * C4], gs[p * C4], ga, be, mean, rstd, swish, s1, s2, sg, sb);
    red_g[threadIdx.x] = sg;
    red_b[threadIdx.x] = sb;
    red_1[threadIdx.x] = s1;
    red_2[threadIdx.x] = s2;
    __syncthreads();
    if (w == 0) {
        for (int r = 1; r < nwalk; ++r) {
            const float4 a = red_g[r * C4 + q], bb = red_b[r * C4 + q];
            sg.x += a.x; sg.y += a.y; sg.z += a.z; sg.w += a.w;
            sb.x += bb.x; sb.y += bb.y; sb.z += bb.z; sb.w += bb.w;
            s1 += red_1[r * C4 + q];
            s2 += red_2[r * C4 + q];
        }
        atomicAdd(dgamma + c, sg.x); atomicAdd(dgamma + c + 1, sg.y); atomicAdd(dgamma + c + 2, sg.z); atomicAdd(dgamma + c + 3, sg.w);
        atomicAdd(dbeta + c, sb.x); atomicAdd(dbeta + c + 1, sb.y); atomicAdd(dbeta + c + 2, sb.z); atomicAdd(dbeta + c + 3, sb.w);
    }
    // the quads of a group are cpg / 4 (1, 2 or 4 ... <= 8) adjacent lanes of walker 0's warps; every lane of those warps takes part
    if (threadIdx.x < ((C4 + 31) & ~31)) {
        const int qpg = cpg >> 2;
        for (int o = 1; o < qpg; o <<= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (w == 0 && (q % qpg) == 0) {
            atomicAdd(sums + ((long long)b * G + g) * 2, (double)s1);
            atomicAdd(sums + ((long long)b * G + g) * 2 + 1, (double)s2);
        }
    }
}

// gx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat * xhat)) with the per-group constants (fp64 once per CTA) staged in shared memory
__global__ void __launch_bounds__(256) gn_bwd_apply_quad_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps, int swish, long long HW, int C, int G,
                                                                const double* __restrict__ sums, float* __restrict__ gx) {
    __shared__ float4 grp[64];                       // (mean, rstd, m1, m2) per group, G <= 64
    const int b = blockIdx.y;
    const int C4 = C >> 2, cpg = C / G;
    if (threadIdx.x < G) {
        const double cnt = (double)HW * cpg;
        float mean, rstd;
        te_mean_rstd(stats, b, G, threadIdx.x, cnt, eps, mean, rstd);
        grp[threadIdx.x] = make_float4(mean, rstd, (float)(sums[((long long)b * G + threadIdx.x) * 2] / cnt),
                                       (float)(sums[((long long)b * G + threadIdx.x) * 2 + 1] / cnt));
    }
    __syncthreads();
    const int q = threadIdx.x % C4, w = threadIdx.x / C4, nwalk = 256 / C4;
    const int c = q * 4;
    const float4 k = grp[c / cpg];
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    const float4* xs = reinterpret_cast<const float4*>(x + (long long)b * HW * C) + q;
    const float4* gs = reinterpret_cast<const float4*>(gy + (long long)b * HW * C) + q;
    float4* os = reinterpret_cast<float4*>(gx + (long long)b * HW * C) + q;
    for (long long p = (long long)blockIdx.x * nwalk + w; p < HW; p += (long long)gridDim.x * nwalk) {
        const float4 xv = xs[p * C4], gv = gs[p * C4];
        const float xh0 = (xv.x - k.x) * k.y, xh1 = (xv.y - k.x) * k.y, xh2 = (xv.z - k.x) * k.y, xh3 = (xv.w - k.x) * k.y;
        const float e0 = te_dn(fmaf(xh0, ga.x, be.x), gv.x, swish) * ga.x, e1 = te_dn(fmaf(xh1, ga.y, be.y), gv.y, swish) * ga.y;
        const float e2 = te_dn(fmaf(xh2, ga.z, be.z), gv.z, swish) * ga.z, e3 = te_dn(fmaf(xh3, ga.w, be.w), gv.w, swish) * ga.w;
        os[p * C4] = make_float4(k.y * (e0 - k.z - xh0 * k.w), k.y * (e1 - k.z - xh1 * k.w), k.y * (e2 - k.z - xh2 * k.w), k.y * (e3 - k.z - xh3 * k.w));
    }
}

// col[(b, oy, ox)][t * C + c] = x[b][oy * stride + t / k - pad][ox * stride + t % k - pad][c]  (zero outside the H x W image)
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const float* __restrict__ x, int B, int H, int W, int C, int k, int stride, int pad, int Ho,
                                                          int Wo, float* __restrict__ col) {
    const long long total = (long long)B * Ho * Wo * k * k * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int t = (int)(r % (k * k));
        r /= (k * k);
        const int ox = (int)(r % Wo);
        r /= Wo;
        const int oy = (int)(r % Ho), b = (int)(r / Ho);
        const int y = oy * stride + t / k - pad, xx = ox * stride + t % k - pad;
        col[i] = (y >= 0 && y < H && xx >= 0 && xx < W) ? __ldg(x + (((long long)b * H + y) * W + xx) * C + c) : 0.f;
    }
}

// one row per block: dS[r][j] = scale * P[r][j] * (dP[r][j] - sum_j' dP[r][j'] P[r][j']), j < n_keys
__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dP, long long ld, int n_keys,
                                                               float scale, float* __restrict__ dS) {
    __shared__ float s_red[8];
    __shared__ float s_dot;
    const long long row = blockIdx.x;
    const float* p = P + row * ld;
    const float* d = dP + row * ld;
    float acc = 0.f;
    for (int j = threadIdx.x; j < n_keys; j += blockDim.x) acc = fmaf(__ldg(p + j), __ldg(d + j), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_red[i];
        s_dot = t;
    }
    __syncthreads();
    const float dot = s_dot;
    float* o = dS + row * ld;
    for (int j = threadIdx.x; j < n_keys; j += blockDim.x) o[j] = scale * __ldg(p + j) * (__ldg(d + j) - dot);
}

// out[c] += sum_p x[p][c]: the bias gradient of a convolution (dY summed over the pixels of the batch).  grid (C / 128 rounded up, row chunks),
// 256 threads = 32 column quads x 8 row lanes; 512-byte row segments per warp, fp32 partials reduced through shared memory, one atomicAdd per
// column and CTA.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long P, int C, long long rows_per_cta, float* __restrict__ out) {
    __shared__ float4 part[8][32];
    const int quad = threadIdx.x & 31, lane_r = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + quad) * 4;
    const long long p0 = (long long)blockIdx.y * rows_per_cta, p1 = (p0 + rows_per_cta < P) ? p0 + rows_per_cta : P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C)
        for (long long p = p0 + lane_r; p < p1; p += 8) {
            const float4 v = *reinterpret_cast<const float4*>(x + p * C + c);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
    part[lane_r][quad] = acc;
    __syncthreads();
    if (lane_r == 0 && c < C) {
        for (int r = 1; r < 8; ++r) {
            acc.x += part[r][quad].x;
            acc.y += part[r][quad].y;
            acc.z += part[r][quad].z;
            acc.w += part[r][quad].w;
        }
        atomicAdd(out + c, acc.x);
        atomicAdd(out + c + 1, acc.y);
        atomicAdd(out + c + 2, acc.z);
        atomicAdd(out + c + 3, acc.w);
    }
}

}  // namespace glare

using namespace glare;

// GroupNorm(G, eps) (+ swish) backward on NHWC fp32: stats = (sum x, sum x^2) per (sample, group) from glare_gn_stats_nhwc_f32;
// sums [B][G][2] fp64 scratch (overwritten); gx NHWC; dgamma / dbeta [C] are ACCUMULATED into (caller zero-fills)
GLARE_API int glare_gn_bwd_nhwc_f32(const float* x, const float* gy, const double* stats, const float* gamma, const float* beta, float eps, int swish,
                                    int B, long long HW, int C, int G, double* sums, float* gx, float* dgamma, float* dbeta, cudaStream_t stream) {
    if (B < 0 || HW < 0 || C <= 0 || G <= 0 || C % G != 0) return GLARE_ERR_BAD_ARG;
    if (B == 0 || HW == 0) return GLARE_OK;
    if (!x || !gy || !stats || !gamma || !beta || !sums || !gx || !dgamma || !dbeta || B > 65535) return GLARE_ERR_BAD_ARG;
#ifndef GLARE_CUDA_EMU
    GLARE_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)B * G, stream));
#else
    memset(sums, 0, sizeof(double) * 2 * (size_t)B * G);
#endif
    const int C4 = C / 4, cpg = C / G;
    const bool quad = (C % 4 == 0) && C4 <= 256 && (C4 & (C4 - 1)) == 0 && (cpg % 4 == 0) && cpg <= 32 && G <= 64 &&
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx) |
                        reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
    if (quad) {
        const int nwalk = 256 / C4;
        long long chunks = (HW + 16LL * nwalk - 1) / (16LL * nwalk);                 // >= 16 pixels per walker
        const long long cap = (148LL * 8 + B - 1) / B;
        if (chunks > cap) chunks = cap;
        TE_LAUNCH(gn_bwd_stats_quad_kernel, dim3((unsigned)chunks, (unsigned)B), 256, stream, x, gy, stats, gamma, beta, eps, swish, HW, C, G, sums, dgamma,
                  dbeta);
        long long blocks = (HW + 4LL * nwalk - 1) / (4LL * nwalk);
        const long long cap2 = (148LL * 16 + B - 1) / B;
        if (blocks > cap2) blocks = cap2;
        TE_LAUNCH(gn_bwd_apply_quad_kernel, dim3((unsigned)blocks, (unsigned)B), 256, stream, x, gy, stats, gamma, beta, eps, swish, HW, C, G, sums, gx);
        GLARE_CHECK_LAUNCH();
        return GLARE_OK;
    }
    long long chunks = (HW + 63) / 64;
    const long long cap = (148LL * 8 + B - 1) / B;
    if (chunks > cap) chunks = cap;
    TE_LAUNCH(gn_bwd_stats_kernel, dim3((unsigned)chunks, (unsigned)B), 256, stream, x, gy, stats, gamma, beta, eps, swish, HW, C, G, sums, dgamma, dbeta);
    long long blocks = (HW * C + 1023) / 1024;
    const long long cap2 = (148LL * 16 + B - 1) / B;
    if (blocks > cap2) blocks = cap2;
    TE_LAUNCH(gn_bwd_apply_kernel, dim3((unsigned)blocks, (unsigned)B), 256, stream, x, gy, stats, gamma, beta, eps, swish, HW, C, G, sums, gx);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// x NHWC [B,H,W,C] -> col [B*Ho*Wo][k*k*C]; k in {1, 3}, stride in {1, 2}, pad = low-side padding (high side: zero fill as far as the taps reach)
GLARE_API int glare_im2col_nhwc_f32(const float* x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, float* col,
                                    cudaStream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (k != 1 && k != 3) || (stride != 1 && stride != 2) || pad < 0 || pad > 1 || Ho <= 0 || Wo <= 0)
        return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !col) return GLARE_ERR_BAD_ARG;
    const long long total = (long long)B * Ho * Wo * k * k * C;
    const long long blocks = (total + 255) / 256;
    TE_LAUNCH(im2col_nhwc_kernel, dim3((unsigned)(blocks < 148LL * 32 ? blocks : 148LL * 32)), 256, stream, x, B, H, W, C, k, stride, pad, Ho, Wo, col);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// P, dP, dS [rows][ld] fp32 (first n_keys columns used): softmax backward of AttnBlock (encoder_decoder.py:181-183: w = softmax(scale * q k))
GLARE_API int glare_attn_softmax_bwd_f32(const float* P, const float* dP, long long rows, long long ld, int n_keys, float scale, float* dS,
                                         cudaStream_t stream) {
    if (rows < 0 || n_keys <= 0 || ld < n_keys) return GLARE_ERR_BAD_ARG;
    if (rows == 0) return GLARE_OK;
    if (!P || !dP || !dS || rows > 0x7fffffffLL) return GLARE_ERR_BAD_ARG;
    TE_LAUNCH(attn_softmax_bwd_kernel, dim3((unsigned)rows), 256, stream, P, dP, ld, n_keys, scale, dS);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// x [P][C] fp32 contiguous, C % 4 == 0 -> out[C] += column sums (out zero-filled by the caller for a plain sum): bias gradients of the
// encoder / decoder convolutions (torch autograd of nn.Conv2d's bias, encoder_decoder.py:88-115)
GLARE_API int glare_colsum_f32(const float* x, long long P, int C, float* out, cudaStream_t stream) {
    if (P < 0 || C <= 0 || (C & 3)) return GLARE_ERR_BAD_ARG;
    if (P == 0) return GLARE_OK;
    if (!x || !out) return GLARE_ERR_BAD_ARG;
    const int col_tiles = (C + 127) / 128;
    long long chunks = (P + 255) / 256;                                  // >= 32 sequential loads per thread
    const long long cap = (148LL * 8 + col_tiles - 1) / col_tiles;
    if (chunks > cap) chunks = cap;
    const long long per = (P + chunks - 1) / chunks;
    chunks = (P + per - 1) / per;
    TE_LAUNCH(colsum_kernel, dim3((unsigned)col_tiles, (unsigned)chunks), 256, stream, x, P, C, per, out);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
