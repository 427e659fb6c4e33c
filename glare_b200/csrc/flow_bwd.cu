// Stage-2 training support for the conditional flow: forward activations of the coupling nets and the backward pass of a FlowStep
// (BASELINE config 4; reference FlowStep.py:75-98 normal_flow + autograd of FlowActNorms.py:48-100, Permutations.py:21-59,
// FlowAffineCouplingsAblation.py:50-151, flow.py:13-70).  Every formula follows oracle/flow_backward.py, the CPU specification that
// tests/test_oracle.py checks against torch autograd and the reference's own gradients (tests/golden/stage2.npz).
//
// STATUS: verified on the CPU (this source through tests/cuda_emu, tests/test_flow_train_cpu.py) and green on B200 since round 2
// (tests/flow_train_gpu_check.py in a child process of tests/test_zz_flow_train_gpu.py; profiles/r70_train_check.log).  Nothing on the
// inference path calls into this file.
//
// Training shapes are small (latent 80x80, batch 4: P = 25 600 pixels, 24 coupling steps x 2 nets x 9 552 parameters), so the pass is a
// sequence of simple kernels over NHWC-flattened [P][C] fp32 buffers -- one thread per pixel, parameters of the packed net block
// (flow.cu NET_* layout) read through the read-only cache or staged in shared memory.  Weight gradients: the tensor-core path of
// csrc/train_wgrad.cu (glare_b200/flow_train.py CudaKernels.wgrad), the skinny kernel at the end of this file for the 9-column ones, the
// split-K fp32 GEMM of dcn_bwd.cu for anything else.  The hoisted 64 -> 3072 conv over ft (DESIGN.md "flow") gets its data gradient from
// the tensor-core conv path and its weight gradient from the same tensor-core path.
#ifdef GLARE_CUDA_EMU
#include "cuda_emu.h"        // tests/cuda_emu: the same kernel source executed on the host (CPU check of indexing and arithmetic)
#define FB_LAUNCH(kern, grid, block, stream, ...) glare_emu::launch(kern, dim3(grid), dim3(block), __VA_ARGS__)
#else
#include "common.cuh"
#define FB_LAUNCH(kern, grid, block, stream, ...) kern<<<grid, block, 0, stream>>>(__VA_ARGS__)
#endif

namespace glare {

// packed per-net parameter block, identical to flow.cu / glare_b200/flow.py
constexpr int FB_C = 64;
constexpr int FB_W1Z = 0, FB_B1 = 576, FB_S1 = 640, FB_W2T = 704, FB_B2 = 4800, FB_S2 = 4864, FB_W3 = 4928, FB_B3 = 9536, FB_S3 = 9544;
constexpr float FB_EPS = 0.0001f;             // affine_eps, FlowAffineCouplingsAblation.py:31

__device__ __forceinline__ void fb_decode(long long p, int h, int w, int& b, int& y, int& x) {
    const long long hw = (long long)h * w;
    b = (int)(p / hw);
    const int r = (int)(p - (long long)b * hw);
    y = r / w;
    x = r - y * w;
}

// ---- forward of one coupling net, activations kept -----------------------------------------------------------------------------
// h1[p][c] = relu((pre[p * pre_ld + c] + sum_t W1z[c][t] z1(p + d_t) + b1[c]) * s1[c])      (z1: plane of row stride z1_ld, or null)
// The post-ReLU value is all the backward needs: the mask is h > 0 and the ActNorm logs gradient sum(g_n * n) only sees n where n > 0.
__global__ void __launch_bounds__(128) flow_tr_net1_kernel(const float* __restrict__ pre, long long pre_ld, const float* __restrict__ z1, long long z1_ld,
                                                           const float* __restrict__ net, long long P, int h, int w, float* __restrict__ n1) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int b, y, x;
    fb_decode(p, h, w, b, y, x);
    float zc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        zc[t] = (z1 != nullptr && yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(z1 + (((long long)b * h + yy) * w + xx) * z1_ld) : 0.f;
    }
    const float* pr = pre + p * pre_ld;
    float* o = n1 + p * FB_C;
    for (int c = 0; c < FB_C; ++c) {
        float v = __ldg(pr + c);
        if (z1 != nullptr) {
#pragma unroll
            for (int t = 0; t < 9; ++t) v = fmaf(__ldg(net + FB_W1Z + c * 9 + t), zc[t], v);
        }
        o[c] = fmaxf((v + __ldg(net + FB_B1 + c)) * __ldg(net + FB_S1 + c), 0.f);
    }
}

// h2[p][o] = relu((sum_i W2T[i][o] h1[p][i] + b2[o]) * s2[o])
__global__ void __launch_bounds__(128) flow_tr_net2_kernel(const float* __restrict__ n1, const float* __restrict__ net, long long P,
                                                           float* __restrict__ n2) {
    __shared__ float s_w[FB_C * FB_C];
    for (int i = threadIdx.x; i < FB_C * FB_C; i += blockDim.x) s_w[i] = __ldg(net + FB_W2T + i);
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float acc[FB_C];
#pragma unroll
    for (int o = 0; o < FB_C; ++o) acc[o] = 0.f;
    const float* a = n1 + p * FB_C;
    for (int i = 0; i < FB_C; ++i) {
        const float xv = __ldg(a + i);
#pragma unroll
        for (int o = 0; o < FB_C; ++o) acc[o] = fmaf(s_w[i * FB_C + o], xv, acc[o]);
    }
    float* out = n2 + p * FB_C;
#pragma unroll
    for (int o = 0; o < FB_C; ++o) out[o] = fmaxf((acc[o] + __ldg(net + FB_B2 + o)) * __ldg(net + FB_S2 + o), 0.f);
}

// hout[p][j] = (sum_{c,t} W3[c][t][j] h2(p + d_t)[c] + b3[j]) * s3[j],  j < 8 (columns >= nout are zero: zero weights / bias)
__global__ void __launch_bounds__(128) flow_tr_net3_kernel(const float* __restrict__ n2, const float* __restrict__ net, long long P, int h, int w,
                                                           float* __restrict__ hout) {
    __shared__ __align__(16) float s_w[FB_C * 72];
    for (int i = threadIdx.x; i < FB_C * 72; i += blockDim.x) s_w[i] = __ldg(net + FB_W3 + i);
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int b, y, x;
    fb_decode(p, h, w, b, y, x);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
        const float* a = n2 + (((long long)b * h + yy) * w + xx) * FB_C;
        for (int c = 0; c < FB_C; ++c) {
            const float xv = __ldg(a + c);
            const float4 wa = *reinterpret_cast<const float4*>(s_w + c * 72 + t * 8), wb = *reinterpret_cast<const float4*>(s_w + c * 72 + t * 8 + 4);
            o[0] = fmaf(wa.x, xv, o[0]); o[1] = fmaf(wa.y, xv, o[1]); o[2] = fmaf(wa.z, xv, o[2]); o[3] = fmaf(wa.w, xv, o[3]);
            o[4] = fmaf(wb.x, xv, o[4]); o[5] = fmaf(wb.y, xv, o[5]); o[6] = fmaf(wb.z, xv, o[6]); o[7] = fmaf(wb.w, xv, o[7]);
        }
    }
    float* out = hout + p * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = (o[j] + __ldg(net + FB_B3 + j)) * __ldg(net + FB_S3 + j);
}

__device__ __forceinline__ float fb_scale(float hraw) { return 1.0f / (1.0f + expf(-(hraw + 2.0f))) + FB_EPS; }

// ActNorm + invertible 1x1 + feature affine of a step (FlowStep.py:75-98 up to the self coupling): z_in NCHW [B,3,h,w], pw = forward
// pointwise block (flow.py pack_pointwise: M[9], bias[3], exp(logs)[3]), hF [P][8] or null (noCoupling) -> t, u, v [P][4] (3 used)
__global__ void __launch_bounds__(256) flow_tr_point_fwd_kernel(const float* __restrict__ z_in, const float* __restrict__ pw, const float* __restrict__ hF,
                                                                long long P, long long hw, float* __restrict__ t_out, float* __restrict__ u_out,
                                                                float* __restrict__ v_out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const long long b = p / hw, q = p - b * hw;
    const float* zb = z_in + b * 3 * hw + q;
    const float t0 = (zb[0] + __ldg(pw + 9)) * __ldg(pw + 12), t1 = (zb[hw] + __ldg(pw + 10)) * __ldg(pw + 13),
                t2 = (zb[2 * hw] + __ldg(pw + 11)) * __ldg(pw + 14);
    float u[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) u[o] = __ldg(pw + 3 * o) * t0 + __ldg(pw + 3 * o + 1) * t1 + __ldg(pw + 3 * o + 2) * t2;
    t_out[p * 4] = t0; t_out[p * 4 + 1] = t1; t_out[p * 4 + 2] = t2; t_out[p * 4 + 3] = 0.f;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        u_out[p * 4 + o] = u[o];
        v_out[p * 4 + o] = hF ? (u[o] + __ldg(hF + p * 8 + 2 * o)) * fb_scale(__ldg(hF + p * 8 + 2 * o + 1)) : u[o];
    }
    u_out[p * 4 + 3] = 0.f;
    v_out[p * 4 + 3] = 0.f;
}

// ---- backward ------------------------------------------------------------------------------------------------------------------------
// (x + shift) * scale, scale = sigmoid(hraw + 2) + eps, logdet += log scale:  gradient wrt (shift, hraw) given g_y and g_ld
__device__ __forceinline__ void fb_affine_bwd(float x, float shift, float hraw, float g_y, float g_ld, float& g_x, float& g_shift, float& g_hraw) {
    const float sc = fb_scale(hraw), s = sc - FB_EPS;
    g_x = g_y * sc;
    g_shift = g_y * sc;
    g_hraw = (g_y * (x + shift) + g_ld / sc) * s * (1.0f - s);
}

// self coupling backward: g_out NCHW [B,3,h,w], v [P][4], hA [P][8] -> g_hA [P][8] (4 used, rest 0), g_v [P][4] (g_v[0] = g_out[0] for now)
__global__ void __launch_bounds__(256) flow_tr_coupling_bwd_a_kernel(const float* __restrict__ g_out, const float* __restrict__ v,
                                                                     const float* __restrict__ hA, float g_ld, long long P, long long hw,
                                                                     float* __restrict__ g_hA, float* __restrict__ g_v) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const long long b = p / hw, q = p - b * hw;
    const float* gb = g_out + b * 3 * hw + q;
    float gh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gh[j] = 0.f;
    float gv[4] = {gb[0], 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j)
        fb_affine_bwd(__ldg(v + p * 4 + 1 + j), __ldg(hA + p * 8 + 2 * j), __ldg(hA + p * 8 + 2 * j + 1), gb[(1 + j) * hw], g_ld, gv[1 + j], gh[2 * j],
                      gh[2 * j + 1]);
#pragma unroll
    for (int j = 0; j < 8; ++j) g_hA[p * 8 + j] = gh[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) g_v[p * 4 + j] = gv[j];
}

// feature affine backward: g_v [P][4] (+ g_z1 [P] into channel 0), u [P][4], hF [P][8] -> g_hF [P][8] (6 used), g_u [P][4]
__global__ void __launch_bounds__(256) flow_tr_coupling_bwd_f_kernel(const float* __restrict__ g_v, const float* __restrict__ g_z1,
                                                                     const float* __restrict__ u, const float* __restrict__ hF, float g_ld, long long P,
                                                                     float* __restrict__ g_hF, float* __restrict__ g_u) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float gh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gu[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float gy = __ldg(g_v + p * 4 + j) + (j == 0 ? __ldg(g_z1 + p) : 0.f);
        fb_affine_bwd(__ldg(u + p * 4 + j), __ldg(hF + p * 8 + 2 * j), __ldg(hF + p * 8 + 2 * j + 1), gy, g_ld, gu[j], gh[2 * j], gh[2 * j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g_hF[p * 8 + j] = gh[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) g_u[p * 4 + j] = gu[j];
}

// Conv2dZeros backward, data part: g_a3[p][j] = g_h[p][j] * s3[j] (written, [P][8]); used by the stencil below and by the weight GEMM
__global__ void __launch_bounds__(256) flow_tr_scale8_kernel(const float* __restrict__ g_h, const float* __restrict__ net, long long P,
                                                             float* __restrict__ g_a3) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * 8) return;
    g_a3[i] = g_h[i] * __ldg(net + FB_S3 + (int)(i & 7));
}

// g_n2[p][c] = (sum_t sum_j W3[c][t][j] g_a3(p - d_t)[j]) * (h2[p][c] > 0)          (transpose of the 3x3 conv, then the ReLU mask)
__global__ void __launch_bounds__(128) flow_tr_net3_dgrad_kernel(const float* __restrict__ g_a3, const float* __restrict__ n2,
                                                                 const float* __restrict__ net, long long P, int h, int w, float* __restrict__ g_n2) {
    __shared__ __align__(16) float s_w[FB_C * 72];
    for (int i = threadIdx.x; i < FB_C * 72; i += blockDim.x) s_w[i] = __ldg(net + FB_W3 + i);
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int b, y, x;
    fb_decode(p, h, w, b, y, x);
    float g[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
        const bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
        const float* gp = g_a3 + (((long long)b * h + yy) * w + xx) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) g[t][j] = in ? __ldg(gp + j) : 0.f;
    }
    for (int c = 0; c < FB_C; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(s_w[c * 72 + t * 8 + j], g[t][j], acc);
        g_n2[p * FB_C + c] = __ldg(n2 + p * FB_C + c) > 0.f ? acc : 0.f;
    }
}

// g_a2 = g_n2 * s2 (written); g_n1[p][i] = (sum_o W2T[i][o] g_a2[o]) * (h1[p][i] > 0); g_a1 = g_n1 * s1 written densely [P][64] and, when
// g_pre != null, into the pre-activation gradient tensor at row stride pre_ld (the operand of the hoisted conv's backward)
__global__ void __launch_bounds__(128) flow_tr_net2_dgrad_kernel(const float* __restrict__ g_n2, const float* __restrict__ n1,
                                                                 const float* __restrict__ net, long long P, float* __restrict__ g_a2,
                                                                 float* __restrict__ g_n1, float* __restrict__ g_a1, float* __restrict__ g_pre,
                                                                 long long pre_ld) {
    __shared__ float s_w[FB_C * FB_C];
    for (int i = threadIdx.x; i < FB_C * FB_C; i += blockDim.x) s_w[i] = __ldg(net + FB_W2T + i);
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float ga[FB_C];
#pragma unroll
    for (int o = 0; o < FB_C; ++o) {
        ga[o] = __ldg(g_n2 + p * FB_C + o) * __ldg(net + FB_S2 + o);
        g_a2[p * FB_C + o] = ga[o];
    }
    for (int i = 0; i < FB_C; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int o = 0; o < FB_C; ++o) acc = fmaf(s_w[i * FB_C + o], ga[o], acc);
        const float gn = __ldg(n1 + p * FB_C + i) > 0.f ? acc : 0.f;
        const float g1 = gn * __ldg(net + FB_S1 + i);
        g_n1[p * FB_C + i] = gn;
        g_a1[p * FB_C + i] = g1;
        if (g_pre != nullptr) g_pre[p * pre_ld + i] = g1;
    }
}

// g_z1[p] = sum_t sum_c W1z[c][t] g_a1(p - d_t)[c]
__global__ void __launch_bounds__(128) flow_tr_net1_zgrad_kernel(const float* __restrict__ g_a1, const float* __restrict__ net, long long P, int h, int w,
                                                                 float* __restrict__ g_z1) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int b, y, x;
    fb_decode(p, h, w, b, y, x);
    float acc = 0.f;
    for (int t = 0; t < 9; ++t) {
        const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
        if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
        const float* gp = g_a1 + (((long long)b * h + yy) * w + xx) * FB_C;
        for (int c = 0; c < FB_C; ++c) acc = fmaf(__ldg(net + FB_W1Z + c * 9 + t), __ldg(gp + c), acc);
    }
    g_z1[p] = acc;
}

// invertible 1x1 + ActNorm backward: g_u [P][4], t [P][4], pw (forward block) -> g_z_in NCHW; sums[0..8] += sum_p g_u[o] t[i] (o*3+i),
// sums[9..11] += sum_p g_t[i] t[i] (logs), sums[12..14] += sum_p g_z[i] (bias)
__global__ void __launch_bounds__(256) flow_tr_point_bwd_kernel(const float* __restrict__ g_u, const float* __restrict__ t, const float* __restrict__ pw,
                                                                long long P, long long hw, float* __restrict__ g_z, float* __restrict__ sums) {
    __shared__ float s_red[8][15];
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float acc[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i] = 0.f;
    if (p < P) {
        const long long b = p / hw, q = p - b * hw;
        const float gu[3] = {__ldg(g_u + p * 4), __ldg(g_u + p * 4 + 1), __ldg(g_u + p * 4 + 2)};
        const float tv[3] = {__ldg(t + p * 4), __ldg(t + p * 4 + 1), __ldg(t + p * 4 + 2)};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float gt = __ldg(pw + i) * gu[0] + __ldg(pw + 3 + i) * gu[1] + __ldg(pw + 6 + i) * gu[2];       // W^T g_u
            const float gz = gt * __ldg(pw + 12 + i);
            g_z[b * 3 * hw + (long long)i * hw + q] = gz;
            acc[9 + i] = gt * tv[i];
            acc[12 + i] = gz;
#pragma unroll
            for (int o = 0; o < 3; ++o) acc[o * 3 + i] = gu[o] * tv[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 15; ++i) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], s);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < 15; ++i) s_red[threadIdx.x >> 5][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 15) {
        float v = 0.f;
        for (int k = 0; k < 8; ++k) v += s_red[k][threadIdx.x];
        atomicAdd(sums + threadIdx.x, v);
    }
}

// col[p][t * C + c] = f(x(p + d_t)[c]) (zero outside the image), f = ReLU when relu != 0: the im2col operand of a 3x3 weight gradient
__global__ void __launch_bounds__(256) flow_tr_im2col3x3_kernel(const float* __restrict__ x, long long ldx, int C, int relu, long long P, int h, int w,
                                                                float* __restrict__ col) {
    const long long total = P * 9 * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long r = i / C;
        const int t = (int)(r % 9);
        const long long p = r / 9;
        int b, y, xx;
        fb_decode(p, h, w, b, y, xx);
        const int yy = y + t / 3 - 1, xq = xx + t % 3 - 1;
        float v = 0.f;
        if (yy >= 0 && yy < h && xq >= 0 && xq < w) {
            v = __ldg(x + (((long long)b * h + yy) * w + xq) * ldx + c);
            if (relu) v = fmaxf(v, 0.f);
        }
        col[i] = v;
    }
}

// out[c] += sum_p a[p * lda + c] * (b ? b[p * ldb + c] : 1)      (bias / logs gradients); grid (chunks), C <= 64
__global__ void __launch_bounds__(256) flow_tr_colsum_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b, long long ldb, int C,
                                                             long long P, long long p_per_cta, float* __restrict__ out) {
    __shared__ float s_acc[64];
    if (threadIdx.x < 64) s_acc[threadIdx.x] = 0.f;
    __syncthreads();
    const int c = threadIdx.x % C, r0 = threadIdx.x / C, rows = blockDim.x / C;
    const long long p0 = (long long)blockIdx.x * p_per_cta, p1 = (p0 + p_per_cta < P) ? p0 + p_per_cta : P;
    float acc = 0.f;
    if (r0 < rows)
        for (long long p = p0 + r0; p < p1; p += rows) acc = fmaf(__ldg(a + p * lda + c), b ? __ldg(b + p * ldb + c) : 1.0f, acc);
    if (r0 < rows) atomicAdd(&s_acc[c], acc);
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(out + threadIdx.x, s_acc[threadIdx.x]);
}

// out[m][n] += sum_p a[p][m] * b[p][n] for a SKINNY left operand (M <= 32 columns): the weight gradient of the z1 input channel of a coupling
// net's first 3x3 conv (a = the 9 im2col columns of z1, b = the 64-channel pre-activation gradient).  256 threads = (256 / N) pixel walkers x N
// columns; every thread keeps M accumulators, reads its b element once per pixel (coalesced rows) and the M a-values of the pixel through the
// read-only cache (the same address for a whole walker).  The 128 x 128-tile split-K GEMM spent 0.29 ms on each of these 24 calls per step.
__global__ void __launch_bounds__(256) gemm_tn_skinny_kernel(const float* __restrict__ a, const float* __restrict__ b, long long P, int M, int N,
                                                             long long p_per_cta, float* __restrict__ out) {
    __shared__ float red[32][256];
    const int n = threadIdx.x % N, w = threadIdx.x / N, nwalk = 256 / N;
    const long long p0 = (long long)blockIdx.x * p_per_cta, p1 = (p0 + p_per_cta < P) ? p0 + p_per_cta : P;
    float acc[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) acc[m] = 0.f;
    if (w < nwalk)
        for (long long p = p0 + w; p < p1; p += nwalk) {
            const float bv = __ldg(b + p * N + n);
#pragma unroll
            for (int m = 0; m < 32; ++m)
                if (m < M) acc[m] = fmaf(__ldg(a + p * M + m), bv, acc[m]);
        }
#pragma unroll
    for (int m = 0; m < 32; ++m) red[m][threadIdx.x] = acc[m];
    __syncthreads();
    // thread t < M * N: element (m, n2) summed over the walkers
    for (int e = threadIdx.x; e < M * N; e += 256) {
        const int m = e / N, n2 = e - m * N;
        float t = 0.f;
        for (int r = 0; r < nwalk; ++r) t += red[m][r * N + n2];
        atomicAdd(out + e, t);
    }
}

}  // namespace glare

using namespace glare;

#define FB_GRID(n, t) (unsigned)(((n) + (t)-1) / (t))

// Forward of one coupling net with the activations its backward needs: pre = pre-activation planes [P] rows of stride pre_ld (the hoisted
// conv's output), z1 = first latent channel, element stride z1_ld between pixels, or NULL (NN_F); net = packed block; n1, n2 [P][64] = the
// post-ReLU hidden activations h1, h2; hout [P][8] = net output (columns >= 4 / 6 are zero).
GLARE_API int glare_flow_train_net_fwd_f32(const float* pre, long long pre_ld, const float* z1, long long z1_ld, const float* net, int B, int h, int w,
                                           float* n1, float* n2, float* hout, cudaStream_t stream) {
    if (B < 0 || h <= 0 || w <= 0 || pre_ld < FB_C) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!pre || !net || !n1 || !n2 || !hout) return GLARE_ERR_BAD_ARG;
    const long long P = (long long)B * h * w;
    FB_LAUNCH(flow_tr_net1_kernel, FB_GRID(P, 128), 128, stream, pre, pre_ld, z1, z1_ld, net, P, h, w, n1);
    FB_LAUNCH(flow_tr_net2_kernel, FB_GRID(P, 128), 128, stream, n1, net, P, n2);
    FB_LAUNCH(flow_tr_net3_kernel, FB_GRID(P, 128), 128, stream, n2, net, P, h, w, hout);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// ActNorm + invertible 1x1 (+ feature affine when hF != NULL) of a step: z_in NCHW [B,3,h,w] -> t, u, v [P][4]
GLARE_API int glare_flow_train_point_fwd_f32(const float* z_in, const float* pw_fwd, const float* hF, int B, int h, int w, float* t, float* u, float* v,
                                             cudaStream_t stream) {
    if (B < 0 || h <= 0 || w <= 0) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!z_in || !pw_fwd || !t || !u || !v) return GLARE_ERR_BAD_ARG;
    const long long P = (long long)B * h * w;
    FB_LAUNCH(flow_tr_point_fwd_kernel, FB_GRID(P, 256), 256, stream, z_in, pw_fwd, hF, P, (long long)h * w, t, u, v);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// which = 0: self coupling backward (g_out NCHW, v, hA -> g_h [P][8], g_v [P][4]); which = 1: feature affine backward (g_v, g_z1, u, hF ->
// g_h [P][8], g_u [P][4]).  g_ld = dL/dlogdet of a sample (the same for every sample of a mean-reduced objective).
GLARE_API int glare_flow_train_coupling_bwd_f32(int which, const float* g_in, const float* g_z1, const float* x, const float* hraw, float g_ld, int B,
                                                int h, int w, float* g_h, float* g_x, cudaStream_t stream) {
    if (B < 0 || h <= 0 || w <= 0 || (which != 0 && which != 1)) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!g_in || !x || !hraw || !g_h || !g_x || (which == 1 && !g_z1)) return GLARE_ERR_BAD_ARG;
    const long long P = (long long)B * h * w;
    if (which == 0) FB_LAUNCH(flow_tr_coupling_bwd_a_kernel, FB_GRID(P, 256), 256, stream, g_in, x, hraw, g_ld, P, (long long)h * w, g_h, g_x);
    else FB_LAUNCH(flow_tr_coupling_bwd_f_kernel, FB_GRID(P, 256), 256, stream, g_in, g_z1, x, hraw, g_ld, P, g_h, g_x);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// Data half of a coupling net's backward: g_h [P][8] (gradient of the net output) with the saved h1, h2 (passed as n1, n2) -> g_a3 [P][8], g_n2, g_a2, g_n1,
// g_a1 [P][64] (g_a1 also scattered to g_pre at row stride pre_ld when g_pre != NULL) and, when g_z1 != NULL, the gradient of the z1 plane.
// The parameter gradients follow from these with glare_dcnv2_bwd_weight_f32 / glare_flow_train_colsum_f32 (glare_b200/flow_train.py).
GLARE_API int glare_flow_train_net_bwd_f32(const float* g_h, const float* n1, const float* n2, const float* net, int B, int h, int w, float* g_a3,
                                           float* g_n2, float* g_a2, float* g_n1, float* g_a1, float* g_pre, long long pre_ld, float* g_z1,
                                           cudaStream_t stream) {
    if (B < 0 || h <= 0 || w <= 0 || (g_pre && pre_ld < FB_C)) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!g_h || !n1 || !n2 || !net || !g_a3 || !g_n2 || !g_a2 || !g_n1 || !g_a1) return GLARE_ERR_BAD_ARG;
    const long long P = (long long)B * h * w;
    FB_LAUNCH(flow_tr_scale8_kernel, FB_GRID(P * 8, 256), 256, stream, g_h, net, P, g_a3);
    FB_LAUNCH(flow_tr_net3_dgrad_kernel, FB_GRID(P, 128), 128, stream, g_a3, n2, net, P, h, w, g_n2);
    FB_LAUNCH(flow_tr_net2_dgrad_kernel, FB_GRID(P, 128), 128, stream, g_n2, n1, net, P, g_a2, g_n1, g_a1, g_pre, pre_ld);
    if (g_z1) FB_LAUNCH(flow_tr_net1_zgrad_kernel, FB_GRID(P, 128), 128, stream, g_a1, net, P, h, w, g_z1);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// invertible 1x1 + ActNorm backward: g_u, t [P][4] -> g_z NCHW [B,3,h,w]; sums[15] += (dW data term [9], dlogs data term [3], dbias [3])
GLARE_API int glare_flow_train_point_bwd_f32(const float* g_u, const float* t, const float* pw_fwd, int B, int h, int w, float* g_z, float* sums,
                                             cudaStream_t stream) {
    if (B < 0 || h <= 0 || w <= 0) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!g_u || !t || !pw_fwd || !g_z || !sums) return GLARE_ERR_BAD_ARG;
    const long long P = (long long)B * h * w;
    FB_LAUNCH(flow_tr_point_bwd_kernel, FB_GRID(P, 256), 256, stream, g_u, t, pw_fwd, P, (long long)h * w, g_z, sums);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// col [P][9*C] (tap-major) = im2col of x [P] rows of stride ldx, optionally through a ReLU
GLARE_API int glare_flow_train_im2col3x3_f32(const float* x, long long ldx, int C, int relu, int B, int h, int w, float* col, cudaStream_t stream) {
    if (B < 0 || h <= 0 || w <= 0 || C <= 0 || ldx < C) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !col) return GLARE_ERR_BAD_ARG;
    const long long P = (long long)B * h * w, total = P * 9 * C;
    const long long blocks = (total + 255) / 256;
    FB_LAUNCH(flow_tr_im2col3x3_kernel, (unsigned)(blocks < 148LL * 32 ? blocks : 148LL * 32), 256, stream, x, ldx, C, relu, P, h, w, col);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// out[c] += sum_p a[p * lda + c] * (b ? b[p * ldb + c] : 1), c < C <= 64 with 256 % C == 0
GLARE_API int glare_flow_train_colsum_f32(const float* a, long long lda, const float* b, long long ldb, int C, long long P, float* out,
                                          cudaStream_t stream) {
    if (P < 0 || C <= 0 || C > 64 || 256 % C != 0 || lda < C || (b && ldb < C)) return GLARE_ERR_BAD_ARG;
    if (P == 0) return GLARE_OK;
    if (!a || !out) return GLARE_ERR_BAD_ARG;
    // >= one CTA per 128 rows (32 sequential loads per thread): round 1's 1024 rows per CTA left 25 CTAs for the 25 600 latent pixels of a
    // training batch, 80 us per call and 288 calls per step (profiles/r50_train_probe_kernel_breakdown.txt)
    long long chunks = (P + 127) / 128;
    if (chunks > 148 * 8) chunks = 148 * 8;
    const long long per = (P + chunks - 1) / chunks;
    chunks = (P + per - 1) / per;
    FB_LAUNCH(flow_tr_colsum_kernel, (unsigned)chunks, 256, stream, a, lda, b, ldb, C, P, per, out);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// out [M][N] += a [P][M]^T b [P][N], fp32; M <= 32, N <= 256 with 256 % N == 0 (out accumulates: zero-fill for a plain product)
GLARE_API int glare_gemm_tn_skinny_f32(const float* a, const float* b, long long P, int M, int N, float* out, cudaStream_t stream) {
    if (P < 0 || M <= 0 || M > 32 || N <= 0 || N > 256 || 256 % N != 0) return GLARE_ERR_BAD_ARG;
    if (P == 0) return GLARE_OK;
    if (!a || !b || !out) return GLARE_ERR_BAD_ARG;
    const int nwalk = 256 / N;
    long long chunks = (P + 16LL * nwalk - 1) / (16LL * nwalk);          // >= 16 pixels per walker
    if (chunks > 148 * 4) chunks = 148 * 4;
    const long long per = (P + chunks - 1) / chunks;
    chunks = (P + per - 1) / per;
    FB_LAUNCH(gemm_tn_skinny_kernel, (unsigned)chunks, 256, stream, a, b, P, M, N, per, out);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
