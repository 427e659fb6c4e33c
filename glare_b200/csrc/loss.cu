// Stage-3 loss kernels: one SSIM level of the reference's MS-SSIM (code/models/modules/pytorch_msssim/__init__.py:20-68 `ssim`, called per
// level by `msssim` :71-97) -- value and gradient with respect to the first image.
//
//   mu1 = G*x, mu2 = G*y, e11 = G*(x x), e22 = G*(y y), e12 = G*(x y)            ('valid' Gaussian window, 11 x 11, sigma 1.5, per plane)
//   v1 = 2 (e12 - mu1 mu2) + C2,  v2 = (e11 - mu1^2) + (e22 - mu2^2) + C2,  cs_map = v1 / v2
//   ssim_map = (2 mu1 mu2 + C1) v1 / ((mu1^2 + mu2^2 + C1) v2);   cs = mean(cs_map), ssim = mean(ssim_map)
//
// forward : per-CTA partial sums of cs_map and ssim_map (summed in fp64 by the caller: a fixed partition, so deterministic).
// backward: given d(objective)/d(cs) and /d(ssim) (device scalars, no host sync), the three window-level gradient maps
//           g_mu = dL/dmu1, g_e11 = dL/de11, g_e12 = dL/de12, then  dL/dx = G^T g_mu + 2 x G^T g_e11 + y G^T g_e12  (G^T: the adjoint of the
//           'valid' correlation = 'full' correlation), plus a quarter of the coarser level's gradient (adjoint of avg_pool2d(2)).
// HBM-bound by construction (3-channel images, < 1 MB per level): the point is parity and keeping the training step on the device.
// The same source runs on the host through tests/cuda_emu (tests/test_losses_cpu.py).
#ifdef GLARE_CUDA_EMU
#include "cuda_emu.h"
#define LS_LAUNCH(kern, grid, block, stream, ...) glare_emu::launch(kern, grid, dim3(block), __VA_ARGS__)
#else
#include "common.cuh"
#define LS_LAUNCH(kern, grid, block, stream, ...) kern<<<grid, block, 0, stream>>>(__VA_ARGS__)
#endif

namespace {

constexpr int T = 16;          // outputs per CTA edge; 256 threads, thread t -> (t / 16, t % 16)
#define TY ((int)(threadIdx.x >> 4))
#define TX ((int)(threadIdx.x & 15))
constexpr int MAXW = 11;       // window_size of the reference (real_size = min(11, H, W))

struct Win {
    float g[MAXW];
};

struct Moments {
    float mu1, mu2, e11, e22, e12;
};

// the (T + ws - 1)^2 input patch of one plane, zero outside the image
__device__ __forceinline__ void load_patch(const float* __restrict__ img, int H, int W, int y0, int x0, int ws, float* sm) {
    const int P = T + ws - 1;
    for (int i = threadIdx.x; i < P * P; i += T * T) {
        const int yy = y0 + i / P, xx = x0 + i % P;
        sm[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(long long)yy * W + xx] : 0.f;
    }
}

__device__ __forceinline__ Moments moments(const float* sx, const float* sy, const Win& w, int ws) {
    const int P = T + ws - 1;
    Moments m = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < ws; ++i) {
        for (int j = 0; j < ws; ++j) {
            const float g = w.g[i] * w.g[j];                       // the reference's 2-D window is the outer product of the 1-D one (:14-17)
            const float a = sx[(TY + i) * P + TX + j], b = sy[(TY + i) * P + TX + j];
            m.mu1 = fmaf(g, a, m.mu1);
            m.mu2 = fmaf(g, b, m.mu2);
            m.e11 = fmaf(g, a * a, m.e11);
            m.e22 = fmaf(g, b * b, m.e22);
            m.e12 = fmaf(g, a * b, m.e12);
        }
    }
    return m;
}

__global__ void __launch_bounds__(T* T) ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, int ws, Win w,
                                                          float C1, float C2, float* __restrict__ part) {
    __shared__ float sx[(T + MAXW - 1) * (T + MAXW - 1)], sy[(T + MAXW - 1) * (T + MAXW - 1)];
    __shared__ float red[2][T * T / 32];
    const int Ho = H - ws + 1, Wo = W - ws + 1;
    const long long plane = blockIdx.z;
    const int y0 = blockIdx.y * T, x0 = blockIdx.x * T;
    load_patch(x + plane * H * W, H, W, y0, x0, ws, sx);
    load_patch(y + plane * H * W, H, W, y0, x0, ws, sy);
    __syncthreads();
    float cs = 0.f, ss = 0.f;
    if (y0 + TY < Ho && x0 + TX < Wo) {
        const Moments m = moments(sx, sy, w, ws);
        const float s11 = m.e11 - m.mu1 * m.mu1, s22 = m.e22 - m.mu2 * m.mu2, s12 = m.e12 - m.mu1 * m.mu2;
        const float v1 = 2.0f * s12 + C2, v2 = s11 + s22 + C2;
        cs = v1 / v2;
        ss = ((2.0f * m.mu1 * m.mu2 + C1) * v1) / ((m.mu1 * m.mu1 + m.mu2 * m.mu2 + C1) * v2);
    }
    for (int o = 16; o; o >>= 1) {
        cs += __shfl_xor_sync(0xffffffffu, cs, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    const int tid = threadIdx.x;
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = cs;
        red[1][tid >> 5] = ss;
    }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < T * T / 32; ++i) {
            a += red[0][i];
            b += red[1][i];
        }
        const long long blk = (plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        part[2 * blk] = a;
        part[2 * blk + 1] = b;
    }
}

// coef[0] = dL/d(sum of cs_map), coef[1] = dL/d(sum of ssim_map)   (the caller divides by the number of window positions)
__global__ void __launch_bounds__(T* T) ssim_bwd_maps_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, int ws, Win w,
                                                               float C1, float C2, const float* __restrict__ coef, float* __restrict__ g_mu,
                                                               float* __restrict__ g_e11, float* __restrict__ g_e12) {
    __shared__ float sx[(T + MAXW - 1) * (T + MAXW - 1)], sy[(T + MAXW - 1) * (T + MAXW - 1)];
    const int Ho = H - ws + 1, Wo = W - ws + 1;
    const long long plane = blockIdx.z;
    const int y0 = blockIdx.y * T, x0 = blockIdx.x * T;
    load_patch(x + plane * H * W, H, W, y0, x0, ws, sx);
    load_patch(y + plane * H * W, H, W, y0, x0, ws, sy);
    __syncthreads();
    const int oy = y0 + TY, ox = x0 + TX;
    if (oy >= Ho || ox >= Wo) return;
    const Moments m = moments(sx, sy, w, ws);
    const float a_cs = coef[0], a_ss = coef[1];
    const float s11 = m.e11 - m.mu1 * m.mu1, s22 = m.e22 - m.mu2 * m.mu2, s12 = m.e12 - m.mu1 * m.mu2;
    const float v1 = 2.0f * s12 + C2, v2 = s11 + s22 + C2;
    const float A = 2.0f * m.mu1 * m.mu2 + C1, Bq = m.mu1 * m.mu1 + m.mu2 * m.mu2 + C1;
    const float cs = v1 / v2, lum = A / Bq;
    // d cs_map: through v1 (dv1/dmu1 = -2 mu2, dv1/de12 = 2) and v2 (dv2/dmu1 = -2 mu1, dv2/de11 = 1)
    const float dcs_dmu = (-2.0f * m.mu2) / v2 + cs * (2.0f * m.mu1) / v2;
    const float dcs_de12 = 2.0f / v2, dcs_de11 = -cs / v2;
    const float dlum_dmu = 2.0f * m.mu2 / Bq - lum * 2.0f * m.mu1 / Bq;
    const float k = a_cs + a_ss * lum;                                  // ssim_map = lum * cs_map
    const long long o = (plane * Ho + oy) * Wo + ox;
    g_mu[o] = k * dcs_dmu + a_ss * cs * dlum_dmu;
    g_e11[o] = k * dcs_de11;
    g_e12[o] = k * dcs_de12;
}

// dx[p] = sum_q G[p - q] (g_mu[q] + 2 x[p] g_e11[q] + y[p] g_e12[q])  (+ 0.25 * coarse[p / 2], the adjoint of the next level's avg_pool2d)
__global__ void __launch_bounds__(T* T) ssim_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, int ws, Win w,
                                                                const float* __restrict__ g_mu, const float* __restrict__ g_e11,
                                                                const float* __restrict__ g_e12, const float* __restrict__ coarse,
                                                                float* __restrict__ dx) {
    __shared__ float sm[3][(T + MAXW - 1) * (T + MAXW - 1)];
    const int Ho = H - ws + 1, Wo = W - ws + 1, P = T + ws - 1;
    const long long plane = blockIdx.z;
    const int y0 = blockIdx.y * T, x0 = blockIdx.x * T;
    const float* maps[3] = {g_mu + plane * Ho * Wo, g_e11 + plane * Ho * Wo, g_e12 + plane * Ho * Wo};
    // window positions q in [p - ws + 1, p]
    for (int i = threadIdx.x; i < P * P; i += T * T) {
        const int qy = y0 - (ws - 1) + i / P, qx = x0 - (ws - 1) + i % P;
        const bool in = qy >= 0 && qy < Ho && qx >= 0 && qx < Wo;
        for (int k = 0; k < 3; ++k) sm[k][i] = in ? maps[k][(long long)qy * Wo + qx] : 0.f;
    }
    __syncthreads();
    const int py = y0 + TY, px = x0 + TX;
    if (py >= H || px >= W) return;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int i = 0; i < ws; ++i) {
        for (int j = 0; j < ws; ++j) {
            // q = p - (i, j): patch row (ws - 1 - i) + threadIdx.y
            const int s = (TY + ws - 1 - i) * P + TX + ws - 1 - j;
            const float g = w.g[i] * w.g[j];
            t0 = fmaf(g, sm[0][s], t0);
            t1 = fmaf(g, sm[1][s], t1);
            t2 = fmaf(g, sm[2][s], t2);
        }
    }
    const long long p = (plane * H + py) * W + px;
    float v = t0 + 2.0f * x[p] * t1 + y[p] * t2;
    if (coarse != nullptr) {
        const int Hc = H / 2, Wc = W / 2;
        if ((py >> 1) < Hc && (px >> 1) < Wc) v += 0.25f * coarse[(plane * Hc + (py >> 1)) * Wc + (px >> 1)];
    }
    dx[p] = v;
}

// ---- elementwise / pooling pieces between the convolutions of the stage-3 step (VGG16 features, Upsample, MS-SSIM pyramid) -----------------
// 256 threads, grid-stride; NHWC tensors are walked as float4 over the channels (C % 4 == 0).

__global__ void __launch_bounds__(256) relu_fwd_kernel(const float4* __restrict__ x, long long n4, float4* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = x[i];
        y[i] = make_float4(v.x > 0.f ? v.x : 0.f, v.y > 0.f ? v.y : 0.f, v.z > 0.f ? v.z : 0.f, v.w > 0.f ? v.w : 0.f);
    }
}

// gx = gy where the forward OUTPUT is positive (nn.ReLU backward)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float4* __restrict__ y, const float4* __restrict__ gy, long long n4,
                                                       float4* __restrict__ gx) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = y[i], g = gy[i];
        gx[i] = make_float4(v.x > 0.f ? g.x : 0.f, v.y > 0.f ? g.y : 0.f, v.z > 0.f ? g.z : 0.f, v.w > 0.f ? g.w : 0.f);
    }
}

__device__ __forceinline__ void pool_pick(float v, int k, float& best, unsigned& which) {
    if (v > best || v != v) {            // strictly greater (the first maximum wins, like ATen's max_pool2d scan order); NaN propagates
        best = v;
        which = (unsigned)k;
    }
}

// nn.MaxPool2d(2, 2) on NHWC: y [B][H/2][W/2][C]; idx: one byte per element (0..3 = position in the 2 x 2 window, row-major), packed 4 per word
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float4* __restrict__ x, int H, int W, int C4, long long n_out4,
                                                           float4* __restrict__ y, unsigned* __restrict__ idx) {
    const int Ho = H / 2, Wo = W / 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_out4; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int ox = (int)(r % Wo);
        r /= Wo;
        const int oy = (int)(r % Ho);
        const long long b = r / Ho;
        const long long base = ((b * H + 2 * oy) * W + 2 * ox) * C4 + c;
        float4 best = x[base];
        unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            const float4 v = x[base + ((long long)(k >> 1) * W + (k & 1)) * C4];
            pool_pick(v.x, k, best.x, w0);
            pool_pick(v.y, k, best.y, w1);
            pool_pick(v.z, k, best.z, w2);
            pool_pick(v.w, k, best.w, w3);
        }
        y[i] = best;
        idx[i] = w0 | (w1 << 8) | (w2 << 16) | (w3 << 24);
    }
}

// gx [B][H][W][C]: the pooled gradient routed to the recorded position, zero elsewhere (rows / columns past 2 * (H / 2) are the caller's zeros)
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float4* __restrict__ gy, const unsigned* __restrict__ idx, int H, int W, int C4,
                                                           long long n_out4, float4* __restrict__ gx) {
    const int Ho = H / 2, Wo = W / 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_out4; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int ox = (int)(r % Wo);
        r /= Wo;
        const int oy = (int)(r % Ho);
        const long long b = r / Ho;
        const long long base = ((b * H + 2 * oy) * W + 2 * ox) * C4 + c;
        const float4 g = gy[i];
        const unsigned w = idx[i];
#pragma unroll
        for (unsigned k = 0; k < 4; ++k)
            gx[base + ((long long)(k >> 1) * W + (k & 1)) * C4] =
                make_float4((w & 255u) == k ? g.x : 0.f, ((w >> 8) & 255u) == k ? g.y : 0.f, ((w >> 16) & 255u) == k ? g.z : 0.f, (w >> 24) == k ? g.w : 0.f);
    }
}

// F.avg_pool2d(x, (2, 2)) on [planes][H][W] (the MS-SSIM pyramid, pytorch_msssim/__init__.py:83-84)
__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ x, int H, int W, long long n_out, float* __restrict__ y) {
    const int Ho = H / 2, Wo = W / 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_out; i += (long long)gridDim.x * 256) {
        const int ox = (int)(i % Wo);
        const long long r = i / Wo;
        const int oy = (int)(r % Ho);
        const float* p = x + ((r / Ho) * H + 2 * oy) * W + 2 * ox;
        y[i] = (p[0] + p[1] + p[W] + p[W + 1]) * 0.25f;
    }
}

// nearest x2 on NHWC (Upsample.forward, encoder_decoder.py:46-48): adjoint = 0 -> y [B][2H][2W][C] = x; adjoint = 1 -> y [B][H][W][C] = the 2 x 2
// sums of x [B][2H][2W][C]
__global__ void __launch_bounds__(256) up2_kernel(const float4* __restrict__ x, int H, int W, int C4, long long n4, int adjoint, float4* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int px = (int)(r % W);
        r /= W;
        const int py = (int)(r % H);
        const long long b = r / H;
        const long long big = ((b * 2 * H + 2 * py) * 2 * W + 2 * px) * C4 + c;
        if (!adjoint) {
            const float4 v = x[i];
            y[big] = v;
            y[big + C4] = v;
            y[big + 2LL * W * C4] = v;
            y[big + 2LL * W * C4 + C4] = v;
        } else {
            const float4 a = x[big], b2 = x[big + C4], c2 = x[big + 2LL * W * C4], d = x[big + 2LL * W * C4 + C4];
            y[i] = make_float4((a.x + b2.x) + (c2.x + d.x), (a.y + b2.y) + (c2.y + d.y), (a.z + b2.z) + (c2.z + d.z), (a.w + b2.w) + (c2.w + d.w));
        }
    }
}

inline unsigned ew_grid(long long n) {
    long long g = (n + 255) / 256;
    return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

bool make_window(int ws, const float* g, Win* w) {
    if (ws < 1 || ws > MAXW || g == nullptr) return false;
    for (int i = 0; i < MAXW; ++i) w->g[i] = i < ws ? g[i] : 0.f;
    return true;
}

}  // namespace

GLARE_API long long glare_ssim_partials(int planes, int H, int W, int ws) {
    const int Ho = H - ws + 1, Wo = W - ws + 1;
    if (planes <= 0 || Ho <= 0 || Wo <= 0) return 0;
    return (long long)planes * ((Ho + T - 1) / T) * ((Wo + T - 1) / T);
}

GLARE_API int glare_ssim_fwd_f32(const float* x, const float* y, int planes, int H, int W, int ws, const float* window_host, float C1, float C2,
                                 float* part, cudaStream_t stream) {
    Win w;
    if (!x || !y || !part || planes <= 0 || planes > 65535 || H < ws || W < ws || !make_window(ws, window_host, &w)) return GLARE_ERR_BAD_ARG;
    const int Ho = H - ws + 1, Wo = W - ws + 1;
    dim3 grid((Wo + T - 1) / T, (Ho + T - 1) / T, planes);
    LS_LAUNCH(ssim_fwd_kernel, grid, T * T, stream, x, y, H, W, ws, w, C1, C2, part);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

GLARE_API int glare_ssim_bwd_f32(const float* x, const float* y, int planes, int H, int W, int ws, const float* window_host, float C1, float C2,
                                 const float* coef, float* g_mu, float* g_e11, float* g_e12, const float* coarse, float* dx,
                                 cudaStream_t stream) {
    Win w;
    if (!x || !y || !coef || !g_mu || !g_e11 || !g_e12 || !dx || planes <= 0 || planes > 65535 || H < ws || W < ws ||
        !make_window(ws, window_host, &w))
        return GLARE_ERR_BAD_ARG;
    const int Ho = H - ws + 1, Wo = W - ws + 1;
    LS_LAUNCH(ssim_bwd_maps_kernel, dim3((Wo + T - 1) / T, (Ho + T - 1) / T, planes), T * T, stream, x, y, H, W, ws, w, C1, C2, coef, g_mu, g_e11, g_e12);
    GLARE_CHECK_LAUNCH();
    LS_LAUNCH(ssim_bwd_apply_kernel, dim3((W + T - 1) / T, (H + T - 1) / T, planes), T * T, stream, x, y, H, W, ws, w, g_mu, g_e11, g_e12, coarse, dx);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// ---- elementwise / pooling entry points (section 8b of include/glare_b200.h) ---------------------------------------------------------------
GLARE_API int glare_relu_f32(const float* x, const float* gy, long long n, float* out, cudaStream_t stream) {
    if (n < 0 || (n & 3)) return GLARE_ERR_BAD_ARG;
    if (n == 0) return GLARE_OK;
    if (!x || !out) return GLARE_ERR_BAD_ARG;
    if (gy == nullptr)
        LS_LAUNCH(relu_fwd_kernel, dim3(ew_grid(n / 4)), 256, stream, reinterpret_cast<const float4*>(x), n / 4, reinterpret_cast<float4*>(out));
    else
        LS_LAUNCH(relu_bwd_kernel, dim3(ew_grid(n / 4)), 256, stream, reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(gy), n / 4,
                  reinterpret_cast<float4*>(out));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

GLARE_API int glare_maxpool2_nhwc_f32(const float* x, int B, int H, int W, int C, float* y, void* idx, cudaStream_t stream) {
    if (B < 0 || H < 2 || W < 2 || C <= 0 || (C & 3)) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !y || !idx) return GLARE_ERR_BAD_ARG;
    const long long n = (long long)B * (H / 2) * (W / 2) * (C / 4);
    LS_LAUNCH(maxpool2_fwd_kernel, dim3(ew_grid(n)), 256, stream, reinterpret_cast<const float4*>(x), H, W, C / 4, n, reinterpret_cast<float4*>(y),
              reinterpret_cast<unsigned*>(idx));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

GLARE_API int glare_maxpool2_nhwc_bwd_f32(const float* gy, const void* idx, int B, int H, int W, int C, float* gx, cudaStream_t stream) {
    if (B < 0 || H < 2 || W < 2 || C <= 0 || (C & 3)) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!gy || !gx || !idx) return GLARE_ERR_BAD_ARG;
    if ((H | W) & 1) GLARE_CUDA(cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)B * H * W * C, stream));
    const long long n = (long long)B * (H / 2) * (W / 2) * (C / 4);
    LS_LAUNCH(maxpool2_bwd_kernel, dim3(ew_grid(n)), 256, stream, reinterpret_cast<const float4*>(gy), reinterpret_cast<const unsigned*>(idx), H, W, C / 4, n,
              reinterpret_cast<float4*>(gx));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

GLARE_API int glare_avgpool2_f32(const float* x, long long planes, int H, int W, float* y, cudaStream_t stream) {
    if (planes < 0 || H < 2 || W < 2) return GLARE_ERR_BAD_ARG;
    if (planes == 0) return GLARE_OK;
    if (!x || !y) return GLARE_ERR_BAD_ARG;
    const long long n = planes * (H / 2) * (W / 2);
    LS_LAUNCH(avgpool2_kernel, dim3(ew_grid(n)), 256, stream, x, H, W, n, y);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// x NHWC; H, W: the SMALL tensor's size.  adjoint = 0: x [B][H][W][C] -> y [B][2H][2W][C]; adjoint = 1: x [B][2H][2W][C] -> y [B][H][W][C]
GLARE_API int glare_up2_nhwc_f32(const float* x, int B, int H, int W, int C, int adjoint, float* y, cudaStream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !y) return GLARE_ERR_BAD_ARG;
    const long long n = (long long)B * H * W * (C / 4);
    LS_LAUNCH(up2_kernel, dim3(ew_grid(n)), 256, stream, reinterpret_cast<const float4*>(x), H, W, C / 4, n, adjoint, reinterpret_cast<float4*>(y));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
