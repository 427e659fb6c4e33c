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
