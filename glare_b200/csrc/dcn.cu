// Modulated deformable convolution (DCNv2) forward, fp32, without the im2col `columns` buffer.
//
// Replaces reference code/models/modules/ops/dcn/src/deform_conv_cuda.cpp:490-569
// (modulated_deform_conv_cuda_forward: per-sample im2col launch + addmm_ + bias) and
// src/deform_conv_cuda_kernel.cu:467-497 (dmcn_im2col_bilinear), :571-633 (modulated_deformable_im2col).
//
//   y[b,co,p] = bias[co] + sum_{c,i,j} W[co,c,i,j] * mask[b,g(c),ij,p] * bilinear(x[b,c], p*stride - pad + (i,j)*dil + off)
//
// Layouts are the reference op's own (deform_conv.py:124-153): x [B,C,H,W], offset [B, dg*2*kh*kw, Ho, Wo]
// (channel g*2*kk + 2*ij + {0: dh, 1: dw}), mask [B, dg*kk, Ho, Wo], y [B,Cout,Ho,Wo]; weights are repacked
// once to [ij][c][co] (glare_dcn_pack_weight_f32) so the B operand loads are coalesced.
//
// Mapping: implicit GEMM, M = 128 consecutive output pixels of one sample, N = 128 output channels,
// K = (group, tap, channel).  The sampling geometry -- four corner offsets and bilinear weights x mask --
// is computed ONCE per (pixel, group, tap) and reused by every channel of the group (the reference recomputes
// it per channel); sampled K-slabs of 8 channels are staged in shared memory, each thread owns an 8x8
// register tile.  fp32 FMA path (parity reference for the tensor-core variant in conv_tc.cu).
#include "common.cuh"

namespace glare {

constexpr int DCN_BM = 128, DCN_BN = 128, DCN_BK = 8, DCN_THREADS = 256;

__global__ void dcn_pack_weight_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int C, int kk) {
    // w [Cout][C][kk] -> wt [kk][C][Cout]
    const long long n = (long long)Cout * C * kk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const long long r = i / Cout;
        const int c = (int)(r % C), t = (int)(r / C);
        wt[i] = w[((long long)co * C + c) * kk + t];
    }
}

struct DcnArgs {
    const float *x, *offset, *mask, *wt, *bias;
    float* y;
    int B, C, H, W, Cout, Ho, Wo, kh, kw, stride, pad, dil, dg;
};

__global__ void __launch_bounds__(DCN_THREADS, 2) dcn_fwd_kernel(const DcnArgs a) {
    __shared__ __align__(16) float s_a[DCN_BK][DCN_BM];      // sampled values   [k][pixel]
    __shared__ __align__(16) float s_b[DCN_BK][DCN_BN];      // weights          [k][co]
    __shared__ int s_off[4][DCN_BM];                         // corner element offsets (-1 = outside)
    __shared__ float s_cw[4][DCN_BM];                        // corner weight * mask

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const long long HoWo = (long long)a.Ho * a.Wo, HW = (long long)a.H * a.W;
    const long long p0 = (long long)blockIdx.x * DCN_BM;
    const int n0 = blockIdx.y * DCN_BN;
    const int kk = a.kh * a.kw, cpg = a.C / a.dg;
    const int tm = tid & 15, tn = tid >> 4;                  // 16 x 16 threads, 8 pixels x 8 channels each

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const float* xb = a.x + (long long)b * a.C * HW;
    for (int g = 0; g < a.dg; ++g) {
        for (int t = 0; t < kk; ++t) {
            __syncthreads();   // previous slab fully consumed before the geometry is overwritten
            if (tid < DCN_BM) {
                const long long p = p0 + tid;
                int o0 = -1, o1 = -1, o2 = -1, o3 = -1;
                float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
                if (p < HoWo) {
                    const int ho = (int)(p / a.Wo), wo = (int)(p - (long long)ho * a.Wo);
                    const int i = t / a.kw, j = t - i * a.kw;
                    const float* ob = a.offset + ((long long)b * a.dg + g) * 2 * kk * HoWo + p;
                    const float oh = ob[(long long)(2 * t) * HoWo], ow = ob[(long long)(2 * t + 1) * HoWo];
                    const float m = a.mask[(((long long)b * a.dg + g) * kk + t) * HoWo + p];
                    // .cu:607-618: h_im = h_in + i*dil + offset_h, sample only inside the open range (-1,H)x(-1,W)
                    const float h_im = (float)(ho * a.stride - a.pad + i * a.dil) + oh;
                    const float w_im = (float)(wo * a.stride - a.pad + j * a.dil) + ow;
                    if (h_im > -1.f && w_im > -1.f && h_im < (float)a.H && w_im < (float)a.W) {
                        const int hl = (int)floorf(h_im), wl = (int)floorf(w_im);
                        const int hh = hl + 1, wh = wl + 1;
                        const float lh = h_im - hl, lw = w_im - wl;
                        const float uh = 1.f - lh, uw = 1.f - lw;
                        if (hl >= 0 && wl >= 0) { o0 = hl * a.W + wl; w0 = uh * uw * m; }
                        if (hl >= 0 && wh <= a.W - 1) { o1 = hl * a.W + wh; w1 = uh * lw * m; }
                        if (hh <= a.H - 1 && wl >= 0) { o2 = hh * a.W + wl; w2 = lh * uw * m; }
                        if (hh <= a.H - 1 && wh <= a.W - 1) { o3 = hh * a.W + wh; w3 = lh * lw * m; }
                    }
                }
                s_off[0][tid] = o0; s_off[1][tid] = o1; s_off[2][tid] = o2; s_off[3][tid] = o3;
                s_cw[0][tid] = w0; s_cw[1][tid] = w1; s_cw[2][tid] = w2; s_cw[3][tid] = w3;
            }
            __syncthreads();
            const int pm = tid & (DCN_BM - 1);
            const int o0 = s_off[0][pm], o1 = s_off[1][pm], o2 = s_off[2][pm], o3 = s_off[3][pm];
            const float w0 = s_cw[0][pm], w1 = s_cw[1][pm], w2 = s_cw[2][pm], w3 = s_cw[3][pm];
            for (int cc = 0; cc < cpg; cc += DCN_BK) {
                // A slab: 8 channels x 128 pixels, 4 samples per thread
#pragma unroll
                for (int q = 0; q < DCN_BK / 2; ++q) {
                    const int kq = (tid >> 7) + 2 * q;
                    const int ch = g * cpg + cc + kq;
                    float v = 0.f;
                    if (cc + kq < cpg) {
                        const float* xc = xb + (long long)ch * HW;
                        if (o0 >= 0) v = fmaf(w0, __ldg(xc + o0), v);
                        if (o1 >= 0) v = fmaf(w1, __ldg(xc + o1), v);
                        if (o2 >= 0) v = fmaf(w2, __ldg(xc + o2), v);
                        if (o3 >= 0) v = fmaf(w3, __ldg(xc + o3), v);
                    }
                    s_a[kq][pm] = v;
                }
                // B slab: wt[t][c][co], 8 channels x 128 co, 4 values per thread
#pragma unroll
                for (int q = 0; q < DCN_BK / 2; ++q) {
                    const int kq = (tid >> 7) + 2 * q;
                    const int ch = g * cpg + cc + kq, co = n0 + pm;
                    float v = 0.f;
                    if (cc + kq < cpg && co < a.Cout) v = __ldg(a.wt + ((long long)t * a.C + ch) * a.Cout + co);
                    s_b[kq][pm] = v;
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < DCN_BK; ++k) {
                    const float4 a0 = *reinterpret_cast<const float4*>(&s_a[k][tm * 8]);
                    const float4 a1 = *reinterpret_cast<const float4*>(&s_a[k][tm * 8 + 4]);
                    const float4 b0 = *reinterpret_cast<const float4*>(&s_b[k][tn * 8]);
                    const float4 b1 = *reinterpret_cast<const float4*>(&s_b[k][tn * 8 + 4]);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                }
                __syncthreads();
            }
        }
    }
    // epilogue: + bias, NCHW store (8 consecutive pixels per thread per channel)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int co = n0 + tn * 8 + j;
        if (co >= a.Cout) continue;
        const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
        float* yb = a.y + ((long long)b * a.Cout + co) * HoWo;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long p = p0 + tm * 8 + i;
            if (p < HoWo) yb[p] = acc[i][j] + bv;
        }
    }
}

}  // namespace glare

using namespace glare;

GLARE_API int glare_dcn_pack_weight_f32(const float* weight, int Cout, int C, int kh, int kw, float* packed_out,
                                         cudaStream_t stream) {
    if (!weight || !packed_out || Cout <= 0 || C <= 0 || kh <= 0 || kw <= 0) return GLARE_ERR_BAD_ARG;
    const long long n = (long long)Cout * C * kh * kw;
    const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    dcn_pack_weight_kernel<<<grid, 256, 0, stream>>>(weight, packed_out, Cout, C, kh * kw);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

GLARE_API int glare_dcnv2_fwd_f32(const float* x, const float* offset, const float* mask, const float* packed_weight,
                                   const float* bias_or_null, int B, int C, int H, int W, int Cout, int kh, int kw,
                                   int stride, int pad, int dil, int deformable_groups, float* y, cudaStream_t stream) {
    if (B == 0) return GLARE_OK;
    if (!x || !offset || !mask || !packed_weight || !y) return GLARE_ERR_BAD_ARG;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || Cout <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0 || dil <= 0 ||
        deformable_groups <= 0 || C % deformable_groups != 0)
        return GLARE_ERR_BAD_ARG;
    DcnArgs a{};
    a.x = x; a.offset = offset; a.mask = mask; a.wt = packed_weight; a.bias = bias_or_null; a.y = y;
    a.B = B; a.C = C; a.H = H; a.W = W; a.Cout = Cout; a.kh = kh; a.kw = kw;
    a.stride = stride; a.pad = pad; a.dil = dil; a.dg = deformable_groups;
    a.Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    a.Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    if (a.Ho <= 0 || a.Wo <= 0) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (B > 65535) return GLARE_ERR_BAD_ARG;
    const long long HoWo = (long long)a.Ho * a.Wo;
    dim3 grid((unsigned)((HoWo + DCN_BM - 1) / DCN_BM), (unsigned)((Cout + DCN_BN - 1) / DCN_BN), (unsigned)B);
    dcn_fwd_kernel<<<grid, DCN_THREADS, 0, stream>>>(a);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
