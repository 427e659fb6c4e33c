// Modulated deformable convolution (DCNv2) backward, 3x3 / stride 1 / pad 1 / dilation 1 / groups 1 (GLARE's configuration).
//
// Replaces reference ops/dcn/src/deform_conv_cuda.cpp:571-685 (modulated_deform_conv_cuda_backward: per-sample GEMM into a
// `columns` buffer, then three kernels) and deform_conv_cuda_kernel.cu:499-567 (gradient / coordinate weights), :636-693
// (modulated_deformable_col2im: grad input by atomicAdd), :696-767 (col2im_coord: grad offset and grad mask).
//
// Dataflow here (glare_b200/dcn_backward.py):
//   dcol[p,(t,c)] = sum_co gout[p,co] W[co,c,t]        1x1 conv on the tensor-core path (conv_tc.cu) with W^T as the filter
//   dcn_bwd_data_kernel: one thread per (sample, pixel, deformable group, tap) walks the group's channels four at a time:
//        val      = bilinear(x[c], pos)                                      (dmcn_im2col_bilinear, .cu:467-497)
//        grad_mask   += dcol * val                                            (.cu:744-746)
//        grad_offset += dcol * mask * d val / d{h,w}                          (dmcn_get_coordinate_weight, .cu:527-567)
//        grad_x[corner] += dcol * mask * corner weight   (128-bit vector atomics)    (dmcn_get_gradient_weight, .cu:499-525)
//        col[p,(t,c)] = mask * val      (optional: the im2col operand of the weight gradient, written in the same pass)
//   dcn_bwd_weight_kernel: grad_W[(t,c),co] += sum_p col[p,(t,c)] gout[p,co] -- split-K SGEMM over the pixels of the whole batch
//        (both operands are pixel-major, i.e. K-major rows), fp32 FMA, partial tiles combined with atomics.
// Layouts: x, grad_x, dcol, col, gout NHWC; offset / mask / grad_offset / grad_mask NCHW exactly as the reference op
// (offset channel g*18 + 2t + {0: dh, 1: dw}, mask channel g*9 + t).
#include "common.cuh"

namespace glare {

struct DcnBwdArgs {
    const float *x, *offset, *mask, *dcol;
    float *grad_x, *grad_offset, *grad_mask, *col;
    int B, C, H, W, dg, cpg;
};

__global__ void __launch_bounds__(128) dcn_bwd_data_kernel(const DcnBwdArgs a) {
    const long long HW = (long long)a.H * a.W;
    const long long total = (long long)a.B * a.dg * 9 * HW;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx % HW;
        long long r = idx / HW;
        const int t = (int)(r % 9);
        r /= 9;
        const int g = (int)(r % a.dg), b = (int)(r / a.dg);
        const int ho = (int)(p / a.W), wo = (int)(p - (long long)ho * a.W);
        const long long obase = ((long long)b * a.dg + g) * 18 * HW + p;
        const float oh = __ldg(a.offset + obase + (long long)(2 * t) * HW), ow = __ldg(a.offset + obase + (long long)(2 * t + 1) * HW);
        const long long mpos = (((long long)b * a.dg + g) * 9 + t) * HW + p;
        const float mk = __ldg(a.mask + mpos);
        const float h_im = (float)(ho - 1 + t / 3) + oh, w_im = (float)(wo - 1 + t % 3) + ow;
        const bool inside = h_im > -1.f && w_im > -1.f && h_im < (float)a.H && w_im < (float)a.W;

        const int hl = (int)floorf(h_im), wl = (int)floorf(w_im);
        const int hh = hl + 1, wh = wl + 1;
        const float lh = h_im - hl, lw = w_im - wl, uh = 1.f - lh, uw = 1.f - lw;
        const bool ok0 = inside && hl >= 0 && wl >= 0, ok1 = inside && hl >= 0 && wh <= a.W - 1;
        const bool ok2 = inside && hh <= a.H - 1 && wl >= 0, ok3 = inside && hh <= a.H - 1 && wh <= a.W - 1;
        const long long pix_b = (long long)b * HW;
        const long long c0 = (pix_b + (ok0 ? (long long)hl * a.W + wl : 0)) * a.C, c1 = (pix_b + (ok1 ? (long long)hl * a.W + wh : 0)) * a.C;
        const long long c2 = (pix_b + (ok2 ? (long long)hh * a.W + wl : 0)) * a.C, c3 = (pix_b + (ok3 ? (long long)hh * a.W + wh : 0)) * a.C;
        const float w0 = uh * uw, w1 = uh * lw, w2 = lh * uw, w3 = lh * lw;

        const long long dbase = (pix_b + p) * 9 * a.C + (long long)t * a.C + (long long)g * a.cpg;
        float gm = 0.f, goh = 0.f, gow = 0.f;
        for (int c = 0; c < a.cpg; c += 4) {
            const int ch = g * a.cpg + c;
            const float4 d = __ldg(reinterpret_cast<const float4*>(a.dcol + dbase + c));
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v0 = ok0 ? __ldg(reinterpret_cast<const float4*>(a.x + c0 + ch)) : z;
            const float4 v1 = ok1 ? __ldg(reinterpret_cast<const float4*>(a.x + c1 + ch)) : z;
            const float4 v2 = ok2 ? __ldg(reinterpret_cast<const float4*>(a.x + c2 + ch)) : z;
            const float4 v3 = ok3 ? __ldg(reinterpret_cast<const float4*>(a.x + c3 + ch)) : z;
            const float dd[4] = {d.x, d.y, d.z, d.w};
            const float a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w};
            const float a2[4] = {v2.x, v2.y, v2.z, v2.w}, a3[4] = {v3.x, v3.y, v3.z, v3.w};
            float cv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float val = w0 * a0[k] + w1 * a1[k] + w2 * a2[k] + w3 * a3[k];
                cv[k] = val * mk;
                gm = fmaf(dd[k], val, gm);
                // d val / dh and d val / dw (dmcn_get_coordinate_weight, bp_dir 0 / 1)
                const float dvh = -uw * a0[k] - lw * a1[k] + uw * a2[k] + lw * a3[k];
                const float dvw = -uh * a0[k] + uh * a1[k] - lh * a2[k] + lh * a3[k];
                goh = fmaf(dd[k] * mk, dvh, goh);
                gow = fmaf(dd[k] * mk, dvw, gow);
            }
            if (a.col) *reinterpret_cast<float4*>(a.col + dbase + c) = make_float4(cv[0], cv[1], cv[2], cv[3]);
            if (a.grad_x) {
                const float4 gd = make_float4(d.x * mk, d.y * mk, d.z * mk, d.w * mk);
                if (ok0) atomicAdd(reinterpret_cast<float4*>(a.grad_x + c0 + ch), make_float4(gd.x * w0, gd.y * w0, gd.z * w0, gd.w * w0));
                if (ok1) atomicAdd(reinterpret_cast<float4*>(a.grad_x + c1 + ch), make_float4(gd.x * w1, gd.y * w1, gd.z * w1, gd.w * w1));
                if (ok2) atomicAdd(reinterpret_cast<float4*>(a.grad_x + c2 + ch), make_float4(gd.x * w2, gd.y * w2, gd.z * w2, gd.w * w2));
                if (ok3) atomicAdd(reinterpret_cast<float4*>(a.grad_x + c3 + ch), make_float4(gd.x * w3, gd.y * w3, gd.z * w3, gd.w * w3));
            }
        }
        if (a.grad_mask) a.grad_mask[mpos] = inside ? gm : 0.f;
        if (a.grad_offset) {
            a.grad_offset[obase + (long long)(2 * t) * HW] = inside ? goh : 0.f;
            a.grad_offset[obase + (long long)(2 * t + 1) * HW] = inside ? gow : 0.f;
        }
    }
}

// grad_wp[m][n] += sum_{p in this CTA's pixel range} col[p][m] * gout[p][n];  128 x 128 tile, 8 x 8 per thread, K chunks of 8 pixels
constexpr int WG_BM = 128, WG_BN = 128, WG_BK = 8, WG_THREADS = 256;

__global__ void __launch_bounds__(WG_THREADS, 2) dcn_bwd_weight_kernel(const float* __restrict__ col, const float* __restrict__ gout,
                                                                       long long P, int M, int N, long long p_per_cta,
                                                                       float* __restrict__ grad_wp) {
    __shared__ __align__(16) float s_a[WG_BK][WG_BM];
    __shared__ __align__(16) float s_b[WG_BK][WG_BN];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * WG_BM, n0 = blockIdx.y * WG_BN;
    const long long k0 = (long long)blockIdx.z * p_per_cta, k1 = (k0 + p_per_cta < P) ? k0 + p_per_cta : P;
    const int tm = tid & 15, tn = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int lm = tid & 127, lk = tid >> 7;
    for (long long k = k0; k < k1; k += WG_BK) {
#pragma unroll
        for (int q = 0; q < WG_BK / 2; ++q) {
            const long long p = k + lk + 2 * q;
            s_a[lk + 2 * q][lm] = (p < k1 && m0 + lm < M) ? __ldg(col + p * M + m0 + lm) : 0.f;
            s_b[lk + 2 * q][lm] = (p < k1 && n0 + lm < N) ? __ldg(gout + p * N + n0 + lm) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < WG_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&s_a[kk][tm * 8]), a1 = *reinterpret_cast<const float4*>(&s_a[kk][tm * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&s_b[kk][tn * 8]), b1 = *reinterpret_cast<const float4*>(&s_b[kk][tn * 8 + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + tm * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tn * 8 + j;
            if (n < N) atomicAdd(grad_wp + (long long)m * N + n, acc[i][j]);
        }
    }
}

}  // namespace glare

using namespace glare;

// x NHWC [B,H,W,C]; offset NCHW [B,18dg,H,W]; mask NCHW [B,9dg,H,W]; dcol NHWC [B,H,W,9C] (channel t*C + c) = W^T applied to grad_output.
// Outputs (any may be NULL): grad_x NHWC [B,H,W,C] ACCUMULATED into (caller zero-fills), grad_offset / grad_mask NCHW (overwritten),
// col NHWC [B,H,W,9C] = the im2col operand mask*bilinear(x) for glare_dcnv2_bwd_weight_f32.
GLARE_API int glare_dcnv2_bwd_data_f32(const float* x, const float* offset, const float* mask, const float* dcol, int B, int C, int H, int W,
                                       int deformable_groups, float* grad_x, float* grad_offset, float* grad_mask, float* col,
                                       cudaStream_t stream) {
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || deformable_groups <= 0 || C % deformable_groups != 0 || (C / deformable_groups) % 4 != 0)
        return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !offset || !mask || !dcol) return GLARE_ERR_BAD_ARG;
    DcnBwdArgs a{};
    a.x = x; a.offset = offset; a.mask = mask; a.dcol = dcol;
    a.grad_x = grad_x; a.grad_offset = grad_offset; a.grad_mask = grad_mask; a.col = col;
    a.B = B; a.C = C; a.H = H; a.W = W; a.dg = deformable_groups; a.cpg = C / deformable_groups;
    const long long total = (long long)B * deformable_groups * 9 * H * W;
    const long long blocks = (total + 127) / 128;
    const int grid = (int)(blocks < 148LL * 64 ? blocks : 148LL * 64);
    dcn_bwd_data_kernel<<<grid, 128, 0, stream>>>(a);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// grad_w_packed [9C][Cout] += col^T gout over P = B*H*W pixels (col [P][9C], gout [P][Cout], both NHWC-flattened).
// grad_weight[co,c,i,j] = grad_w_packed[(3i+j)*C + c][co].
GLARE_API int glare_dcnv2_bwd_weight_f32(const float* col, const float* gout, long long P, int KC, int Cout, float* grad_w_packed,
                                         cudaStream_t stream) {
    if (P < 0 || KC <= 0 || Cout <= 0) return GLARE_ERR_BAD_ARG;
    if (P == 0) return GLARE_OK;
    if (!col || !gout || !grad_w_packed) return GLARE_ERR_BAD_ARG;
    const int gm = (KC + WG_BM - 1) / WG_BM, gn = (Cout + WG_BN - 1) / WG_BN;
    long long split = (148LL * 4 + gm * gn - 1) / (gm * gn);
    const long long max_split = (P + 1023) / 1024;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    if (split > 65535) split = 65535;
    long long per = (P + split - 1) / split;
    per = (per + WG_BK - 1) / WG_BK * WG_BK;
    split = (P + per - 1) / per;
    dcn_bwd_weight_kernel<<<dim3(gm, gn, (unsigned)split), WG_THREADS, 0, stream>>>(col, gout, P, KC, Cout, per, grad_w_packed);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
