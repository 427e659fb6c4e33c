/*
 * oracle_kernels.c -- CPU restatement (plain C, scalar) of the two GLARE hot-path operators whose
 * results must be reproduced index-/sample-exactly.
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg as the checker.  Nothing in glare_b200/ (the product) may call into this file.
 *
 * Build:  make -C oracle        (gcc -O2 -ffp-contract=off: every multiply/add below rounds exactly
 *                                where it is written; the only fused operations are explicit fmaf)
 *
 * Parity pin: oracle/gen_golden.py checks both functions against the reference's own Python modules
 * (imported from /root/reference) and commits the vectors under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* ---------------------------------------------------------------------------------------------
 * VectorQuantizer2.forward  (reference: code/models/modules/quantize.py:271-312)
 *   d[t,k] = sum(z_t^2) + sum(e_k^2) - 2 * einsum('bd,dn->bn')        quantize.py:280-282
 *   idx[t] = argmin_k d[t,k]  (lowest index among equal minima)         quantize.py:284
 *   z_q[t] = E[idx[t]]                                                  quantize.py:285
 * fp32 recipe (bit-compared against torch CPU / MKL sgemm with K=3):
 *   tn  = fl(fl(fl(t0*t0)+fl(t1*t1))+fl(t2*t2))            each square rounded, no FMA
 *   cn  = same for the code vector
 *   dot = fma(t2,c2, fma(t1,c1, fl(t0*c0)))                one multiply then an FMA chain in k order
 *   d   = fl(fl(tn+cn) - fl(2*dot))
 * z is NCHW [B,3,h,w] (the layout the reference module receives, quantize.py:276 permutes it);
 * idx is written in (b, y, x) token order, z_q back in NCHW (quantize.py:301).
 * ------------------------------------------------------------------------------------------- */
void glare_oracle_vq_f32(const float *z, const float *codebook, int B, int hw, int K,
                         int64_t *idx_out, float *zq_out, float *dmin_out /* may be NULL */)
{
    for (int b = 0; b < B; ++b) {
        const float *z0 = z + (size_t)b * 3 * hw, *z1 = z0 + hw, *z2 = z1 + hw;
        for (int p = 0; p < hw; ++p) {
            const float t0 = z0[p], t1 = z1[p], t2 = z2[p];
            const float tn = ((t0 * t0) + (t1 * t1)) + (t2 * t2);
            float best = INFINITY;
            int besti = 0;
            for (int k = 0; k < K; ++k) {
                const float c0 = codebook[3 * k], c1 = codebook[3 * k + 1], c2 = codebook[3 * k + 2];
                const float cn = ((c0 * c0) + (c1 * c1)) + (c2 * c2);
                const float dot = fmaf(t2, c2, fmaf(t1, c1, t0 * c0));
                const float d = (tn + cn) - (2.0f * dot);
                /* torch.argmin: first minimum wins; a NaN distance is propagated as the minimum */
                if (d < best || (d != d && best == best)) { best = d; besti = k; }
            }
            idx_out[(size_t)b * hw + p] = besti;
            float *q = zq_out + (size_t)b * 3 * hw;
            /* quantize.py:298 straight-through: z_q = z + (z_q - z).detach(), evaluated in fp32 */
            q[p] = t0 + (codebook[3 * besti] - t0);
            q[hw + p] = t1 + (codebook[3 * besti + 1] - t1);
            q[2 * hw + p] = t2 + (codebook[3 * besti + 2] - t2);
            if (dmin_out) dmin_out[(size_t)b * hw + p] = best;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Modulated deformable im2col  (reference: ops/dcn/src/deform_conv_cuda_kernel.cu:467-497 bilinear,
 * :571-633 kernel; host loop deform_conv_cuda.cpp:539-560).  3x3-general: kernel kh x kw, stride,
 * pad, dilation, `dg` deformable groups.  Layouts exactly as the reference op receives them:
 *   x      [B, C, H, W]
 *   offset [B, dg*2*kh*kw, Ho, Wo]   channel = g*2*kh*kw + 2*(i*kw+j) + {0: dh, 1: dw}
 *   mask   [B, dg*kh*kw,   Ho, Wo]   channel = g*kh*kw + (i*kw+j)
 *   cols   [B, C*kh*kw, Ho*Wo]       row = c*kh*kw + i*kw + j   (per-sample matrix the GEMM consumes)
 * A sample is taken only if  h_im > -1 && w_im > -1 && h_im < H && w_im < W  (.cu:618); each of
 * the four corners contributes only when inside the image (.cu:481-492).
 * ------------------------------------------------------------------------------------------- */
static float bilinear(const float *im, int W, int H, float h, float w)
{
    int h_low = (int)floorf(h), w_low = (int)floorf(w);
    int h_high = h_low + 1, w_high = w_low + 1;
    float lh = h - h_low, lw = w - w_low;
    float hh = 1 - lh, hw = 1 - lw;
    float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (h_low >= 0 && w_low >= 0) v1 = im[h_low * W + w_low];
    if (h_low >= 0 && w_high <= W - 1) v2 = im[h_low * W + w_high];
    if (h_high <= H - 1 && w_low >= 0) v3 = im[h_high * W + w_low];
    if (h_high <= H - 1 && w_high <= W - 1) v4 = im[h_high * W + w_high];
    float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

void glare_oracle_dcn_im2col_f32(const float *x, const float *offset, const float *mask,
                                 int B, int C, int H, int W, int kh, int kw, int stride, int pad,
                                 int dil, int dg, float *cols)
{
    const int Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    const int Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    const int cpg = C / dg, kk = kh * kw;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const int g = c / cpg;
            const float *im = x + ((size_t)b * C + c) * H * W;
            const float *off = offset + ((size_t)b * dg + g) * 2 * kk * Ho * Wo;
            const float *msk = mask + ((size_t)b * dg + g) * kk * Ho * Wo;
            for (int i = 0; i < kh; ++i)
                for (int j = 0; j < kw; ++j) {
                    float *col = cols + (((size_t)b * C + c) * kk + i * kw + j) * Ho * Wo;
                    const float *oh = off + (size_t)(2 * (i * kw + j)) * Ho * Wo;
                    const float *ow = oh + (size_t)Ho * Wo;
                    const float *m = msk + (size_t)(i * kw + j) * Ho * Wo;
                    for (int y = 0; y < Ho; ++y)
                        for (int xx = 0; xx < Wo; ++xx) {
                            const int p = y * Wo + xx;
                            const float h_im = (y * stride - pad) + i * dil + oh[p];
                            const float w_im = (xx * stride - pad) + j * dil + ow[p];
                            float val = 0.f;
                            if (h_im > -1 && w_im > -1 && h_im < H && w_im < W)
                                val = bilinear(im, W, H, h_im, w_im);
                            col[p] = val * m[p];
                        }
                }
        }
}
