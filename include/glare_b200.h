/*
 * glare_b200.h -- C ABI of libglare_b200.so: the B200 (sm_100a) kernels behind GLARE's hot path.
 *
 * This is the drop-in boundary.  The reference's only native plug-in is the pybind11 torch extension
 * `deform_conv_ext` (code/models/modules/ops/dcn/src/deform_conv_ext.cpp:150-164); every other operator on
 * the path is a Python nn.Module calling ATen.  Each entry point below names the reference interface it
 * replaces (paths relative to /root/reference/code/models/modules unless stated).  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference adds at each of those call sites.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, explicit sizes/strides (in elements), the CUDA stream to launch on.
 *     No torch / ATen types, no global mutable state, re-entrant, asynchronous w.r.t. the host
 *     (deform_conv_cuda.cpp launches on at::cuda::getCurrentCUDAStream(); the caller passes that stream).
 *   - the caller allocates every output (deform_conv.py:147 `input.new_empty(shape)` convention).
 *   - return value: 0 = launched; < 0 = GLARE_ERR_* argument error (nothing launched); > 0 = cudaError_t.
 *     (the reference only printf()s launch errors, deform_conv_cuda_kernel.cu:794-798; here the Python shim
 *     raises RuntimeError on any non-zero code.)
 *   - tensors are dense and contiguous in the layout stated per function (the reference TORCH_CHECKs
 *     contiguity, deform_conv_cuda.cpp:497-498).
 */
#ifndef GLARE_B200_H_
#define GLARE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define GLARE_OK 0
#define GLARE_ERR_BAD_ARG (-1)
#define GLARE_ERR_UNSUPPORTED (-2)

#define GLARE_ABI_VERSION 1
int glare_abi_version(void);
/* static string for a return code of this library (GLARE_ERR_* or cudaError_t) */
const char* glare_error_string(int code);

/* ------------------------------------------------------------------------------------------------------
 * (1) VectorQuantizer2.forward -- quantize.py:271-312
 *     d = |z|^2 + |e|^2 - 2 z.e (:280-282), argmin (:284, lowest index among equal minima), embedding
 *     gather (:285), straight-through z_q = z + (e - z) (:298), NCHW<->NHWC rearranges (:276, :301) fused.
 *     Indices are bit-exact w.r.t. the reference's CPU fp32 evaluation order.
 * ---------------------------------------------------------------------------------------------------- */
/* codebook [K,3] fp32 -> packed [K,4] = {e0,e1,e2,|e|^2}; run once per codebook */
int glare_vq_pack_codebook_f32(const float* codebook, int K, float* packed_out, cudaStream_t stream);
/* z [B,3,hw] fp32 (NCHW) -> idx int64 [B*hw] in (b,y,x) order, z_q [B,3,hw] fp32 (NCHW) */
int glare_vq_argmin_gather_f32(const float* z_nchw, const float* packed_codebook, int B, int hw, int K,
                               long long* idx_out, float* zq_nchw_out, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (2) FlowStep.normal_flow / reverse_flow -- FlowStep.py:75-119, with ActNorm2d (FlowActNorms.py:48-100),
 *     InvertibleConv1x1 (Permutations.py:21-59) and CondAffineSeparatedAndCond
 *     (FlowAffineCouplingsAblation.py:50-151; NN = flow.py:13-70 Conv2d/Conv2dZeros) fused into one launch.
 *
 *     The first 3x3 conv of each coupling net is linear; its conditioning-feature part ("pre-activation
 *     planes": 64 channels per net) is produced by the dense conv path for all steps at once, see DESIGN.md.
 *
 *     Packed net block (GLARE_FLOW_NET_FLOATS floats, built by glare_b200.flow.pack_net):
 *       [0,576)      w1z  [64][9]      first-layer weights of the z1 input channel (zeros for NN_F)
 *       [576,640)    b1   [64]         ActNorm bias          [640,704)   s1 = exp(ActNorm logs)
 *       [704,4800)   w2t  [64 in][64 out]                    1x1 conv, transposed
 *       [4800,4864)  b2                                      [4864,4928) s2
 *       [4928,9536)  w3   [64][9][8]   last 3x3 conv, output channel padded to 8
 *       [9536,9544)  b3   [8]                                [9544,9552) s3 = exp(3 * logs)
 *     Pointwise block (16 floats): M[9] row-major (W in normal_flow, fp64-inverted W^-1 in reverse_flow),
 *       ActNorm bias[3], ActNorm scale[3] (exp(logs) normal / exp(-logs) reverse), pad.
 * ---------------------------------------------------------------------------------------------------- */
#define GLARE_FLOW_NET_FLOATS 9552
#define GLARE_FLOW_PW_FLOATS 16
int glare_flow_net_floats(void);
/* NN tail (ActNorm+ReLU, 1x1+ActNorm+ReLU, 3x3 -> nout in {4,6}) of n_steps nets in one launch; used for
 * NN_F = feature_extract (FlowAffineCouplingsAblation.py:121-128) of every step, which never sees z.
 * p is addressed as p[b*batch_stride + s*step_stride + c*chan_stride + pixel*pix_stride] (NCHW planes: chan_stride = h*w,
 * pix_stride = 1; NHWC: chan_stride = 1, pix_stride = channels); out[b][s][nout][h*w] through its two strides. */
int glare_flow_cond_tail_f32(const float* p, long long p_batch_stride, long long p_step_stride, long long p_chan_stride,
                             long long p_pix_stride, const float* nets, int n_steps, int nout, int B, int h, int w, float* out, long long out_batch_stride,
                             long long out_step_stride, cudaStream_t stream);
/* One FlowStep.  direction 0 = normal_flow, 1 = reverse_flow; coupling 0 = "noCoupling" step.
 * z_in/z_out [B,3,h,w] (must not alias); pA [b][64][h*w] = NN_A first-layer pre-activations from ft;
 * hF [b][6][h*w] = NN_F output; logdet [B] or NULL receives += sum(log scale) (negated in reverse). */
int glare_flow_step_f32(int direction, int coupling, const float* z_in, float* z_out, const float* pA,
                        long long pA_batch_stride, long long pA_chan_stride, long long pA_pix_stride, const float* hF, long long hF_batch_stride, const float* netA,
                        const float* pw, int B, int h, int w, float* logdet, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (3) modulated_deform_conv_forward -- ops/dcn/src/deform_conv_ext.cpp:52-147 (pybind signature),
 *     deform_conv_cuda.cpp:490-569 (host loop), deform_conv_cuda_kernel.cu:571-633 (im2col kernel).
 *     x [B,C,H,W]; offset [B, dg*2*kh*kw, Ho, Wo]; mask [B, dg*kh*kw, Ho, Wo]; y [B,Cout,Ho,Wo]; groups = 1
 *     (the only value GLARE uses, deformableDecoder_arch.py:151-152).  No `columns` / `ones` scratch.
 * ---------------------------------------------------------------------------------------------------- */
/* weight [Cout,C,kh,kw] -> packed [kh*kw][C][Cout]; run once per weight update */
int glare_dcn_pack_weight_f32(const float* weight, int Cout, int C, int kh, int kw, float* packed_out,
                              cudaStream_t stream);
int glare_dcnv2_fwd_f32(const float* x, const float* offset, const float* mask, const float* packed_weight,
                        const float* bias_or_null, int B, int C, int H, int W, int Cout, int kh, int kw, int stride,
                        int pad, int dil, int deformable_groups, float* y, cudaStream_t stream);

/* modulated_deform_conv_backward -- ops/dcn/src/deform_conv_ext.cpp:109-147 (pybind signature), deform_conv_cuda.cpp:571-685,
 * kernels deform_conv_cuda_kernel.cu:499-567, 636-767.  3x3, stride 1, pad 1, dilation 1, groups 1.  x / grad_x / dcol / col NHWC,
 * offset / mask and their gradients NCHW as in the reference op.  dcol [B,H,W,9C] (channel t*C + c) is W^T applied to grad_output
 * (a 1x1 glare_conv2d_nhwc_tc with the transposed filter).  grad_x is accumulated into (caller zero-fills); outputs may be NULL. */
int glare_dcnv2_bwd_data_f32(const float* x, const float* offset, const float* mask, const float* dcol, int B, int C, int H, int W,
                             int deformable_groups, float* grad_x, float* grad_offset, float* grad_mask, float* col,
                             cudaStream_t stream);
/* grad_w_packed [9C][Cout] += col^T gout over P = B*H*W pixels; grad_weight[co,c,i,j] = grad_w_packed[(3i+j)*C + c][co] */
int glare_dcnv2_bwd_weight_f32(const float* col, const float* gout, long long P, int KC, int Cout, float* grad_w_packed,
                               cudaStream_t stream);

/* DCNv2Pack.forward tail on tensor cores (deformableDecoder_arch.py:141-152): consumes the RAW conv_offset output
 * (chunk / cat / sigmoid fused), x and offmask NHWC fp32, weights packed by glare_conv_pack_weight, y NHWC fp32.
 * 3x3, stride 1, pad 1, dilation 1, groups 1; (C / deformable_groups) % 64 == 0 (mode 0) or % 32 == 0 (modes 1, 2). */
int glare_dcnv2_pack_fwd_nhwc_tc(int mode, const float* x, const float* offmask, const void* w, const void* w_lo,
                                 const float* bias_or_null, float* y, int B, int H, int W, int C, int Cout,
                                 int deformable_groups, cudaStream_t stream);
/* The reference OPERATOR's inputs on the same tensor-core kernel (ModulatedDeformConvFunction.forward, ops/dcn/deform_conv.py:124-153 ->
 * deform_conv_ext.modulated_deform_conv_forward, src/deform_conv_ext.cpp:125-134) for 3x3 / stride 1 / pad 1 / dilation 1 / groups 1:
 * offset_mask NHWC [B,H,W,27*dg] holds the op's offset tensor in channels [0,18dg) and its mask tensor (already sigmoid-ed) in
 * [18dg,27dg).  Same shape limits as glare_dcnv2_pack_fwd_nhwc_tc. */
int glare_dcnv2_fwd_nhwc_tc(int mode, const float* x, const float* offset_mask, const void* w, const void* w_lo,
                            const float* bias_or_null, float* y, int B, int H, int W, int C, int Cout, int deformable_groups,
                            cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (4) Dense convolutions on tcgen05 tensor cores -- the cuDNN calls behind nn.Conv2d in ResnetBlock
 *     (encoder_decoder.py:88-115), Upsample.conv (:38-53), AttnBlock q/k/v/proj_out (:146-165), nin_shortcut,
 *     WarpBlock.offset / conv_offset (deformableDecoder_arch.py:282) and the flow nets' first layers (flow.py:13-52).
 *     3x3 (pad 1) or 1x1, stride 1.  Activations NHWC, weights packed [Cout][kh*kw][Cin].
 *     mode 0 = bf16 operands, 1 = tf32, 2 = 3xTF32 (hi/lo split operands, ~fp32 accuracy); fp32 accumulate.
 *     Requires Cin % 64 == 0 (mode 0) / Cin % 32 == 0 (modes 1,2) and Cout % 4 == 0, else GLARE_ERR_UNSUPPORTED.
 * ---------------------------------------------------------------------------------------------------- */
int glare_conv_tc_elem_bytes(int mode);
/* OIHW fp32 -> packed operand(s); out_lo only for mode 2 */
int glare_conv_pack_weight(int mode, const float* w_oihw, int Cout, int Cin, int ksize, void* out_hi, void* out_lo,
                           cudaStream_t stream);
/* fp32 activations (n elements, n % 4 == 0) -> bf16 copy (mode 0) or tf32 hi/lo split (mode 2) */
int glare_conv_prep_act(int mode, const float* x, long long n, void* out_hi, void* out_lo, cudaStream_t stream);
int glare_conv2d_nhwc_tc(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo, const float* bias,
                         const float* residual, float* y, int B, int H, int W, int Cin, int Cout, int ksize,
                         cudaStream_t stream);

/* Downsample.forward (encoder_decoder.py:68-72): zero pad right/bottom by 1 + 3x3 stride-2 conv; y NHWC [B,(Hin-2)/2+1,(Win-2)/2+1,Cout] */
int glare_conv2d_nhwc_tc_down2(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo, const float* bias,
                               float* y, int B, int Hin, int Win, int Cin, int Cout, cudaStream_t stream);
/* one sub-pixel phase (a, b) of Upsample.forward (encoder_decoder.py:49-53, nearest x2 + 3x3 conv): 2x2 filter of pre-summed taps on the
 * low-resolution x NHWC [B,H,W,Cin], writes pixels (2i+a, 2j+b) of y NHWC [B,2H,2W,Cout]; w packed [Cout][4][Cin] for this phase */
int glare_conv2d_nhwc_tc_up2_phase(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo, const float* bias,
                                   float* y, int B, int H, int W, int Cin, int Cout, int a, int b, cudaStream_t stream);
/* general form: kind 0 = stride-1 conv, 1 = Downsample conv, 2 = sub-pixel phase (pa, pb) of Upsample+conv.  gn_stats (optional, kinds 0
 * and 1, Cout in {128, 256, 384, 512}, [B][32][2] fp64) receives the GroupNorm(32) sum / sum of squares of the OUTPUT (the statistics of the
 * Normalize that consumes it, encoder_decoder.py:34-35): the epilogue sums its staged output tiles into per-(tile, warp) partials in
 * gn_scratch (>= glare_conv_gn_scratch_floats(B, Hout, Wout) floats; every entry written once, no atomics) and a second small launch
 * reduces them in fp64 in a fixed order. */
int glare_conv2d_nhwc_tc_g(int mode, int kind, const void* x, const void* x_lo, const void* w, const void* w_lo, const float* bias,
                           const float* residual, float* y, int B, int Hin, int Win, int Cin, int Cout, int ksize, int pa, int pb,
                           double* gn_stats, float* gn_scratch, long long gn_scratch_floats, cudaStream_t stream);
long long glare_conv_gn_scratch_floats(int B, int H, int W);
/* extended form: ldy = output pixel stride (elements, >= Cout, % 4 == 0); w_batch_stride != 0 -> per-sample weights
 * w + n * w_batch_stride (the attention GEMMs S = Q K^T and O = P V, encoder_decoder.py:176-187) */
int glare_conv2d_nhwc_tc_ex(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo, const float* bias,
                            const float* residual, float* y, int B, int H, int W, int Cin, int Cout, int ksize, long long ldy,
                            long long w_batch_stride, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (4b) AttnBlock core -- encoder_decoder.py:176-187: softmax(q^T k * C^-0.5) and h = v w^T around the two GEMMs above.
 * ---------------------------------------------------------------------------------------------------- */
/* S [rows][lds] fp32 logits -> P [rows][ldp] = softmax(scale*S[:, :n_keys]) as operand (0 bf16 | 1 fp32 | 2 tf32 hi+lo),
 * zero in [n_keys, n_pad) */
int glare_attn_softmax_rows(int out_mode, const float* S, long long rows, long long lds, int n_keys, int n_pad, float scale,
                            void* out_hi, void* out_lo, long long ldp, cudaStream_t stream);
/* v NHWC [B][N][C] fp32 -> V^T [B][C][Np] operand(s) */
int glare_attn_transpose_v(int out_mode, const float* v, int B, int N, int C, int Np, void* out_hi, void* out_lo,
                           cudaStream_t stream);
/* Softmax fused into the epilogues of the two GEMMs (mode 4 = bf16x3 operands; same reference lines).  The scores GEMM emits
 * p~ = exp(scale * q.k - ref(row)) with ref(row) = scale * |q_row| * max_j |k_j| - margin (an upper bound of the row's logits minus
 * margin, so no row maximum is needed before the GEMM) directly as the operand of the second GEMM, plus partial row sums; the second
 * GEMM scales row i by 1 / sum_j p~_ij.  A row whose sum leaves [1e-24, 3e38] sets *flag (the caller re-runs with the exact path). */
int glare_attn_row_norm(const float* x, long long rows, int C, long long rows_per_sample, float* norm_out, unsigned* max_bits,
                        cudaStream_t stream);
/* Default row reference (round 2): ref_out[r] = scale * max_{j < n_sub} s_sub[r * lds + j] + offset, from the scores of every query against a
 * strided SAMPLE of the keys (one small GEMM).  The sampled maximum is a lower bound of the row maximum, so the fused path only leaves its
 * window when the true maximum exceeds the sampled one by > ~128 nats, however loose |q||k| is.  Passed to glare_attn_scores_exp_tc as
 * q_row_norm with key_norm_max = NULL. */
int glare_attn_row_ref(const float* s_sub, long long rows, long long lds, int n_sub, float scale, float offset, float* ref_out,
                       cudaStream_t stream);
int glare_attn_scores_exp_tc(int mode, const void* q, const void* k, int rows_h, int rows_w, int C, int n_keys, int n_pad, float scale,
                             float margin, const float* q_row_norm, const unsigned* key_norm_max, void* p_out, float* row_sum_part,
                             long long part_stride, int* n_blocks_host, cudaStream_t stream);
int glare_attn_row_sum_finish(const float* part, long long part_stride, int n_blocks, long long rows, float* row_scale, int* flag,
                              cudaStream_t stream);
int glare_attn_pv_tc(int mode, const void* p, long long ldp, const void* vt, long long ldvt, const float* row_scale, const float* residual,
                     void* y, int rows_h, int rows_w, int n_keys, int C, long long ldy, int pack_out, cudaStream_t stream);
/* stride-1 conv writing its output as the bf16x3 operand of the next tensor-core GEMM (q / k projections, the WarpBlock offset conv) and,
 * optionally, per-row partial sums of squares per output block; glare_attn_row_norm_finish turns those into |row| and the per-sample maximum */
int glare_conv2d_nhwc_tc_pack(int mode, const void* x, const void* w, const float* bias, void* y_operand, int B, int H, int W, int Cin, int Cout,
                              int ksize, float* row_sq_part, long long part_stride, int* n_blocks_host, cudaStream_t stream);
int glare_attn_row_norm_finish(const float* part, long long part_stride, int n_blocks, long long rows, long long rows_per_sample,
                               float* norm_out, unsigned* max_bits, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (5) Normalize = GroupNorm(32, eps 1e-6) (+ swish) -- encoder_decoder.py:29-35 as used by ResnetBlock.forward
 *     (:117-137), AttnBlock.forward (:168-171) and the norm_out heads.  NHWC fp32 in; the apply pass emits the
 *     tensor-core operand of the following conv (out_mode 0 bf16, 1 fp32, 2 tf32 hi + lo).
 * ---------------------------------------------------------------------------------------------------- */
int glare_gn_stats_nhwc_f32(const float* x, int B, long long HW, int C, int G, double* stats, cudaStream_t stream);
int glare_gn_apply_nhwc(int out_mode, const float* x, const double* stats, const float* gamma, const float* beta, float eps,
                        int swish, int B, long long HW, int C, int G, void* out_hi, void* out_lo, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (6) Pre/post-processing of the entry points, batched on the device -- infer_dataset_lol.py:124-128,135 and
 *     infer_unpaired.py:40-42,81-88,121-122,130.  mode 0 = np.pad 'reflect' (LOL eval), 1 = cv2.BORDER_REFLECT (unpaired).
 * ---------------------------------------------------------------------------------------------------- */
int glare_preprocess_u8(const uint8_t* img_nhwc, int B, int H, int W, int pad_top, int pad_bottom, int pad_left, int pad_right, int mode,
                        float* out_nchw, cudaStream_t stream);
int glare_postprocess_u8(const float* y, long long sb, long long sc, long long sh, long long sw, int B, int y0, int x0, int H, int W,
                         uint8_t* out_nhwc, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (2b) Stage-2 training support for the flow (train_stage2.py / LLFlowVQGAN2_arch.py:75-122): forward activations of one coupling
 *      net and the backward pass of FlowStep.normal_flow (FlowStep.py:75-98) -- the autograd of FlowActNorms.py:48-100,
 *      Permutations.py:21-59, FlowAffineCouplingsAblation.py:50-151, flow.py:13-70 -- over NHWC-flattened [P][C] fp32 buffers,
 *      P = B*h*w.  Formulas: oracle/flow_backward.py (checked against autograd on the CPU).  Weight gradients are formed from the
 *      buffers below with glare_im2col_t_operand_bf16x3 + glare_conv2d_nhwc_tc_ex (tensor cores), glare_gemm_tn_skinny_f32 /
 *      glare_dcnv2_bwd_weight_f32 (grad[M][N] += a[P][M]^T b[P][N]) and glare_flow_train_colsum_f32; the host side is
 *      glare_b200/flow_train.py.  Green on B200 (tests/flow_train_gpu_check.py).
 * ---------------------------------------------------------------------------------------------------- */
int glare_flow_train_net_fwd_f32(const float* pre, long long pre_ld, const float* z1, long long z1_ld, const float* net, int B, int h, int w,
                                 float* h1, float* h2, float* hout, cudaStream_t stream);
int glare_flow_train_point_fwd_f32(const float* z_in, const float* pw_fwd, const float* hF, int B, int h, int w, float* t, float* u, float* v,
                                   cudaStream_t stream);
int glare_flow_train_coupling_bwd_f32(int which, const float* g_in, const float* g_z1, const float* x, const float* hraw, float g_ld, int B, int h,
                                      int w, float* g_h, float* g_x, cudaStream_t stream);
int glare_flow_train_net_bwd_f32(const float* g_h, const float* h1, const float* h2, const float* net, int B, int h, int w, float* g_a3, float* g_n2,
                                 float* g_a2, float* g_n1, float* g_a1, float* g_pre, long long pre_ld, float* g_z1, cudaStream_t stream);
int glare_flow_train_point_bwd_f32(const float* g_u, const float* t, const float* pw_fwd, int B, int h, int w, float* g_z, float* sums,
                                   cudaStream_t stream);
int glare_flow_train_im2col3x3_f32(const float* x, long long ldx, int C, int relu, int B, int h, int w, float* col, cudaStream_t stream);
int glare_flow_train_colsum_f32(const float* a, long long lda, const float* b, long long ldb, int C, long long P, float* out, cudaStream_t stream);
/* out [M][N] += a [P][M]^T b [P][N], fp32, for a skinny left operand (M <= 32; N <= 256, 256 % N == 0): the weight gradient of the z1 input
 * channel of a coupling net's first conv (FlowAffineCouplingsAblation.py:143-151 under autograd) and of the 3-channel conv_in (27 columns). */
int glare_gemm_tn_skinny_f32(const float* a, const float* b, long long P, int M, int N, float* out, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (5b) Backward of the condition encoder's blocks for stage-2 training (autograd of encoder_decoder.py:29-35, 68-72, 117-137, 168-192):
 *      GroupNorm (+ swish) backward, the im2col operand of a conv weight gradient (dW = col^T dY through glare_dcnv2_bwd_weight_f32), and the
 *      softmax backward of AttnBlock.  Data gradients of the convolutions and the attention matmuls reuse the tensor-core conv path with
 *      flipped / transposed operands (glare_b200/encoder_train.py).  Verified on the CPU (kernel source run on the host) and on B200 (tests/flow_train_gpu_check.py).
 * ---------------------------------------------------------------------------------------------------- */
int glare_gn_bwd_nhwc_f32(const float* x, const float* gy, const double* stats, const float* gamma, const float* beta, float eps, int swish, int B,
                          long long HW, int C, int G, double* sums, float* gx, float* dgamma, float* dbeta, cudaStream_t stream);
int glare_im2col_nhwc_f32(const float* x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, float* col, cudaStream_t stream);
/* Transposed bf16x3 operand of im2col(x) for the tensor-core weight gradient (torch autograd of the encoder's nn.Conv2d layers,
 * encoder_decoder.py:68-72,88-115,146-165): x NHWC [B,H,W,C] fp32, C % 32 == 0 -> out [nch][k*k*C][2*chunk] bf16, row = tap*C + c, the
 * B*Ho*Wo output pixels as the K dimension in chunks of `chunk` (% 32 == 0; zero past the end).  k = 1: the transposed operand of dY. */
int glare_im2col_t_operand_bf16x3(const float* x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, int chunk,
                                  void* out, cudaStream_t stream);
/* The same for single-piece bf16 operands (mode 0, the bf16 training configuration): out [nch][k*k*C][chunk] bf16, chunk % 64 == 0. */
int glare_im2col_t_operand_bf16(const float* x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, int chunk,
                                void* out, cudaStream_t stream);
int glare_attn_softmax_bwd_f32(const float* P, const float* dP, long long rows, long long ld, int n_keys, float scale, float* dS,
                               cudaStream_t stream);
/* out[C] += column sums of x [P][C] fp32 (C % 4 == 0): the bias gradient of nn.Conv2d (dY summed over the pixels of the batch). */
int glare_colsum_f32(const float* x, long long P, int C, float* out, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (7) Elementwise glue of the AFT decoder -- deformableDecoder_arch.py:587-590 (Mix: enc * m + h * (1 - m)) and :567
 *     (h + x_vq * mean(h) / mean(x_vq)): out[n][i] = a[n][i] * alpha[n * alpha_stride] + b[n][i] * beta[n * beta_stride], each product
 *     rounded before the sum like the reference's separate ATen kernels.  Strides 0 (one scalar) or 1 (per sample).
 * ---------------------------------------------------------------------------------------------------- */
int glare_aft_axpby_f32(const float* a, const float* b, const float* alpha, const float* beta, int alpha_stride, int beta_stride, int B,
                        long long n_per_sample, float* out, cudaStream_t stream);
/* WarpBlock.forward's torch.cat([x_vq, h], dim=1) (deformableDecoder_arch.py:286) fused with the operand conversion of the offset conv that
 * consumes it: a NHWC [P][Ca], b NHWC [P][Cb] fp32 -> bf16x3 (mode 4) operand of the concatenated tensor, out [P][2 * (Ca + Cb)] bf16. */
int glare_aft_cat_operand(const float* a, const float* b, long long P, int Ca, int Cb, void* out, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * (8) Stage-3 loss: one level of MS-SSIM -- modules/pytorch_msssim/__init__.py:20-68 (`ssim`, 'valid' 11 x 11 Gaussian window per plane,
 *     C1 = (0.01 L)^2, C2 = (0.03 L)^2), called per level by `msssim` :71-97 from VQLLFLOWD_model.py:221.  x, y: [planes][H][W] fp32
 *     (planes = B * C), window_host: the `ws` 1-D Gaussian weights in HOST memory (ws = min(11, H, W)).
 *     fwd: part[2 * b], part[2 * b + 1] = sums of cs_map / ssim_map over CTA b's window positions; glare_ssim_partials gives the CTA count.
 *     bwd: coef (device) = {dL/d(sum cs_map), dL/d(sum ssim_map)}; g_mu / g_e11 / g_e12: scratch [planes][H-ws+1][W-ws+1];
 *          dx[planes][H][W] = dL/dx of this level + 0.25 * coarse[planes][H/2][W/2] (the next level's dx through avg_pool2d(2); may be NULL).
 * ---------------------------------------------------------------------------------------------------- */
long long glare_ssim_partials(int planes, int H, int W, int ws);
int glare_ssim_fwd_f32(const float* x, const float* y, int planes, int H, int W, int ws, const float* window_host, float C1, float C2,
                       float* part, cudaStream_t stream);
int glare_ssim_bwd_f32(const float* x, const float* y, int planes, int H, int W, int ws, const float* window_host, float C1, float C2,
                       const float* coef, float* g_mu, float* g_e11, float* g_e12, const float* coarse, float* dx, cudaStream_t stream);

/* (8b) The elementwise / pooling pieces between the convolutions of the stage-3 step, fp32, NHWC tensors with C % 4 == 0:
 *   glare_relu_f32            gy == NULL: out = max(x, 0) (nn.ReLU of vgg16.features);  gy != NULL: out = gy where x > 0 (x = the forward OUTPUT)
 *   glare_maxpool2_nhwc_f32   nn.MaxPool2d(2, 2): y [B][H/2][W/2][C], idx = one byte per output element (position 0..3 in the window;
 *                             the first maximum wins, NaN propagates), B * (H/2) * (W/2) * C bytes
 *   glare_maxpool2_nhwc_bwd_f32  gx [B][H][W][C] = gy routed to the recorded positions, zero elsewhere
 *   glare_avgpool2_f32        F.avg_pool2d(x, (2, 2)) on [planes][H][W] (pytorch_msssim/__init__.py:83-84)
 *   glare_up2_nhwc_f32        nearest x2 of Upsample.forward (encoder_decoder.py:46-48) on [B][H][W][C] (adjoint = 0) and its adjoint, the
 *                             2 x 2 sums of [B][2H][2W][C] (adjoint = 1); H, W are the small tensor's size in both directions */
int glare_relu_f32(const float* x, const float* gy, long long n, float* out, cudaStream_t stream);
int glare_maxpool2_nhwc_f32(const float* x, int B, int H, int W, int C, float* y, void* idx, cudaStream_t stream);
int glare_maxpool2_nhwc_bwd_f32(const float* gy, const void* idx, int B, int H, int W, int C, float* gx, cudaStream_t stream);
int glare_avgpool2_f32(const float* x, long long planes, int H, int W, float* y, cudaStream_t stream);
int glare_up2_nhwc_f32(const float* x, int B, int H, int W, int C, int adjoint, float* y, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GLARE_B200_H_ */
