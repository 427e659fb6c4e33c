// Weight-gradient operand builder for the stage-2 training step (BASELINE config 4): what torch autograd computes for every nn.Conv2d of
// the condition encoder (encoder_decoder.py:88-115 ResnetBlock convs, :68-72 Downsample, :146-165 AttnBlock 1x1 convs),
//     dW[co][tap * C + c] = sum_p col[p][tap * C + c] * dY[p][co],        col = im2col(x),
// runs on the tensor-core GEMM path as a batch of chunked products  (col_c^T) (dY_c^T)^T  whose two operands must be K-major in the
// REDUCTION index p.  Round 2's first version built them with an fp32 im2col [P][9C] (1.9 GB per 128-channel 320x320 conv), a transposed
// copy and an operand-conversion pass: 55 ms of the 170 ms of kernel time per step (profiles/r51_train_probe_kernel_breakdown.txt).  This
// kernel writes the transposed bf16x3 operand [chunk][tap * C + c][pixel] DIRECTLY from the NHWC activation: one read of x per tap, one operand
// write, no fp32 columns.  With k = 1 it is the plain transposed operand of dY.
#include "common.cuh"

namespace glare {

// grid (pixel groups of 32, C / 32, k * k); block (32, 8).  out [nch][k*k*C][2 * chunk] bf16: row m = tap * C + c, pixel j of the chunk at
// (j >> 5) * 64 + (j & 31) (a1) and + 32 (a2) -- the mode-4 operand layout with the pixels as the K dimension.
__global__ void __launch_bounds__(256) im2col_t_operand_kernel(const float* __restrict__ x, int B, int H, int W, int C, int k, int stride, int pad,
                                                               int Ho, int Wo, long long P, int chunk, __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[32][33];
    const int tap = blockIdx.z, dy = tap / k, dx = tap - dy * k;
    const int c0 = blockIdx.y * 32;
    const long long p0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    // load: 32 pixels x 32 channels, coalesced over the channels of a pixel; taps outside the image and pixels past P read as zero
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        float v = 0.f;
        if (p < P) {
            const long long hw = (long long)Ho * Wo;
            const int b = (int)(p / hw);
            const int rem = (int)(p - (long long)b * hw);
            const int oy = rem / Wo, ox = rem - oy * Wo;
            const int iy = oy * stride + dy - pad, ix = ox * stride + dx - pad;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long long)b * H + iy) * W + ix) * C + c0 + tx);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    // store: row m = tap * C + c, 32 consecutive pixels of one K group: a1 pieces then a2 pieces (64 bytes each)
    const int ci = (int)(p0 / chunk);
    const int j0 = (int)(p0 - (long long)ci * chunk);                       // multiple of 32
    const long long M = (long long)k * k * C;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const float v = tile[tx][r];                                          // pixel tx, channel c0 + r
        __nv_bfloat16 a1, a2;
        split_b3(v, a1, a2);
        __nv_bfloat16* row = out + (((long long)ci * M + (long long)tap * C + c0 + r) * chunk + j0) * 2;
        row[tx] = a1;
        row[32 + tx] = a2;
    }
}

}  // namespace glare

using namespace glare;

// x NHWC [B,H,W,C] fp32 (C % 32 == 0) -> transposed bf16x3 operand of its im2col for a k x k conv (k in {1, 3}) with the given stride and
// low-side padding, output size Ho x Wo: out [nch][k*k*C][2 * chunk] bf16, nch = ceil(B*Ho*Wo / chunk), chunk % 32 == 0; pixels past
// B*Ho*Wo are zero.  Feeds the batched GEMM of glare_conv2d_nhwc_tc_ex (per-sample weights = the dY operand built with k = 1).
GLARE_API int glare_im2col_t_operand_bf16x3(const float* x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, int chunk,
                                            void* out, cudaStream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (C & 31) || (k != 1 && k != 3) || stride < 1 || stride > 2 || pad < 0 || Ho <= 0 || Wo <= 0 ||
        chunk <= 0 || (chunk & 31))
        return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !out) return GLARE_ERR_BAD_ARG;
    const long long P = (long long)B * Ho * Wo;
    const long long nch = (P + chunk - 1) / chunk;
    const long long groups = nch * (chunk / 32);
    if (groups > 0x7fffffffLL || C / 32 > 65535) return GLARE_ERR_UNSUPPORTED;
    im2col_t_operand_kernel<<<dim3((unsigned)groups, (unsigned)(C / 32), (unsigned)(k * k)), dim3(32, 8), 0, stream>>>(
        x, B, H, W, C, k, stride, pad, Ho, Wo, P, chunk, reinterpret_cast<__nv_bfloat16*>(out));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
