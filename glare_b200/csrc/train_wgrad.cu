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

// grid (pixel groups of 32 in packs of 8, C / 32, k * k); 256 threads.  out [nch][k*k*C][2 * chunk] bf16: row m = tap * C + c, pixel j of
// the chunk at (j >> 5) * 64 + (j & 31) (a1) and + 32 (a2) -- the mode-4 operand layout with the pixels as the K dimension.
// One CTA moves a 256-pixel x 32-channel tile of one tap: warp w gathers pixel group w (lane = channel, 128-byte reads per pixel) into shared
// memory, then writes 4 channel rows of 1 KB each as 16-byte stores (lane = 8 consecutive pixels of one group: its a1 and its a2 octet).
// The tile's column index is pix + pix / 8 over an odd row stride, which makes both phases bank-conflict free.  (Round 2's first version
// moved 32 x 32 tiles with 2-byte stores: 147 K CTAs for a 128-channel 256x256 conv, about a tenth of the HBM rate.)
constexpr int WG_PIX = 256, WG_STRIDE = WG_PIX + WG_PIX / 8 + 1;

// X3 = false: the single-piece bf16 operand (mode 0), out [nch][k*k*C][chunk], pixel j of the chunk at element j.
template <bool X3>
__global__ void __launch_bounds__(256) im2col_t_operand_kernel(const float* __restrict__ x, int B, int H, int W, int C, int k, int stride, int pad,
                                                               int Ho, int Wo, long long P, int chunk, long long groups,
                                                               __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[32 * WG_STRIDE];
    const int tap = blockIdx.z, dy = tap / k, dx = tap - dy * k;
    const int c0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // gather: warp = pixel group; lane j resolves pixel j's source once, the warp then reads the 32 channels of each pixel
    {
        const long long p = ((long long)blockIdx.x * 8 + warp) * 32 + lane;
        long long src = -1;
        if (p < P) {
            const long long hw = (long long)Ho * Wo;
            const int b = (int)(p / hw);
            const int rem = (int)(p - (long long)b * hw);
            const int oy = rem / Wo, ox = rem - oy * Wo;
            const int iy = oy * stride + dy - pad, ix = ox * stride + dx - pad;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) src = (((long long)b * H + iy) * W + ix) * C + c0;
        }
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const long long s = __shfl_sync(0xffffffffu, src, j);
            const int pix = warp * 32 + j;
            tile[lane * WG_STRIDE + pix + (pix >> 3)] = s >= 0 ? __ldg(x + s + lane) : 0.f;
        }
    }
    __syncthreads();
    // scatter: lane -> (pixel group g, octet q); rows warp * 4 .. + 3
    const int g = lane >> 2, q = lane & 3;
    const long long gi = (long long)blockIdx.x * 8 + g;
    if (gi >= groups) return;
    const long long p0 = gi * 32;
    const long long ci = p0 / chunk;
    const int j0 = (int)(p0 - ci * chunk);                                  // multiple of 32
    const long long M = (long long)k * k * C;
    const int col = (g * 32 + q * 8) + (g * 4 + q);                           // pix + pix / 8 of the octet's first pixel
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = warp * 4 + rr;
        const float* t = tile + r * WG_STRIDE + col;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (X3) {
                __nv_bfloat16 a1, a2, b1, b2;
                split_b3(t[2 * i], a1, a2);
                split_b3(t[2 * i + 1], b1, b2);
                hi[i] = (uint32_t)__bfloat16_as_ushort(a1) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
                lo[i] = (uint32_t)__bfloat16_as_ushort(a2) | ((uint32_t)__bfloat16_as_ushort(b2) << 16);
            } else {
                hi[i] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(t[2 * i])) |
                        ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(t[2 * i + 1])) << 16);
            }
        }
        if (X3) {
            __nv_bfloat16* row = out + ((ci * M + (long long)tap * C + c0 + r) * chunk + j0) * 2 + q * 8;
            *reinterpret_cast<uint4*>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(row + 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        } else {
            __nv_bfloat16* row = out + (ci * M + (long long)tap * C + c0 + r) * chunk + j0 + q * 8;
            *reinterpret_cast<uint4*>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        }
    }
}

}  // namespace glare

using namespace glare;

template <bool X3>
static int im2col_t_operand(const float* x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, int chunk, void* out,
                            cudaStream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (C & 31) || (k != 1 && k != 3) || stride < 1 || stride > 2 || pad < 0 || Ho <= 0 || Wo <= 0 ||
        chunk <= 0 || (chunk & (X3 ? 31 : 63)))
        return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !out) return GLARE_ERR_BAD_ARG;
    const long long P = (long long)B * Ho * Wo;
    const long long nch = (P + chunk - 1) / chunk;
    const long long groups = nch * (chunk / 32);
    if (groups > 0x7fffffffLL || C / 32 > 65535) return GLARE_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return GLARE_ERR_BAD_ARG;
    im2col_t_operand_kernel<X3><<<dim3((unsigned)((groups + 7) / 8), (unsigned)(C / 32), (unsigned)(k * k)), 256, 0, stream>>>(
        x, B, H, W, C, k, stride, pad, Ho, Wo, P, chunk, groups, reinterpret_cast<__nv_bfloat16*>(out));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// x NHWC [B,H,W,C] fp32 (C % 32 == 0) -> transposed bf16x3 operand of its im2col for a k x k conv (k in {1, 3}) with the given stride and
// low-side padding, output size Ho x Wo: out [nch][k*k*C][2 * chunk] bf16, nch = ceil(B*Ho*Wo / chunk), chunk % 32 == 0; pixels past
// B*Ho*Wo are zero.  Feeds the batched GEMM of glare_conv2d_nhwc_tc_ex (per-sample weights = the dY operand built with k = 1).
GLARE_API int glare_im2col_t_operand_bf16x3(const float* x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, int chunk,
                                            void* out, cudaStream_t stream) {
    return im2col_t_operand<true>(x, B, H, W, C, k, stride, pad, Ho, Wo, chunk, out, stream);
}

// the same for the single-piece bf16 operand (mode 0): out [nch][k*k*C][chunk] bf16, chunk % 64 == 0
GLARE_API int glare_im2col_t_operand_bf16(const float* x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo, int chunk,
                                          void* out, cudaStream_t stream) {
    return im2col_t_operand<false>(x, B, H, W, C, k, stride, pad, Ho, Wo, chunk, out, stream);
}
