// Conditional normalizing-flow step kernels (one launch per FlowStep, both directions).
//
// Replaces, per step, reference code/models/modules/FlowStep.py:75-119 (normal_flow / reverse_flow):
//   ActNorm2d             FlowActNorms.py:48-100        z <- (z + b) * exp(logs)      | z * exp(-logs) - b
//   InvertibleConv1x1     Permutations.py:21-59         z <- W z                      | W^-1 z  (fp64 inverse on host)
//   CondAffineSeparatedAndCond  FlowAffineCouplingsAblation.py:50-151
//       feature affine    z  <- (z + shiftFt) * scaleFt | z / scaleFt - shiftFt        (scale/shift = NN_F(ft))
//       self affine       z2 <- (z2 + shift) * scale    | z2 / scale - shift           (scale/shift = NN_A(cat[z1, ft]))
//   NN = Conv2d 3x3 (no bias) + ActNorm + ReLU -> Conv2d 1x1 + ActNorm + ReLU -> Conv2dZeros 3x3 * exp(3 logs)
//                          FlowAffineCouplingsAblation.py:143-151, flow.py:13-70
//
// Work split (DESIGN.md "flow"): the first 3x3 conv of both nets is linear in its input, and 64 of its
// 65 (NN_A) / all 64 (NN_F) input channels are the conditioning features `ft`, which do not depend on z.
// Those contributions ("pre-activation planes" p, 64 channels per net per step) are produced for all 24
// coupling steps at once by one dense 64 -> 3072 conv over ft (tensor-core conv path).  The kernel here
// fuses everything that remains of a step into ONE launch on a 14x14 pixel tile with a 1-pixel halo:
//   (a) h1 = relu(actnorm1(p + conv3x3_{1->64}(z1)))      16x16 region, thread per pixel
//   (b) h2 = relu(actnorm2(W2 h1))                        1x1, in place in shared memory (64 KB)
//   (c) hA = (conv3x3_{64->4}(h2) + b3) * exp(3 logs3)    14x14 interior
//   (d) coupling, feature affine, 1x1 invertible conv, ActNorm, per-sample logdet
// The same kernel in RAW mode (no z) evaluates NN_F's tail for all steps in one batched launch, because
// NN_F never sees z.  All arithmetic is fp32 FMA on CUDA cores: K=64 contractions on a 3-channel latent
// feeding an argmin are not tensor-core work, and z must stay fp32-faithful for the VQ lookup.
#include "common.cuh"

namespace glare {

constexpr int FLOW_TILE = 14;                 // output pixels per tile edge
constexpr int FLOW_REG = 16;                  // region edge = tile + 1-pixel halo on each side
constexpr int FLOW_THREADS = FLOW_REG * FLOW_REG;
constexpr int FLOW_C = 64;                    // hidden channels (LOL.yml network_G.flow hidden_channels)

// packed per-net parameter block (floats); mirrored by glare_b200/flow.py and include/glare_b200.h
constexpr int NET_W1Z = 0;                    // [64][9]   weights of the z1 input channel (zeros for NN_F)
constexpr int NET_B1 = 576;                   // [64]      ActNorm bias
constexpr int NET_S1 = 640;                   // [64]      exp(ActNorm logs)
constexpr int NET_W2T = 704;                  // [64 in][64 out]
constexpr int NET_B2 = 4800;
constexpr int NET_S2 = 4864;
constexpr int NET_W3 = 4928;                  // [64][9][8]  (out channel padded to 8)
constexpr int NET_B3 = 9536;                  // [8]
constexpr int NET_S3 = 9544;                  // [8]  exp(3 * logs)
constexpr int NET_FLOATS = 9552;
constexpr int PW_FLOATS = 16;                 // M[9] row-major (out,in), an_bias[3], an_scale[3], pad

enum { FLOW_RAW = 0, FLOW_INV = 1, FLOW_FWD = 2 };

struct FlowArgs {
    const float* p;          // pre-activation planes of this net: [b][step][64][h*w] through the strides below
    long long p_bs, p_ss, p_cs, p_ps;   // batch / step / channel / pixel strides (NCHW planes: cs = h*w, ps = 1; NHWC: cs = 1, ps = C)
    const float* net;        // packed nets, one per step
    long long net_ss;
    float* out;              // RAW: [b][step][nout][h*w]
    long long out_bs, out_ss;
    const float* z_in;       // [B,3,h,w]
    float* z_out;
    const float* hF;         // NN_F output of this step: [b][6][h*w]
    long long hF_bs;
    const float* pw;         // PW_FLOATS
    float* logdet;           // [B] or null
    int B, h, w, tiles_x;
};

__device__ __forceinline__ float scale_of(float hraw) {
    // FlowAffineCouplingsAblation.py:124-135: sigmoid(h + 2) + affine_eps(1e-4)
    return 1.0f / (1.0f + expf(-(hraw + 2.0f))) + 0.0001f;
}

template <int MODE, int NOUT>
__global__ void __launch_bounds__(FLOW_THREADS, 2) flow_tail_kernel(const FlowArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_net = reinterpret_cast<float*>(smem_raw);                 // NET_FLOATS
    float* s_h = s_net + NET_FLOATS;                                   // [64][256]
    float* s_z1 = s_h + FLOW_C * FLOW_THREADS;                         // [18][20]
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_red[FLOW_THREADS / 32];
    __shared__ float pwv[PW_FLOATS];

    const int tid = threadIdx.x;
    if (MODE != FLOW_RAW && tid < PW_FLOATS) pwv[tid] = __ldg(a.pw + tid);
    const int b = blockIdx.y, step = blockIdx.z;
    const int ty = blockIdx.x / a.tiles_x, tx = blockIdx.x - ty * a.tiles_x;
    const int r = tid >> 4, c = tid & 15;
    const int gy = ty * FLOW_TILE - 1 + r, gx = tx * FLOW_TILE - 1 + c;
    const bool in_img = gy >= 0 && gy < a.h && gx >= 0 && gx < a.w;
    const long long hw = (long long)a.h * a.w;
    const long long pix = (long long)gy * a.w + gx;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, NET_FLOATS * 4);
        bulk_g2s(s_net, a.net + (long long)step * a.net_ss, NET_FLOATS * 4, &bar);
    }

    if (MODE != FLOW_RAW) {
        // z1 on the region plus one more halo pixel (18x18); zero outside the image (conv zero padding)
        const float* zb = a.z_in + (long long)b * 3 * hw;
        const float* hfb = a.hF + (long long)b * a.hF_bs;
        for (int i = tid; i < 18 * 18; i += FLOW_THREADS) {
            const int rr = i / 18, cc = i - rr * 18;
            const int y = ty * FLOW_TILE - 2 + rr, x = tx * FLOW_TILE - 2 + cc;
            float v = 0.f;
            if (y >= 0 && y < a.h && x >= 0 && x < a.w) {
                const long long q = (long long)y * a.w + x;
                if (MODE == FLOW_INV) {
                    v = zb[q];
                } else {
                    // forward: NN_A sees channel 0 after ActNorm, W and the feature affine (FlowStep.py:75-98)
                    float t0 = (zb[q] + pwv[9]) * pwv[12];
                    float t1 = (zb[hw + q] + pwv[10]) * pwv[13];
                    float t2 = (zb[2 * hw + q] + pwv[11]) * pwv[14];
                    float u = pwv[0] * t0 + pwv[1] * t1 + pwv[2] * t2;
                    v = (u + hfb[q]) * scale_of(hfb[hw + q]);
                }
            }
            s_z1[rr * 20 + cc] = v;
        }
        __syncthreads();
    }

    // ---- (a) first layer: pre-activation plane + z1 contribution, ActNorm, ReLU ------------------------
    float zc[9];
    if (MODE != FLOW_RAW) {
#pragma unroll
        for (int t = 0; t < 9; ++t) zc[t] = s_z1[(r + t / 3) * 20 + (c + t % 3)];
    }
    const float* pb = a.p + (long long)b * a.p_bs + (long long)step * a.p_ss + (in_img ? pix * a.p_ps : 0);
    mbar_wait(&bar, 0);
#pragma unroll 1
    for (int c0 = 0; c0 < FLOW_C; c0 += 8) {
        float pv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pv[j] = in_img ? __ldg(pb + (long long)(c0 + j) * a.p_cs) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = pv[j];
            if (MODE != FLOW_RAW) {
                const float* wz = s_net + NET_W1Z + (c0 + j) * 9;
#pragma unroll
                for (int t = 0; t < 9; ++t) v = fmaf(wz[t], zc[t], v);
            }
            v = (v + s_net[NET_B1 + c0 + j]) * s_net[NET_S1 + c0 + j];
            s_h[(c0 + j) * FLOW_THREADS + tid] = fmaxf(v, 0.f);
        }
    }

    // ---- (b) 1x1 conv 64 -> 64 + ActNorm + ReLU, in place (each thread owns its pixel column) -----------
    {
        float acc[FLOW_C];
#pragma unroll
        for (int o = 0; o < FLOW_C; ++o) acc[o] = 0.f;
#pragma unroll 2
        for (int ci = 0; ci < FLOW_C; ++ci) {
            const float x = s_h[ci * FLOW_THREADS + tid];
            const float4* w4 = reinterpret_cast<const float4*>(s_net + NET_W2T + ci * FLOW_C);
#pragma unroll
            for (int o4 = 0; o4 < FLOW_C / 4; ++o4) {
                const float4 wv = w4[o4];
                acc[4 * o4 + 0] = fmaf(wv.x, x, acc[4 * o4 + 0]);
                acc[4 * o4 + 1] = fmaf(wv.y, x, acc[4 * o4 + 1]);
                acc[4 * o4 + 2] = fmaf(wv.z, x, acc[4 * o4 + 2]);
                acc[4 * o4 + 3] = fmaf(wv.w, x, acc[4 * o4 + 3]);
            }
        }
        // outside the image h2 is the zero padding of the last 3x3 conv
#pragma unroll
        for (int o = 0; o < FLOW_C; ++o) {
            const float v = (acc[o] + s_net[NET_B2 + o]) * s_net[NET_S2 + o];
            s_h[o * FLOW_THREADS + tid] = in_img ? fmaxf(v, 0.f) : 0.f;
        }
    }
    __syncthreads();

    // ---- (c) 3x3 conv 64 -> NOUT on the interior, (d) step epilogue --------------------------------------
    const bool interior = r >= 1 && r <= FLOW_TILE && c >= 1 && c <= FLOW_TILE;
    const bool active = interior && in_img;
    float ld = 0.f;
    if (active) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll 2
        for (int ci = 0; ci < FLOW_C; ++ci) {
            const float* hrow = s_h + ci * FLOW_THREADS + (r - 1) * FLOW_REG + (c - 1);
            const float4* w4 = reinterpret_cast<const float4*>(s_net + NET_W3 + ci * 72);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float x = hrow[(t / 3) * FLOW_REG + (t % 3)];
                const float4 wa = w4[2 * t];
                o[0] = fmaf(wa.x, x, o[0]);
                o[1] = fmaf(wa.y, x, o[1]);
                o[2] = fmaf(wa.z, x, o[2]);
                o[3] = fmaf(wa.w, x, o[3]);
                if (NOUT > 4) {
                    const float4 wb = w4[2 * t + 1];
                    o[4] = fmaf(wb.x, x, o[4]);
                    o[5] = fmaf(wb.y, x, o[5]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NOUT; ++j) o[j] = (o[j] + s_net[NET_B3 + j]) * s_net[NET_S3 + j];

        if (MODE == FLOW_RAW) {
            float* ob = a.out + (long long)b * a.out_bs + (long long)step * a.out_ss + pix;
#pragma unroll
            for (int j = 0; j < NOUT; ++j) ob[(long long)j * hw] = o[j];
        } else {
            const float* zb = a.z_in + (long long)b * 3 * hw + pix;
            const float* hfb = a.hF + (long long)b * a.hF_bs + pix;
            float hf[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) hf[j] = __ldg(hfb + (long long)j * hw);
            float z0 = zb[0], z1 = zb[hw], z2 = zb[2 * hw];
            const float sc0 = scale_of(o[1]), sc1 = scale_of(o[3]);
            const float sf0 = scale_of(hf[1]), sf1 = scale_of(hf[3]), sf2 = scale_of(hf[5]);
            if (MODE == FLOW_INV) {
                z1 = z1 / sc0 - o[0];                       // FlowAffineCouplingsAblation.py:88-91
                z2 = z2 / sc1 - o[2];
                z0 = z0 / sf0 - hf[0];                      // :106-108
                z1 = z1 / sf1 - hf[2];
                z2 = z2 / sf2 - hf[4];
                float y0 = pwv[0] * z0 + pwv[1] * z1 + pwv[2] * z2;   // Permutations.py:55-56 (W^-1)
                float y1 = pwv[3] * z0 + pwv[4] * z1 + pwv[5] * z2;
                float y2 = pwv[6] * z0 + pwv[7] * z1 + pwv[8] * z2;
                z0 = y0 * pwv[12] - pwv[9];                 // FlowActNorms.py:64,98-99 (scale = exp(-logs))
                z1 = y1 * pwv[13] - pwv[10];
                z2 = y2 * pwv[14] - pwv[11];
                ld = -(logf(sc0) + logf(sc1) + logf(sf0) + logf(sf1) + logf(sf2));
            } else {
                z0 = (z0 + pwv[9]) * pwv[12];
                z1 = (z1 + pwv[10]) * pwv[13];
                z2 = (z2 + pwv[11]) * pwv[14];
                float y0 = pwv[0] * z0 + pwv[1] * z1 + pwv[2] * z2;
                float y1 = pwv[3] * z0 + pwv[4] * z1 + pwv[5] * z2;
                float y2 = pwv[6] * z0 + pwv[7] * z1 + pwv[8] * z2;
                z0 = (y0 + hf[0]) * sf0;                    // FlowAffineCouplingsAblation.py:56-59
                z1 = (y1 + hf[2]) * sf1;
                z2 = (y2 + hf[4]) * sf2;
                z1 = (z1 + o[0]) * sc0;                     // :75-78
                z2 = (z2 + o[2]) * sc1;
                ld = logf(sc0) + logf(sc1) + logf(sf0) + logf(sf1) + logf(sf2);
            }
            float* zo = a.z_out + (long long)b * 3 * hw + pix;
            zo[0] = z0;
            zo[hw] = z1;
            zo[2 * hw] = z2;
        }
    }
    if (MODE != FLOW_RAW && a.logdet != nullptr) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, s);
        if ((tid & 31) == 0) s_red[tid >> 5] = ld;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < FLOW_THREADS / 32; ++i) t += s_red[i];
            atomicAdd(a.logdet + b, t);
        }
    }
}

// noCoupling steps (FlowUpsamplerNet.py:95-106): ActNorm + invertible 1x1 only
template <bool INV>
__global__ void __launch_bounds__(256) flow_pointwise_kernel(const float* __restrict__ z_in, float* __restrict__ z_out,
                                                             const float* __restrict__ pw, int B, long long hw) {
    float pwv[PW_FLOATS];
#pragma unroll
    for (int i = 0; i < PW_FLOATS; ++i) pwv[i] = __ldg(pw + i);
    const long long n = (long long)B * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / hw, q = i - b * hw;
        const float* zb = z_in + b * 3 * hw + q;
        float z0 = zb[0], z1 = zb[hw], z2 = zb[2 * hw];
        if (INV) {
            float y0 = pwv[0] * z0 + pwv[1] * z1 + pwv[2] * z2;
            float y1 = pwv[3] * z0 + pwv[4] * z1 + pwv[5] * z2;
            float y2 = pwv[6] * z0 + pwv[7] * z1 + pwv[8] * z2;
            z0 = y0 * pwv[12] - pwv[9];
            z1 = y1 * pwv[13] - pwv[10];
            z2 = y2 * pwv[14] - pwv[11];
        } else {
            z0 = (z0 + pwv[9]) * pwv[12];
            z1 = (z1 + pwv[10]) * pwv[13];
            z2 = (z2 + pwv[11]) * pwv[14];
            float y0 = pwv[0] * z0 + pwv[1] * z1 + pwv[2] * z2;
            float y1 = pwv[3] * z0 + pwv[4] * z1 + pwv[5] * z2;
            float y2 = pwv[6] * z0 + pwv[7] * z1 + pwv[8] * z2;
            z0 = y0; z1 = y1; z2 = y2;
        }
        float* zo = z_out + b * 3 * hw + q;
        zo[0] = z0;
        zo[hw] = z1;
        zo[2 * hw] = z2;
    }
}

constexpr size_t FLOW_SMEM = (size_t)(NET_FLOATS + FLOW_C * FLOW_THREADS + 18 * 20) * sizeof(float);

template <int MODE, int NOUT>
static int launch_flow_tail(const FlowArgs& a, int n_steps, cudaStream_t stream) {
    GLARE_CUDA(cudaFuncSetAttribute(flow_tail_kernel<MODE, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)FLOW_SMEM));
    const int tiles_y = (a.h + FLOW_TILE - 1) / FLOW_TILE;
    dim3 grid(a.tiles_x * tiles_y, a.B, n_steps);
    flow_tail_kernel<MODE, NOUT><<<grid, FLOW_THREADS, FLOW_SMEM, stream>>>(a);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

}  // namespace glare

using namespace glare;

GLARE_API int glare_flow_net_floats(void) { return NET_FLOATS; }

// NN tail for `n_steps` nets at once (used for NN_F of all coupling steps: it never sees z).
GLARE_API int glare_flow_cond_tail_f32(const float* p, long long p_batch_stride, long long p_step_stride,
                                        long long p_chan_stride, long long p_pix_stride, const float* nets, int n_steps, int nout, int B, int h, int w, float* out,
                                        long long out_batch_stride, long long out_step_stride, cudaStream_t stream) {
    if (B < 0 || h < 0 || w < 0 || n_steps < 0 || (nout != 4 && nout != 6)) return GLARE_ERR_BAD_ARG;
    if (B == 0 || h == 0 || w == 0 || n_steps == 0) return GLARE_OK;
    if (!p || !nets || !out) return GLARE_ERR_BAD_ARG;
    if (B > 65535 || n_steps > 65535) return GLARE_ERR_BAD_ARG;
    FlowArgs a{};
    a.p = p; a.p_bs = p_batch_stride; a.p_ss = p_step_stride; a.p_cs = p_chan_stride; a.p_ps = p_pix_stride;
    a.net = nets; a.net_ss = NET_FLOATS;
    a.out = out; a.out_bs = out_batch_stride; a.out_ss = out_step_stride;
    a.B = B; a.h = h; a.w = w; a.tiles_x = (w + FLOW_TILE - 1) / FLOW_TILE;
    return nout == 6 ? launch_flow_tail<FLOW_RAW, 6>(a, n_steps, stream) : launch_flow_tail<FLOW_RAW, 4>(a, n_steps, stream);
}

// One FlowStep.  direction: 0 = normal_flow (encode), 1 = reverse_flow (decode).  coupling: 0 for the
// "noCoupling" steps (pA / hF / netA ignored).  z_out must not alias z_in.  logdet ([B], may be null) is
// incremented by the data-dependent terms only (sum log scale; negated in reverse); the constant
// ActNorm / invconv terms are added by the caller (they depend only on the weights).
GLARE_API int glare_flow_step_f32(int direction, int coupling, const float* z_in, float* z_out, const float* pA,
                                   long long pA_batch_stride, long long pA_chan_stride, long long pA_pix_stride, const float* hF, long long hF_batch_stride,
                                   const float* netA, const float* pw, int B, int h, int w, float* logdet,
                                   cudaStream_t stream) {
    if (B < 0 || h < 0 || w < 0 || (direction != 0 && direction != 1)) return GLARE_ERR_BAD_ARG;
    if (B == 0 || h == 0 || w == 0) return GLARE_OK;
    if (!z_in || !z_out || !pw || z_in == z_out) return GLARE_ERR_BAD_ARG;
    if (!coupling) {
        const long long hw = (long long)h * w, n = (long long)B * hw;
        const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
        if (direction == 1) flow_pointwise_kernel<true><<<grid, 256, 0, stream>>>(z_in, z_out, pw, B, hw);
        else flow_pointwise_kernel<false><<<grid, 256, 0, stream>>>(z_in, z_out, pw, B, hw);
        GLARE_CHECK_LAUNCH();
        return GLARE_OK;
    }
    if (!pA || !hF || !netA || B > 65535) return GLARE_ERR_BAD_ARG;
    FlowArgs a{};
    a.p = pA; a.p_bs = pA_batch_stride; a.p_ss = 0; a.p_cs = pA_chan_stride; a.p_ps = pA_pix_stride;
    a.net = netA; a.net_ss = 0;
    a.z_in = z_in; a.z_out = z_out;
    a.hF = hF; a.hF_bs = hF_batch_stride;
    a.pw = pw; a.logdet = logdet;
    a.B = B; a.h = h; a.w = w; a.tiles_x = (w + FLOW_TILE - 1) / FLOW_TILE;
    return direction == 1 ? launch_flow_tail<FLOW_INV, 4>(a, 1, stream) : launch_flow_tail<FLOW_FWD, 4>(a, 1, stream);
}
