// VectorQuantizer2 lookup: per-token L2 distance to the codebook, argmin, embedding gather.
// Replaces reference code/models/modules/quantize.py:271-312 (distance :280-282, argmin :284,
// gather :285, NCHW<->NHWC rearranges :276/:301 fused away).
//
// Bit-exact contract: the distance the reference computes on CPU is
//   d = fl(fl(tn + cn) - fl(2*dot)),  tn/cn = ((v0*v0)+(v1*v1))+(v2*v2) without FMA,
//   dot = fma(t2,c2, fma(t1,c1, fl(t0*c0)));   argmin = lowest index among equal minima.
// 2*dot is exact, so fl(a - 2*dot) == fma(-2, dot, a): one FFMA, same bits.
//
// Mapping: persistent CTAs (<= 148), codebook packed {c0,c1,c2,cn} (128 KB for K=8192) staged once
// per CTA into shared memory with bulk-async (TMA) copies; 8 warps split the K codes, each lane keeps
// TPT tokens in registers (one broadcast LDS.128 per code feeds TPT independent FMA chains); (d,idx)
// pairs are min-reduced across the 8 code slices in ascending slice order with a strict '<', which
// preserves the first-minimum tie-break.
#include "common.cuh"
#include <math_constants.h>

namespace glare {

constexpr int VQ_THREADS = 256;
constexpr int VQ_WARPS = VQ_THREADS / 32;

__global__ void vq_pack_codebook_kernel(const float* __restrict__ cb, float4* __restrict__ packed, int K) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float c0 = cb[3 * k], c1 = cb[3 * k + 1], c2 = cb[3 * k + 2];
    float cn = __fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fmul_rn(c2, c2));
    packed[k] = make_float4(c0, c1, c2, cn);
}

template <int TPT>
__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_argmin_gather_kernel(const float* __restrict__ z,        // [B,3,hw]
                        const float4* __restrict__ packed,  // [K] {c0,c1,c2,|c|^2}
                        int B, int hw, int K, long long* __restrict__ idx_out, float* __restrict__ zq_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* s_cb = reinterpret_cast<float4*>(smem_raw);                          // K float4
    float* s_d = reinterpret_cast<float*>(smem_raw + (size_t)K * sizeof(float4));  // [VQ_WARPS][32*TPT]
    int* s_i = reinterpret_cast<int*>(s_d + VQ_WARPS * 32 * TPT);                // [VQ_WARPS][32*TPT]
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t total = (uint32_t)K * 16u;
        mbar_arrive_expect_tx(&bar, total);
        const uint32_t chunk = 32768u;
        for (uint32_t off = 0; off < total; off += chunk) {
            uint32_t n = total - off < chunk ? total - off : chunk;
            bulk_g2s(smem_raw + off, reinterpret_cast<const unsigned char*>(packed) + off, n, &bar);
        }
    }

    const long long N = (long long)B * hw;
    const int tile_tokens = 32 * TPT;
    const long long n_tiles = (N + tile_tokens - 1) / tile_tokens;
    const int kslice = (K + VQ_WARPS - 1) / VQ_WARPS;
    const int k0 = warp * kslice, k1 = min(K, k0 + kslice);
    bool cb_ready = false;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        float t0[TPT], t1[TPT], t2[TPT], tn[TPT], best[TPT];
        int besti[TPT];
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
            long long t = tile * tile_tokens + j * 32 + lane;
            float a = 0.f, b = 0.f, c = 0.f;
            if (t < N) {
                long long bi = t / hw, p = t - bi * hw;
                const float* zb = z + bi * 3 * hw;
                a = zb[p];
                b = zb[hw + p];
                c = zb[2 * (long long)hw + p];
            }
            t0[j] = a; t1[j] = b; t2[j] = c;
            tn[j] = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
            best[j] = CUDART_INF_F;
            besti[j] = k0;
        }
        if (!cb_ready) {
            mbar_wait(&bar, 0);
            cb_ready = true;
        }
#pragma unroll 4
        for (int k = k0; k < k1; ++k) {
            const float4 c = s_cb[k];
#pragma unroll
            for (int j = 0; j < TPT; ++j) {
                float dot = __fmaf_rn(t2[j], c.z, __fmaf_rn(t1[j], c.y, __fmul_rn(t0[j], c.x)));
                float d = __fmaf_rn(-2.0f, dot, __fadd_rn(tn[j], c.w));
                if (d < best[j]) { best[j] = d; besti[j] = k; }
            }
        }
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
            s_d[warp * tile_tokens + j * 32 + lane] = best[j];
            s_i[warp * tile_tokens + j * 32 + lane] = besti[j];
        }
        __syncthreads();
        for (int tt = tid; tt < tile_tokens; tt += VQ_THREADS) {
            long long t = tile * tile_tokens + tt;
            if (t < N) {
                float bd = s_d[tt];
                int bi_ = s_i[tt];
#pragma unroll
                for (int w = 1; w < VQ_WARPS; ++w) {
                    float d = s_d[w * tile_tokens + tt];
                    if (d < bd) { bd = d; bi_ = s_i[w * tile_tokens + tt]; }
                }
                long long bi = t / hw, p = t - bi * hw;
                if (!(bd < CUDART_INF_F)) {
                    // no finite distance seen (NaN/Inf input): replay torch.argmin semantics exactly --
                    // first NaN wins, else first minimum.
                    const float* zb = z + bi * 3 * hw;
                    float a = zb[p], b = zb[hw + p], c3 = zb[2 * (long long)hw + p];
                    float tnn = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c3, c3));
                    bd = CUDART_INF_F; bi_ = 0;
                    for (int k = 0; k < K; ++k) {
                        const float4 c = s_cb[k];
                        float dot = __fmaf_rn(c3, c.z, __fmaf_rn(b, c.y, __fmul_rn(a, c.x)));
                        // the reference's two roundings, fl(fl(tn + cn) - fl(2 dot)): when 2 dot overflows while dot does not, this is
                        // inf - inf = NaN where the single FMA of the main loop gives inf.  Only here can that matter: fl(2 dot) = inf
                        // implies tn = inf, so the main loop (d < best) never selects such a code either way.
                        float d = __fsub_rn(__fadd_rn(tnn, c.w), __fmul_rn(2.0f, dot));
                        if (d < bd || (d != d && bd == bd)) { bd = d; bi_ = k; }
                    }
                }
                idx_out[t] = bi_;
                const float4 c = s_cb[bi_];
                // quantize.py:298 straight-through estimator, evaluated in fp32 exactly as the reference does:
                // z_q = z + (e - z)  (two roundings; differs from e in the last bit for some tokens)
                const float* zb = z + bi * 3 * hw;
                const float a = zb[p], b = zb[hw + p], c3 = zb[2 * (long long)hw + p];
                float* q = zq_out + bi * 3 * hw;
                q[p] = __fadd_rn(a, __fsub_rn(c.x, a));
                q[hw + p] = __fadd_rn(b, __fsub_rn(c.y, b));
                q[2 * (long long)hw + p] = __fadd_rn(c3, __fsub_rn(c.z, c3));
            }
        }
        __syncthreads();
    }
    if (!cb_ready) mbar_wait(&bar, 0);  // never leave a bulk copy in flight
}

}  // namespace glare

using namespace glare;

GLARE_API int glare_vq_pack_codebook_f32(const float* codebook, int K, float* packed_out, cudaStream_t stream) {
    if (!codebook || !packed_out || K <= 0) return GLARE_ERR_BAD_ARG;
    vq_pack_codebook_kernel<<<(K + 255) / 256, 256, 0, stream>>>(codebook, reinterpret_cast<float4*>(packed_out), K);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

GLARE_API int glare_vq_argmin_gather_f32(const float* z_nchw, const float* packed_codebook, int B, int hw, int K,
                                          long long* idx_out, float* zq_nchw_out, cudaStream_t stream) {
    if (B < 0 || hw < 0 || K <= 0) return GLARE_ERR_BAD_ARG;
    const long long N = (long long)B * hw;
    if (N == 0) return GLARE_OK;                     // empty batch: nothing to launch (pointers may be null)
    if (!z_nchw || !packed_codebook || !idx_out || !zq_nchw_out) return GLARE_ERR_BAD_ARG;
    // tokens per thread: 4 when there is enough work to fill the chip, else spread thinner
    const int TPT = (N >= (long long)kNumSMs * 128) ? 4 : (N >= (long long)kNumSMs * 64 ? 2 : 1);
    const int tile_tokens = 32 * TPT;
    const size_t smem = (size_t)K * 16 + (size_t)VQ_WARPS * tile_tokens * 8;
    if (smem > 227 * 1024) return GLARE_ERR_BAD_ARG;
    const long long n_tiles = (N + tile_tokens - 1) / tile_tokens;
    const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
#define LAUNCH_VQ(T)                                                                                        \
    do {                                                                                                    \
        GLARE_CUDA(cudaFuncSetAttribute(vq_argmin_gather_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)smem));                                                        \
        vq_argmin_gather_kernel<T><<<grid, VQ_THREADS, smem, stream>>>(                                     \
            z_nchw, reinterpret_cast<const float4*>(packed_codebook), B, hw, K, idx_out, zq_nchw_out);      \
    } while (0)
    if (TPT == 4) LAUNCH_VQ(4);
    else if (TPT == 2) LAUNCH_VQ(2);
    else LAUNCH_VQ(1);
#undef LAUNCH_VQ
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
