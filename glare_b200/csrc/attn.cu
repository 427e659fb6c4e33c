// AttnBlock core (reference encoder_decoder.py:176-187): w = softmax(q^T k * C^-0.5, dim=keys); h = v w^T.
// Single head, d = C = 512, N = h*w tokens (16 275 at 600x400).  The two matmuls run on the tcgen05 GEMM path
// (conv_tc.cu, per-sample weights: S = Q K^T with W = K[n]; O = P V with W = V[n]^T); this file holds the
// memory-bound pieces in between, written so that each N x N matrix is touched once per pass:
//   attn_softmax_rows : reads one band of S (fp32), applies the C^-0.5 scale and the row softmax, and emits P
//                       directly as the tensor-core operand of the second GEMM (bf16 | fp32 | tf32 hi+lo), zero
//                       in the padded key columns.
//   attn_transpose_v  : V [N][C] (NHWC) -> V^T [C][Np] operand(s), zero padded keys.
// The score matrix is processed in bands of query rows sized to stay L2-resident (glare_b200/dense.py).
#include "common.cuh"

namespace glare {

__device__ __forceinline__ float tf32_hi_a(float x) { return tf32_round(x); }

template <int OUT>
__global__ void __launch_bounds__(256) attn_softmax_rows_kernel(const float* __restrict__ S, long long lds, int n_keys,
                                                                int n_pad, float scale, void* __restrict__ out_hi,
                                                                float* __restrict__ out_lo, long long ldp) {
    __shared__ float s_red[8];
    __shared__ float s_bcast;
    const long long row = blockIdx.x;
    const float* s = S + row * lds;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // pass 1: max of the scaled logits (row stays in L1/L2 for the next passes)
    float m = -INFINITY;
    for (int i = tid * 4; i < n_keys; i += 1024) {
        if (i + 3 < n_keys) {
            const float4 v = *reinterpret_cast<const float4*>(s + i);
            m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) );
        } else {
            for (int j = i; j < n_keys; ++j) m = fmaxf(m, s[j]);
        }
    }
    // scale > 0, so max(scale * s) = scale * max(s); the reference multiplies first (encoder_decoder.py:181-182)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    if (tid == 0) {
        float t = s_red[0];
        for (int i = 1; i < 8; ++i) t = fmaxf(t, s_red[i]);
        s_bcast = t * scale;
    }
    __syncthreads();
    const float mx = s_bcast;
    // pass 2: sum of exp
    float sum = 0.f;
    for (int i = tid * 4; i < n_keys; i += 1024) {
        if (i + 3 < n_keys) {
            const float4 v = *reinterpret_cast<const float4*>(s + i);
            sum += expf(v.x * scale - mx) + expf(v.y * scale - mx) + expf(v.z * scale - mx) + expf(v.w * scale - mx);
        } else {
            for (int j = i; j < n_keys; ++j) sum += expf(s[j] * scale - mx);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_red[i];
        s_bcast = 1.0f / t;
    }
    __syncthreads();
    const float inv = s_bcast;
    // pass 3: normalise and emit the operand (n_pad is a multiple of 4)
    for (int i = tid * 4; i < n_pad; i += 1024) {
        float p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = (i + j < n_keys) ? expf(s[i + j] * scale - mx) * inv : 0.f;
        const long long o = row * ldp + i;
        if (OUT == 0) {
            __nv_bfloat162 a = __floats2bfloat162_rn(p[0], p[1]), b = __floats2bfloat162_rn(p[2], p[3]);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&a);
            u.y = *reinterpret_cast<uint32_t*>(&b);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_hi) + o) = u;
        } else if (OUT == 1) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_hi) + o) = make_float4(p[0], p[1], p[2], p[3]);
        } else if (OUT == 4) {
            store_b3_4(reinterpret_cast<__nv_bfloat16*>(out_hi), o, p[0], p[1], p[2], p[3]);
        } else {
            const float4 h = make_float4(tf32_hi_a(p[0]), tf32_hi_a(p[1]), tf32_hi_a(p[2]), tf32_hi_a(p[3]));
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_hi) + o) = h;
            if (OUT == 2) *reinterpret_cast<float4*>(out_lo + o) = make_float4(p[0] - h.x, p[1] - h.y, p[2] - h.z, p[3] - h.w);
            else store_x4(reinterpret_cast<__nv_bfloat16*>(out_lo), o, p[0], p[1], p[2], p[3], h.x, h.y, h.z, h.w);
        }
    }
}

// Same contract, one pass over S: the row (<= 256 threads x 64 values) lives in registers between the max, the
// exponentials (evaluated once) and the normalised operand write.
template <int OUT>
__global__ void __launch_bounds__(256) attn_softmax_rows_reg_kernel(const float* __restrict__ S, long long lds, int n_keys,
                                                                    int n_pad, float scale, void* __restrict__ out_hi,
                                                                    float* __restrict__ out_lo, long long ldp) {
    __shared__ float s_red[8];
    __shared__ float s_bcast;
    const long long row = blockIdx.x;
    const float4* s4 = reinterpret_cast<const float4*>(S + row * lds);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4 v[16];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = (tid + 256 * i) * 4;
        float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (e < n_keys) {
            t = __ldg(s4 + tid + 256 * i);                              // lds % 4 == 0, padded columns are readable
            if (e + 1 >= n_keys) t.y = -INFINITY;
            if (e + 2 >= n_keys) t.z = -INFINITY;
            if (e + 3 >= n_keys) t.w = -INFINITY;
        }
        v[i] = t;
        m = fmaxf(m, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    if (tid == 0) {
        float t = s_red[0];
        for (int i = 1; i < 8; ++i) t = fmaxf(t, s_red[i]);
        s_bcast = t * scale;
    }
    __syncthreads();
    const float mx = s_bcast;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v[i].x = expf(v[i].x * scale - mx);                             // exp(-inf) = 0 for masked / padded columns
        v[i].y = expf(v[i].y * scale - mx);
        v[i].z = expf(v[i].z * scale - mx);
        v[i].w = expf(v[i].w * scale - mx);
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_red[i];
        s_bcast = 1.0f / t;
    }
    __syncthreads();
    const float inv = s_bcast;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = (tid + 256 * i) * 4;
        if (e >= n_pad) continue;
        const float p0 = v[i].x * inv, p1 = v[i].y * inv, p2 = v[i].z * inv, p3 = v[i].w * inv;
        const long long o = row * ldp + e;
        if (OUT == 0) {
            __nv_bfloat162 a = __floats2bfloat162_rn(p0, p1), b = __floats2bfloat162_rn(p2, p3);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&a);
            u.y = *reinterpret_cast<uint32_t*>(&b);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_hi) + o) = u;
        } else if (OUT == 1) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_hi) + o) = make_float4(p0, p1, p2, p3);
        } else if (OUT == 4) {
            store_b3_4(reinterpret_cast<__nv_bfloat16*>(out_hi), o, p0, p1, p2, p3);
        } else {
            const float4 h = make_float4(tf32_hi_a(p0), tf32_hi_a(p1), tf32_hi_a(p2), tf32_hi_a(p3));
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_hi) + o) = h;
            if (OUT == 2) *reinterpret_cast<float4*>(out_lo + o) = make_float4(p0 - h.x, p1 - h.y, p2 - h.z, p3 - h.w);
            else store_x4(reinterpret_cast<__nv_bfloat16*>(out_lo), o, p0, p1, p2, p3, h.x, h.y, h.z, h.w);
        }
    }
}

// v [B][N][C] fp32 -> vt [B][C][Np] operand(s); 32x32 tiles through shared memory
template <int OUT>
__global__ void __launch_bounds__(256) attn_transpose_v_kernel(const float* __restrict__ v, int N, int C, int Np,
                                                               void* __restrict__ out_hi, float* __restrict__ out_lo) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
    const float* vb = v + (long long)b * N * C;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, c = c0 + tx;
        tile[r][tx] = (n < N && c < C) ? __ldg(vb + (long long)n * C + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, n = n0 + tx;
        if (c < C && n < Np) {
            const float x = tile[tx][r];
            const long long o = ((long long)b * C + c) * Np + n;
            if (OUT == 0) reinterpret_cast<__nv_bfloat16*>(out_hi)[o] = __float2bfloat16_rn(x);
            else if (OUT == 1) reinterpret_cast<float*>(out_hi)[o] = x;
            else if (OUT == 4) {
                __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(out_hi);
                const long long xi = (o >> 5) * 64 + (o & 31);
                split_b3(x, xb[xi], xb[xi + 32]);
            } else {
                const float h = tf32_hi_a(x);
                reinterpret_cast<float*>(out_hi)[o] = h;
                if (OUT == 2) out_lo[o] = x - h;
                else {
                    __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(out_lo);
                    const long long xi = (o >> 5) * 64 + (o & 31);
                    xb[xi] = __float2bfloat16_rn(x);
                    xb[xi + 32] = __float2bfloat16_rn(x - h);
                }
            }
        }
    }
}

// ---- softmax fused into the two GEMM epilogues (mode 4) ---------------------------------------------------------------
// softmax is invariant to the per-row reference that is subtracted before exp, and fp32 keeps full relative precision over
// ~76 decades, so the scores GEMM does not need the exact row maximum: any ref(row) with  max_j l_ij - ref(row)  in about
// [-115, +60] (l = scale * s) gives the same P up to rounding.  ref(row) = scale |q_row| max_j |k_j| - margin is a Cauchy-Schwarz
// upper bound of every logit of the row minus `margin`, known BEFORE the GEMM, so its epilogue can emit
//     p~_ij = exp(l_ij - ref_i)     (<= e^margin, the bf16x3 operand of the P V GEMM)   and partial row sums,
// and the P V GEMM's epilogue multiplies row i by 1 / sum_j p~_ij.  The N x N matrix is written once and read once (the
// separate softmax pass read it and wrote it again).  The bound is loose when the row maximum sits far below |q||k|: rows whose sum
// falls under 1e-24 (more than ~115 below the bound) or is not finite set a device flag, and the host re-runs the block with the exact
// three-kernel path (dense.py).
//   attn_row_norm_kernel : |x_row| of q (per row) and max_j |k_j| (per sample, atomicMax on the bits of a non-negative float)
//   attn_row_sum_finish  : row_scale = 1 / sum of the partial row sums, fixed summation order; raises the flag
__global__ void __launch_bounds__(256) attn_row_norm_kernel(const float* __restrict__ x, long long rows, int C, long long rows_per_sample,
                                                            float* __restrict__ norm_out, unsigned* __restrict__ max_bits) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float s = 0.f;
    int i = lane;
    for (; i + 96 < C / 4; i += 128) {                                   // four independent 128-bit loads in flight per lane
        const float4 v0 = __ldg(xr + i), v1 = __ldg(xr + i + 32), v2 = __ldg(xr + i + 64), v3 = __ldg(xr + i + 96);
        s = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v0.z, v0.z, fmaf(v0.w, v0.w, s))));
        s = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, fmaf(v1.z, v1.z, fmaf(v1.w, v1.w, s))));
        s = fmaf(v2.x, v2.x, fmaf(v2.y, v2.y, fmaf(v2.z, v2.z, fmaf(v2.w, v2.w, s))));
        s = fmaf(v3.x, v3.x, fmaf(v3.y, v3.y, fmaf(v3.z, v3.z, fmaf(v3.w, v3.w, s))));
    }
    for (; i < C / 4; i += 32) {
        const float4 v = __ldg(xr + i);
        s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const float nrm = sqrtf(s) * 1.00001f;                       // covers the rounding of the sum and of the GEMM's operands
        if (norm_out) norm_out[row] = nrm;
        if (max_bits) atomicMax(max_bits + row / rows_per_sample, __float_as_uint(nrm));
    }
}

__global__ void __launch_bounds__(256) attn_row_sum_finish_kernel(const float* __restrict__ part, long long part_stride, int n_blocks,
                                                                  long long rows, float* __restrict__ row_scale, int* __restrict__ flag) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += __ldg(part + (long long)b * part_stride + r);
    if (!(s >= 1e-24f && s <= 3e38f)) atomicOr(flag, 1);              // also catches NaN
    row_scale[r] = 1.0f / s;
}

// Row reference from SAMPLED scores (the default since round 2): s_sub [rows][lds] = q_i . k_j for a strided subset of the sample's keys
// (one small GEMM, 128 keys = 1/127 of the scores GEMM at 600x400); ref_i = scale * max_j s_sub[i][j] + offset.  The sampled maximum L_i
// is a LOWER bound of the row maximum m_i, so with offset = 50 the largest p~ of the row is e^(m_i - L_i - 50) >= e^-50 (the row sum
// never underflows) and it overflows only if m_i - L_i > ~128: the fused path now fails only for rows whose true maximum exceeds the
// maximum over 128 sampled keys by more than 128 nats -- independent of how loose |q||k| is (trained checkpoints: norms of tens,
// logits of a few units).  attn_row_sum_finish still raises the flag for such rows and the host falls back to the exact path.
__global__ void __launch_bounds__(256) attn_row_ref_kernel(const float* __restrict__ s_sub, long long rows, long long lds, int n_sub, float scale,
                                                           float offset, float* __restrict__ ref_out) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* sr = s_sub + row * lds;
    float m = -INFINITY;
    for (int i = lane; i < n_sub; i += 32) m = fmaxf(m, __ldg(sr + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) ref_out[row] = fmaf(m, scale, offset);              // NaN scores propagate: the row sum check catches them
}

// |row| from the partial sums of squares a conv epilogue left per (output block, row) (glare_conv2d_nhwc_tc_pack)
__global__ void __launch_bounds__(256) attn_row_norm_finish_kernel(const float* __restrict__ part, long long part_stride, int n_blocks, long long rows,
                                                                   long long rows_per_sample, float* __restrict__ norm_out,
                                                                   unsigned* __restrict__ max_bits) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float nrm = 0.f;
    if (r < rows) {
        float s = 0.f;
        for (int b = 0; b < n_blocks; ++b) s += __ldg(part + (long long)b * part_stride + r);
        nrm = sqrtf(s) * 1.00001f;
        if (norm_out) norm_out[r] = nrm;
    }
    if (max_bits) {                                                  // one atomic per warp and sample it touches (warps rarely straddle two)
        const long long smp = r < rows ? r / rows_per_sample : -1;
        const long long smp0 = __shfl_sync(0xffffffffu, smp, 0);
        if (__all_sync(0xffffffffu, smp == smp0)) {
            float m = nrm;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((threadIdx.x & 31) == 0 && smp0 >= 0) atomicMax(max_bits + smp0, __float_as_uint(m));
        } else if (smp >= 0) {
            atomicMax(max_bits + smp, __float_as_uint(nrm));
        }
    }
}

}  // namespace glare

using namespace glare;

// ref_out[r] = scale * max_{j < n_sub} s_sub[r * lds + j] + offset: the per-row softmax reference of glare_attn_scores_exp_tc (key_norm_max
// null) from the scores against a sampled subset of the keys
GLARE_API int glare_attn_row_ref(const float* s_sub, long long rows, long long lds, int n_sub, float scale, float offset, float* ref_out,
                                 cudaStream_t stream) {
    if (rows < 0 || n_sub <= 0 || lds < n_sub || !(scale > 0.f)) return GLARE_ERR_BAD_ARG;
    if (rows == 0) return GLARE_OK;
    if (!s_sub || !ref_out || (rows + 7) / 8 > 0x7fffffffLL) return GLARE_ERR_BAD_ARG;
    attn_row_ref_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(s_sub, rows, lds, n_sub, scale, offset, ref_out);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// norm_out[r] = sqrt(sum_{b < n_blocks} part[b * part_stride + r]) (may be null); max_bits[r / rows_per_sample] = max over the sample (may be null)
GLARE_API int glare_attn_row_norm_finish(const float* part, long long part_stride, int n_blocks, long long rows, long long rows_per_sample,
                                         float* norm_out, unsigned* max_bits, cudaStream_t stream) {
    if (rows < 0 || n_blocks <= 0 || part_stride < rows || rows_per_sample <= 0 || (!norm_out && !max_bits)) return GLARE_ERR_BAD_ARG;
    if (rows == 0) return GLARE_OK;
    if (!part) return GLARE_ERR_BAD_ARG;
    attn_row_norm_finish_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, stream>>>(part, part_stride, n_blocks, rows, rows_per_sample, norm_out, max_bits);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// x [rows][C] fp32 (C % 4 == 0): norm_out[row] = |x_row| (may be null); max_bits[row / rows_per_sample] = max over the sample's rows
// (may be null; the caller zeroes it first)
GLARE_API int glare_attn_row_norm(const float* x, long long rows, int C, long long rows_per_sample, float* norm_out, unsigned* max_bits,
                                  cudaStream_t stream) {
    if (rows < 0 || C <= 0 || (C & 3) || rows_per_sample <= 0 || (!norm_out && !max_bits)) return GLARE_ERR_BAD_ARG;
    if (rows == 0) return GLARE_OK;
    if (!x || (rows + 7) / 8 > 0x7fffffffLL) return GLARE_ERR_BAD_ARG;
    attn_row_norm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(x, rows, C, rows_per_sample, norm_out, max_bits);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// row_scale[r] = 1 / sum_{b < n_blocks} part[b * part_stride + r]; *flag |= 1 when a sum is < 1e-24, infinite or NaN
GLARE_API int glare_attn_row_sum_finish(const float* part, long long part_stride, int n_blocks, long long rows, float* row_scale, int* flag,
                                        cudaStream_t stream) {
    if (rows < 0 || n_blocks <= 0 || part_stride < rows) return GLARE_ERR_BAD_ARG;
    if (rows == 0) return GLARE_OK;
    if (!part || !row_scale || !flag) return GLARE_ERR_BAD_ARG;
    attn_row_sum_finish_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, stream>>>(part, part_stride, n_blocks, rows, row_scale, flag);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// S [rows][lds] fp32 logits (first n_keys columns valid) -> P [rows][ldp] operand (out_mode 0 bf16, 1 fp32, 2 tf32 hi+lo),
// P = softmax(scale * S) over the keys, zero in columns [n_keys, n_pad).
GLARE_API int glare_attn_softmax_rows(int out_mode, const float* S, long long rows, long long lds, int n_keys, int n_pad, float scale,
                                      void* out_hi, void* out_lo, long long ldp, cudaStream_t stream) {
    if (out_mode < 0 || out_mode > 4 || rows < 0 || n_keys <= 0 || n_pad < n_keys || (n_pad & 3) || (lds & 3) || (ldp & 3) || lds < n_keys ||
        ldp < n_pad || scale <= 0.f || (out_mode >= 3 && ((ldp & 31) || (n_pad & 31))))
        return GLARE_ERR_BAD_ARG;
    if (rows == 0) return GLARE_OK;
    if (!S || !out_hi || ((out_mode == 2 || out_mode == 3) && !out_lo) || rows > 0x7fffffffLL) return GLARE_ERR_BAD_ARG;
    float* lo = reinterpret_cast<float*>(out_lo);
    if (n_pad <= 256 * 64) {                         // whole row in registers: one pass over S (N = 16 275 at 600x400)
        if (out_mode == 0) attn_softmax_rows_reg_kernel<0><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
        else if (out_mode == 1) attn_softmax_rows_reg_kernel<1><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
        else if (out_mode == 2) attn_softmax_rows_reg_kernel<2><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
        else if (out_mode == 3) attn_softmax_rows_reg_kernel<3><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
        else attn_softmax_rows_reg_kernel<4><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
        GLARE_CHECK_LAUNCH();
        return GLARE_OK;
    }
    if (out_mode == 0) attn_softmax_rows_kernel<0><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
    else if (out_mode == 1) attn_softmax_rows_kernel<1><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
    else if (out_mode == 2) attn_softmax_rows_kernel<2><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
    else if (out_mode == 3) attn_softmax_rows_kernel<3><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
    else attn_softmax_rows_kernel<4><<<(unsigned)rows, 256, 0, stream>>>(S, lds, n_keys, n_pad, scale, out_hi, lo, ldp);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// v NHWC [B][N][C] fp32 -> V^T [B][C][Np] operand(s) (zero for keys >= N)
GLARE_API int glare_attn_transpose_v(int out_mode, const float* v, int B, int N, int C, int Np, void* out_hi, void* out_lo,
                                     cudaStream_t stream) {
    if (out_mode < 0 || out_mode > 4 || B < 0 || N <= 0 || C <= 0 || Np < N || (out_mode >= 3 && (Np & 31))) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!v || !out_hi || ((out_mode == 2 || out_mode == 3) && !out_lo) || B > 65535) return GLARE_ERR_BAD_ARG;
    dim3 grid((Np + 31) / 32, (C + 31) / 32, B);
    float* lo = reinterpret_cast<float*>(out_lo);
    if (out_mode == 0) attn_transpose_v_kernel<0><<<grid, 256, 0, stream>>>(v, N, C, Np, out_hi, lo);
    else if (out_mode == 1) attn_transpose_v_kernel<1><<<grid, 256, 0, stream>>>(v, N, C, Np, out_hi, lo);
    else if (out_mode == 2) attn_transpose_v_kernel<2><<<grid, 256, 0, stream>>>(v, N, C, Np, out_hi, lo);
    else if (out_mode == 3) attn_transpose_v_kernel<3><<<grid, 256, 0, stream>>>(v, N, C, Np, out_hi, lo);
    else attn_transpose_v_kernel<4><<<grid, 256, 0, stream>>>(v, N, C, Np, out_hi, lo);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
