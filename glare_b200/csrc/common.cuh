// Shared device/host helpers for the glare_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

// exported C-ABI symbol (the library is built with -fvisibility=hidden)
#define GLARE_API extern "C" __attribute__((visibility("default")))

#define GLARE_OK 0
#define GLARE_ERR_BAD_ARG (-1)
#define GLARE_ERR_UNSUPPORTED (-2)

// Every C-ABI entry point returns 0 or a cudaError_t (positive) / GLARE_ERR_* (negative).
#define GLARE_CHECK_LAUNCH()                                  \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return (int)e__;              \
    } while (0)

#define GLARE_CUDA(call)                                      \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return (int)e__;              \
    } while (0)

namespace glare {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + bulk async copy (TMA 1-D) ------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// fp32 -> nearest tf32 value (round to nearest, low 13 mantissa bits zero).  hi of the 3xTF32 split; lo = x - hi is exact
// in fp32 and |lo| <= 2^-12 |x|, so the dropped lo*lo term is below fp32 resolution.
__device__ __forceinline__ float tf32_round(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u & 0xFFFFE000u);
}

// Mode 3 ("tf32 + 2 x bf16"): D += tf32(A_hi) tf32(B_hi) + bf16(A_lo) bf16(B) + bf16(A) bf16(B_lo).  The two small cross terms
// tolerate bf16 (|lo| <= 2^-12 |x|, bf16 adds 2^-9 relative to them: ~1e-6 of the product), and bf16 MMAs run at twice the
// tf32 rate, so the fp32-grade product costs 2 instead of 3 tf32-equivalents.  Cross-term operands live in ONE bf16 "x"
// tensor, interleaved per 32-element K chunk: [32 x bf16(x) | 32 x bf16(x - hi)] = one 128-byte swizzle row per chunk.
// For flat element index e (row lengths are multiples of 32): bf16(x) at (e/32)*64 + e%32, bf16(lo) 32 further.
__device__ __forceinline__ void store_x4(__nv_bfloat16* xbase, long long e, float v0, float v1, float v2, float v3, float h0, float h1,
                                         float h2, float h3) {
    const long long xi = (e >> 5) * 64 + (e & 31);
    __nv_bfloat162 a = __floats2bfloat162_rn(v0, v1), b = __floats2bfloat162_rn(v2, v3);
    __nv_bfloat162 c = __floats2bfloat162_rn(v0 - h0, v1 - h1), d = __floats2bfloat162_rn(v2 - h2, v3 - h3);
    uint2 u, w;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    w.x = *reinterpret_cast<uint32_t*>(&c); w.y = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint2*>(xbase + xi) = u;
    *reinterpret_cast<uint2*>(xbase + xi + 32) = w;
}

// Mode 4 ("bf16x3"): x = a1 + a2 + O(2^-17 x) with a1 = bf16(x), a2 = bf16(x - a1); D += A1 B1 + A1 B2 + A2 B1 (three bf16 passes =
// 1.5 tf32-equivalents, 16-bit-significand products).  ONE operand tensor per side, interleaved like the mode-3 x tensor:
// [32 x a1 | 32 x a2] per 32-element K chunk, i.e. 4 bytes per element instead of 8.
__device__ __forceinline__ void split_b3(float v, __nv_bfloat16& a1, __nv_bfloat16& a2) {
    a1 = __float2bfloat16_rn(v);
    a2 = __float2bfloat16_rn(v - __bfloat162float(a1));
}
__device__ __forceinline__ void store_b3_4(__nv_bfloat16* xbase, long long e, float v0, float v1, float v2, float v3) {
    // same values as split_b3 per element; pairs are rounded and packed by one F2FP each
    const long long xi = (e >> 5) * 64 + (e & 31);
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v0, v1), h23 = __floats2bfloat162_rn(v2, v3);
    uint2 u, w;
    u.x = *reinterpret_cast<const uint32_t*>(&h01);
    u.y = *reinterpret_cast<const uint32_t*>(&h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v0 - __uint_as_float(u.x << 16), v1 - __uint_as_float(u.x & 0xffff0000u));
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(v2 - __uint_as_float(u.y << 16), v3 - __uint_as_float(u.y & 0xffff0000u));
    w.x = *reinterpret_cast<const uint32_t*>(&l01);
    w.y = *reinterpret_cast<const uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(xbase + xi) = u;
    *reinterpret_cast<uint2*>(xbase + xi + 32) = w;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace glare
