// Minimal CPU execution shim for plain CUDA C kernels (no tensor cores, TMA or inline PTX).  TEST INFRASTRUCTURE.
//
// A kernel source that sticks to threadIdx / blockIdx / __shared__ / __syncthreads / warp shuffles / atomicAdd / __ldg compiles with
// g++ -x c++ -DGLARE_CUDA_EMU against this header and runs on the host: blocks one after another, the threads of a block as real OS
// threads (so __syncthreads and shuffles have their CUDA meaning: std::barrier per block, an exchange buffer per warp), __shared__
// variables as function-local statics (one block is resident at a time).  It checks a kernel's indexing and arithmetic without a GPU;
// it says nothing about performance, memory coalescing or hardware-only behaviour.  Used for csrc/flow_bwd.cu
// (tests/test_flow_train_cpu.py), whose launches go through the FB_LAUNCH macro for this purpose.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define GLARE_API extern "C" __attribute__((visibility("default")))
#define GLARE_OK 0
#define GLARE_ERR_BAD_ARG (-1)
#define GLARE_ERR_UNSUPPORTED (-2)
#define GLARE_CHECK_LAUNCH() do { } while (0)
#define GLARE_CUDA(call) do { (void)(call); } while (0)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

typedef void* cudaStream_t;

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

static inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }

namespace glare_emu {
struct WarpState {
    uint32_t buf[32];
    std::unique_ptr<std::barrier<>> bar;
};
struct BlockState {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<WarpState> warps;
};
inline BlockState*& cur_block() { static BlockState* b = nullptr; return b; }
inline std::mutex& atomic_mutex() { static std::mutex m; return m; }
}  // namespace glare_emu

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

static inline void __syncthreads() { glare_emu::cur_block()->bar->arrive_and_wait(); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// every lane of the warp must call (full mask), as in the kernels this shim is used for
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    glare_emu::WarpState& w = glare_emu::cur_block()->warps[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    w.buf[lane] = __float_as_uint(v);
    w.bar->arrive_and_wait();
    const float r = __uint_as_float(w.buf[lane ^ lane_mask]);
    w.bar->arrive_and_wait();
    return r;
}
static inline float atomicAdd(float* p, float v) {
    std::lock_guard<std::mutex> g(glare_emu::atomic_mutex());
    const float old = *p;
    *p = old + v;
    return old;
}
static inline double atomicAdd(double* p, double v) {
    std::lock_guard<std::mutex> g(glare_emu::atomic_mutex());
    const double old = *p;
    *p = old + v;
    return old;
}

namespace glare_emu {
template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, A... args) {
    gridDim = grid;
    blockDim = block;
    const unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                BlockState st;
                st.bar = std::make_unique<std::barrier<>>(nthreads);
                st.warps.resize((nthreads + 31) / 32);
                for (unsigned wi = 0; wi < st.warps.size(); ++wi) {
                    const unsigned lanes = (wi + 1) * 32 <= nthreads ? 32 : nthreads - wi * 32;
                    st.warps[wi].bar = std::make_unique<std::barrier<>>(lanes);
                }
                cur_block() = &st;
                std::vector<std::thread> th;
                th.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t)
                    th.emplace_back([&, t]() {
                        threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        blockIdx = dim3(bx, by, bz);
                        kernel(args...);
                        st.bar->arrive_and_drop();               // an exited thread no longer takes part in __syncthreads
                    });
                for (auto& x : th) x.join();
            }
}
}  // namespace glare_emu
