// DCNv2Pack.forward tail on tensor cores: chunk/cat/sigmoid of the conv_offset output + modulated deformable 3x3 conv.
//
// Replaces reference deformableDecoder_arch.py:141-152 (o1,o2,mask = chunk(conv_offset(feat)); offset = cat(o1,o2);
// mask = sigmoid(mask); modulated_deform_conv(...)) and ops/dcn/src/deform_conv_cuda.cpp:490-569 +
// deform_conv_cuda_kernel.cu:467-497,571-633 for the configuration GLARE uses: 3x3, stride 1, pad 1, dilation 1,
// groups 1, deformable_groups dg (4).  The general fp32 operator stays in dcn.cu.
//
//   y[n,h,w,co] = bias[co] + sum_{g,t,c in g} W[co,c,t] * sigmoid(m[g,t]) * bilinear(x[n,:,:,c], (h,w) - 1 + t + off[g,t])
//
// Same persistent tcgen05 implicit-GEMM skeleton as conv_tc.cu (M = 8x16 pixel tile, N = BN output channels, TMEM
// double-buffered accumulator, TMA for the weight operand), but the A operand is PRODUCED in shared memory by eight
// sampler warps instead of loaded by TMA: per (tap, 128-byte channel chunk) each 8-lane group gathers the four
// bilinear corners of one pixel as whole 128-byte NHWC channel lines (L1-cached, coalesced), blends them with the
// corner weights x mask (geometry computed once per (pixel, group, tap) for all channels of the chunk), and stores the
// operand row in the SWIZZLE_128B pattern the UMMA descriptor expects (16-byte chunk j of row r lives at chunk
// j ^ (r & 7)); fence.proxy.async + mbarrier publish the stage to the MMA thread.  No `columns` buffer, no separate
// offset/mask tensors; the 3xTF32 hi/lo split is taken directly from the fp32 sample.
// Layouts: x NHWC fp32 [B,H,W,C]; om = raw conv_offset output NHWC fp32 [B,H,W,3*dg*9] (channels [0,18dg) offsets in the
// reference's (g, tap, {dh,dw}) order, [18dg,27dg) mask logits); weights packed [Cout][9][C] (glare_conv_pack_weight).
#include <cuda.h>
#include <stdlib.h>

#include "tc.cuh"

namespace glare {

constexpr int DT_FIXED_THREADS = 192;     // warp 0 TMA(weights), 1 MMA, 2-5 epilogue; then two groups of SW sampler warps (SW = 4 or 8)
constexpr int DT_A_BYTES = 128 * 128;
constexpr int DT_TH = 8, DT_TW = 16;        // pixel tile (compile-time: m / TW and m % TW sit in the sampler's inner loop)

struct DcnTcArgs {
    const float* x;
    const float* om;
    const float* bias;
    float* y;
    int B, H, W, C, Cout, dg, cpg, TH, TW, tiles_x, tiles_y, n_blocks, kchunks;   // kchunks = C / BKE channel chunks per tap
    int total_tiles;
    int cpg_shift;           // log2(cpg) when C / deformable_groups is a power of two, else -1
    int mask_prob;           // the mask channels of om already hold sigmoid(m) (operator-level entry: offset / mask tensors of the reference op)
    int stages;              // ring depth actually used (<= DcnCfg::STAGES): fewer stages leave more of the SM's 256 KB to the L1 cache,
                             // which is what serves the bilinear corner gathers (each 128-byte channel line is touched ~36 times per tile)
};

constexpr int DT_OM_MAX = 108;            // 27 * deformable_groups floats of conv_offset output per pixel staged in smem (dg <= 4)

template <int MODE, int BN, bool OMS = false>
struct DcnCfg {
    static constexpr bool XB = MODE == 3;                      // tf32 main term + two bf16 cross terms (common.cuh store_x4)
    static constexpr bool B3 = MODE == 4;                      // bf16x3: one interleaved bf16 tile per operand (common.cuh split_b3)
    static constexpr bool X3 = MODE == 2 || XB;
    static constexpr bool TF32 = MODE >= 1 && MODE <= 3;
    static constexpr int BKE = (TF32 || B3) ? 32 : 64;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = (DT_A_BYTES + B_BYTES) * (X3 ? 2 : 1);
    static constexpr int OM_BYTES = OMS ? 128 * DT_OM_MAX * 4 : 0;        // the tile's offsets + sigmoid(mask), loaded once per tile
    static constexpr int SMEM_BUDGET = 227 * 1024 - 2048 - OM_BYTES;
    static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : (STAGES_RAW < 1 ? 1 : STAGES_RAW);
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_DYN = STAGES * STAGE_BYTES + OM_BYTES + 1024;
};

__device__ __forceinline__ float tf32_hi_d(float x) { return tf32_round(x); }

template <int MODE, int BN, bool OMS, int SW>
__global__ void __launch_bounds__(DT_FIXED_THREADS + 64 * SW, 1)
dcn_tc_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo, const DcnTcArgs a) {
    using Cfg = DcnCfg<MODE, BN, OMS>;
    const int STAGES = a.stages;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* const smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmB);
        if (Cfg::X3) tma_prefetch_desc(&tmBlo);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1 + 32 * SW);      // weight TMA (expect_tx) + the sampler threads of one group
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar[0], 1);
        mbar_init(&tmem_full_bar[1], 1);
        mbar_init(&tmem_empty_bar[0], 128);
        mbar_init(&tmem_empty_bar[1], 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&s_tmem_base, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    const int k_iters = 9 * a.kchunks;
    const int tiles_xy = a.tiles_y * a.tiles_x;

    if (warp == 0) {
        // ===================== weight TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                const int nb = tile % a.n_blocks;
                // K order = channel chunk major, tap minor (the accumulation order is free): the nine taps of one 32-channel chunk read the same
                // ~10 x 18 pixel window of 128-byte channel lines back to back, so the window (23 KB) stays in the small L1 that the operand
                // rings leave; tap-major order put the other chunks' windows (4 - 8 x 23 KB) between two uses of a line
                for (int kc = 0; kc < a.kchunks; ++kc)
                    for (int tap = 0; tap < 9; ++tap, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait_backoff(&empty_bar[s], ph ^ 1, 256);
                        uint8_t* st = smem_al + (size_t)s * Cfg::STAGE_BYTES;
                        const int kcoord = tap * a.C + kc * Cfg::BKE;
                        mbar_arrive_expect_tx(&full_bar[s], Cfg::B_BYTES * (Cfg::X3 ? 2 : 1));
                        tma_load_3d(st + DT_A_BYTES, &tmB, &full_bar[s], (Cfg::B3 ? 2 : 1) * kcoord, nb * BN, 0);
                        if (Cfg::X3)
                            tma_load_3d(st + 2 * DT_A_BYTES + Cfg::B_BYTES, &tmBlo, &full_bar[s], (Cfg::XB ? 2 : 1) * kcoord, nb * BN, 0);
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(Cfg::TF32 ? 2 : 1, 128, BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
                mbar_wait_backoff(&tmem_empty_bar[as], aph ^ 1, 128);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int ki = 0; ki < k_iters; ++ki, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    // (in THIS kernel the MMA thread waits for the samplers nearly all the time: back off here too; <= 128 ns per ~1 us stage)
                    mbar_wait_backoff(&full_bar[s], ph, 128);
                    tc_fence_after();
                    const uint32_t sa = smem_base + (uint32_t)s * Cfg::STAGE_BYTES;
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + DT_A_BYTES);
                    const uint64_t dal = umma_desc_sw128(sa + DT_A_BYTES + Cfg::B_BYTES);
                    const uint64_t dbl = umma_desc_sw128(sa + 2 * DT_A_BYTES + Cfg::B_BYTES);
                    if (Cfg::B3) {
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            const uint64_t adv = (uint64_t)(jj * 2);
                            umma_ss<false>(d_tmem, da + adv, db + adv, idesc, (ki | jj) != 0 ? 1u : 0u);   // A1 B1
                            umma_ss<false>(d_tmem, da + adv, db + 4 + adv, idesc, 1u);                     // A1 B2
                            umma_ss<false>(d_tmem, da + 4 + adv, db + adv, idesc, 1u);                     // A2 B1
                        }
                    }
#pragma unroll
                    for (int j = 0; j < (Cfg::B3 ? 0 : 4); ++j) {
                        const uint64_t adv = (uint64_t)(j * 2);
                        umma_ss<Cfg::TF32>(d_tmem, da + adv, db + adv, idesc, (ki | j) != 0 ? 1u : 0u);
                        if (Cfg::X3 && !Cfg::XB) {
                            umma_ss<true>(d_tmem, dal + adv, db + adv, idesc, 1u);
                            umma_ss<true>(d_tmem, da + adv, dbl + adv, idesc, 1u);
                        }
                    }
                    if (Cfg::XB) {
                        constexpr uint32_t idesc_bf = umma_idesc(1, 128, BN);
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            const uint64_t adv = (uint64_t)(jj * 2);
                            umma_ss<false>(d_tmem, dal + 4 + adv, dbl + adv, idesc_bf, 1u);      // A_lo * B
                            umma_ss<false>(d_tmem, dal + adv, dbl + 4 + adv, idesc_bf, 1u);      // A * B_lo
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full_bar[as]);
            }
        }
    } else if (warp < 6) {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int py = m / DT_TW, px = m - py * DT_TW;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
            const int nb = tile % a.n_blocks;
            int r = tile / a.n_blocks;
            const int n = r / tiles_xy;
            r -= n * tiles_xy;
            const int ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
            const int gy = ty * DT_TH + py, gx = tx * DT_TW + px;
            const bool valid = gy < a.H && gx < a.W;
            const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait_backoff(&tmem_full_bar[as], aph, 1024);
            tc_fence_after();
            float* yrow = a.y + (((long long)n * a.H + gy) * a.W + gx) * a.Cout;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c0, v);
                tmem_ld_wait();
                const int co = nb * BN + c0;
                if (valid && co < a.Cout) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (co + j + 4 <= a.Cout) {
                            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                   __uint_as_float(v[j + 3]));
                            if (a.bias) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + co + j));
                                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                            }
                            *reinterpret_cast<float4*>(yrow + co + j) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[as]);
        }
    } else {
        // ===================== samplers: produce the A operand =====================
        // 8 warps in two groups; group gsel fills the stages with (k-iteration & 1) == gsel, so two stages are being
        // sampled while a third is being multiplied.  Within a stage a warp owns 32 tile rows: 8 lanes x 16 bytes per
        // operand row, 4 rows per step, 8 steps issued as two batches of 16 independent 128-bit corner loads.
        constexpr int RW = 128 / SW;                   // tile rows per sampler warp: 32 (two 16-row passes) or 16
        constexpr int QB = SW == 8 ? 2 : 4;            // pixel rows per lane whose corner loads are in flight together (x 4 corners)
        const int sw = (warp - 6) % SW;                // rows [RW sw, RW sw + RW) of the tile
        const int gsel = (warp - 6) / SW;
        const int sub = lane >> 3, j = lane & 7;
        constexpr int CPL = (Cfg::TF32 || Cfg::B3) ? 4 : 8;   // channels per lane per stage
        constexpr int V4 = CPL / 4;
        const int om_c = 27 * a.dg;
        float* const s_om = reinterpret_cast<float*>(smem_al + (size_t)STAGES * Cfg::STAGE_BYTES);   // [128][om_c] when OMS
        // integer divisions by run-time values cost ~25 instructions each and the sampler is instruction-bound (ncu r31: 250 instructions per
        // lane and (pixel, tap, chunk)): ring slot / phase are carried as counters, the deformable group comes from a shift when
        // C / deformable_groups is a power of two (GLARE: 32 / 64), the tile geometry is compile-time
        uint32_t it = 0;
        int ring_s = 0;
        uint32_t ring_ph = 0;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            int r = tile / a.n_blocks;
            const int n = r / tiles_xy;
            r -= n * tiles_xy;
            const int ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
            const float* xn = a.x + (long long)n * a.H * a.W * a.C;
            if (OMS) {
                // stage the tile's conv_offset output once (offsets raw, mask through the sigmoid): the geometry of every
                // (pixel, group, tap) then starts from shared memory instead of a dependent global load per use
                named_bar_sync(3, 64 * SW);                    // all sampler warps are done with the previous tile's values
                const int st_id = threadIdx.x - DT_FIXED_THREADS;
                // a tile row's 16 pixels x om_c values are one contiguous run in global memory AND in s_om ([m][om_c], m = 16 py + px): copy
                // run by run with the channel index carried as a counter (no division by the run-time om_c; ncu r57: 12 % of the kernel's
                // instructions sat in this loop)
                const int run = DT_TW * om_c, step_ch = (64 * SW) % om_c, mask_ch = 18 * a.dg;
                const int px_valid = a.W - tx * DT_TW < DT_TW ? a.W - tx * DT_TW : DT_TW;
                for (int py = 0; py < DT_TH; ++py) {
                    const int gy = ty * DT_TH + py;
                    const float* src = a.om + (((long long)n * a.H + gy) * a.W + tx * DT_TW) * om_c;
                    float* dst = s_om + py * run;
                    const int n_valid = gy < a.H ? px_valid * om_c : 0;
                    int ch = st_id % om_c;
                    for (int i = st_id; i < run; i += 64 * SW) {
                        float v = 0.f;
                        if (i < n_valid) {
                            v = __ldg(src + i);
                            if (ch >= mask_ch && !a.mask_prob) v = 1.0f / (1.0f + expf(-v));
                        }
                        dst[i] = v;
                        ch += step_ch;
                        if (ch >= om_c) ch -= om_c;
                    }
                }
                named_bar_sync(3, 64 * SW);
            }
            for (int kc = 0; kc < a.kchunks; ++kc) {
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    const int ti = tap / 3, tj = tap - ti * 3;
                    const int s = ring_s;                                  // == it % STAGES, (it / STAGES) & 1
                    const uint32_t ph = ring_ph;
                    if (++ring_s == STAGES) { ring_s = 0; ring_ph ^= 1u; }
                    if ((int)(it & 1) != gsel) continue;
                    const int cbase = kc * Cfg::BKE + j * CPL;             // first channel of this lane
                    const int g = a.cpg_shift >= 0 ? (cbase >> a.cpg_shift) : cbase / a.cpg;   // its deformable group (cpg % CPL == 0)
                    // geometry of one (pixel, group, tap): the four bilinear corner weights (x mask) and corner pixel offsets
                    auto geometry = [&](const int m, float (&gw)[4], int (&go)[4]) {
                        const int py = m / DT_TW, px = m - py * DT_TW;
                        const int gy = ty * DT_TH + py, gx = tx * DT_TW + px;
#pragma unroll
                        for (int k = 0; k < 4; ++k) { gw[k] = 0.f; go[k] = 0; }
                        if (gy < a.H && gx < a.W) {
                            float oh, ow, mk;
                            if (OMS) {
                                const float* omp = s_om + m * om_c;
                                oh = omp[g * 18 + 2 * tap]; ow = omp[g * 18 + 2 * tap + 1]; mk = omp[18 * a.dg + g * 9 + tap];
                            } else {
                                const float* omp = a.om + (((long long)n * a.H + gy) * a.W + gx) * om_c;
                                oh = __ldg(omp + g * 18 + 2 * tap); ow = __ldg(omp + g * 18 + 2 * tap + 1);
                                mk = __ldg(omp + 18 * a.dg + g * 9 + tap);
                                if (!a.mask_prob) mk = 1.0f / (1.0f + expf(-mk));
                            }
                            const float h_im = (float)(gy - 1 + ti) + oh, w_im = (float)(gx - 1 + tj) + ow;   // .cu:607-612
                            if (h_im > -1.f && w_im > -1.f && h_im < (float)a.H && w_im < (float)a.W) {            // .cu:618
                                const int hl = (int)floorf(h_im), wl = (int)floorf(w_im);
                                const int hh = hl + 1, wh = wl + 1;
                                const float lh = h_im - hl, lw = w_im - wl, uh = 1.f - lh, uw = 1.f - lw;
                                if (hl >= 0 && wl >= 0) { gw[0] = uh * uw * mk; go[0] = hl * a.W + wl; }
                                if (hl >= 0 && wh <= a.W - 1) { gw[1] = uh * lw * mk; go[1] = hl * a.W + wh; }
                                if (hh <= a.H - 1 && wl >= 0) { gw[2] = lh * uw * mk; go[2] = hh * a.W + wl; }
                                if (hh <= a.H - 1 && wh <= a.W - 1) { gw[3] = lh * lw * mk; go[3] = hh * a.W + wh; }
                            }
                        }
                    };
                    // The 8 lanes of a `sub` group work on the same pixels (8 x 4 channels = the 32-channel chunk).  When the chunk lies in
                    // one deformable group (C / deformable_groups a multiple of the chunk: GLARE's 32 / 64) they would all compute the same
                    // geometry: lane j computes it for ONE of the lane's RPL pixel rows of this stage and the others fetch it by shuffle
                    // (8 shuffles instead of ~100 instructions per row; the sampler is instruction-issue bound).
                    constexpr int RPL = RW / 4;                            // pixel rows per lane per stage: 4 (SW = 8) or 8 (SW = 4)
                    const bool share = (a.cpg % Cfg::BKE) == 0;              // the chunk's BKE channels (32, or 64 for bf16 operands) lie in one group
                    float myw[4];
                    int myo[4];
                    if (share) {
                        const int ridx = j & (RPL - 1), hq_m = ridx / QB, q4_m = ridx - hq_m * QB;
                        geometry(sw * RW + (hq_m / (4 / QB)) * 16 + ((hq_m % (4 / QB)) * QB + q4_m) + 4 * sub, myw, myo);
                    }
                    mbar_wait_bounded(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem_al + (size_t)s * Cfg::STAGE_BYTES;
#pragma unroll
                    for (int hq = 0; hq < (RW / 16) * (4 / QB); ++hq) {
                        // 16-row pass `half`, pixel rows q4 = qb .. qb + QB - 1 of it; row = 16 half + q4 + 4 sub, so the four rows one warp
                        // instruction stores differ in bit 2 of the row index pairwise (two shared-memory wavefronts instead of four)
                        const int half = hq / (4 / QB), qb = (hq % (4 / QB)) * QB;
                        float cw[QB][4];
                        int co[QB][4];
#pragma unroll
                        for (int q4 = 0; q4 < QB; ++q4) {
                            if (share) {
                                const int src = (lane & 24) + hq * QB + q4;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    cw[q4][k] = __shfl_sync(0xffffffffu, myw[k], src);
                                    co[q4][k] = __shfl_sync(0xffffffffu, myo[k], src);
                                }
                            } else {
                                geometry(sw * RW + half * 16 + (qb + q4) + 4 * sub, cw[q4], co[q4]);
                            }
                        }
                        // corners with zero weight read pixel 0 of the image (valid memory) and contribute nothing
                        float4 xv[QB][4][V4];
#pragma unroll
                        for (int q4 = 0; q4 < QB; ++q4)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float4* p = reinterpret_cast<const float4*>(xn + (co[q4][k] * a.C + cbase));      // < 2^31 per image (host check)
#pragma unroll
                                for (int v4 = 0; v4 < V4; ++v4) xv[q4][k][v4] = __ldg(p + v4);
                            }
#pragma unroll
                        for (int q4 = 0; q4 < QB; ++q4) {
                            const int m = sw * RW + half * 16 + (qb + q4) + 4 * sub;
                            float acc[CPL];
#pragma unroll
                            for (int v4 = 0; v4 < V4; ++v4) {
                                acc[4 * v4 + 0] = cw[q4][0] * xv[q4][0][v4].x;
                                acc[4 * v4 + 1] = cw[q4][0] * xv[q4][0][v4].y;
                                acc[4 * v4 + 2] = cw[q4][0] * xv[q4][0][v4].z;
                                acc[4 * v4 + 3] = cw[q4][0] * xv[q4][0][v4].w;
#pragma unroll
                                for (int k = 1; k < 4; ++k) {
                                    acc[4 * v4 + 0] = fmaf(cw[q4][k], xv[q4][k][v4].x, acc[4 * v4 + 0]);
                                    acc[4 * v4 + 1] = fmaf(cw[q4][k], xv[q4][k][v4].y, acc[4 * v4 + 1]);
                                    acc[4 * v4 + 2] = fmaf(cw[q4][k], xv[q4][k][v4].z, acc[4 * v4 + 2]);
                                    acc[4 * v4 + 3] = fmaf(cw[q4][k], xv[q4][k][v4].w, acc[4 * v4 + 3]);
                                }
                            }
                            const uint32_t off = (uint32_t)m * 128u + (uint32_t)((j ^ (m & 7)) << 4);     // SWIZZLE_128B
                            if (Cfg::B3) {
                                // single x tile: bytes [0,64) = a1 of the 32 channels, [64,128) = a2; this lane owns 8 bytes at 8j in each half
                                const __nv_bfloat162 h01 = __floats2bfloat162_rn(acc[0], acc[1]), h23 = __floats2bfloat162_rn(acc[2], acc[3]);
                                uint2 u, w2;
                                u.x = *reinterpret_cast<const uint32_t*>(&h01);
                                u.y = *reinterpret_cast<const uint32_t*>(&h23);
                                const __nv_bfloat162 l01 = __floats2bfloat162_rn(acc[0] - __uint_as_float(u.x << 16),
                                                                                 acc[1] - __uint_as_float(u.x & 0xffff0000u));
                                const __nv_bfloat162 l23 = __floats2bfloat162_rn(acc[2] - __uint_as_float(u.y << 16),
                                                                                 acc[3] - __uint_as_float(u.y & 0xffff0000u));
                                w2.x = *reinterpret_cast<const uint32_t*>(&l01);
                                w2.y = *reinterpret_cast<const uint32_t*>(&l23);
                                uint8_t* xt = st + (uint32_t)m * 128u + (uint32_t)((j & 1) * 8);
                                *reinterpret_cast<uint2*>(xt + (((j >> 1) ^ (m & 7)) << 4)) = u;
                                *reinterpret_cast<uint2*>(xt + (((4 + (j >> 1)) ^ (m & 7)) << 4)) = w2;
                            } else if (!Cfg::TF32) {
                                __nv_bfloat162 b0 = __floats2bfloat162_rn(acc[0], acc[1]), b1 = __floats2bfloat162_rn(acc[2], acc[3]);
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(acc[CPL - 4], acc[CPL - 3]),
                                               b3 = __floats2bfloat162_rn(acc[CPL - 2], acc[CPL - 1]);
                                uint4 u;
                                u.x = *reinterpret_cast<uint32_t*>(&b0); u.y = *reinterpret_cast<uint32_t*>(&b1);
                                u.z = *reinterpret_cast<uint32_t*>(&b2); u.w = *reinterpret_cast<uint32_t*>(&b3);
                                *reinterpret_cast<uint4*>(st + off) = u;
                            } else if (!Cfg::X3) {
                                *reinterpret_cast<float4*>(st + off) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                            } else if (!Cfg::XB) {
                                const float4 h = make_float4(tf32_hi_d(acc[0]), tf32_hi_d(acc[1]), tf32_hi_d(acc[2]), tf32_hi_d(acc[3]));
                                *reinterpret_cast<float4*>(st + off) = h;
                                *reinterpret_cast<float4*>(st + DT_A_BYTES + Cfg::B_BYTES + off) =
                                    make_float4(acc[0] - h.x, acc[1] - h.y, acc[2] - h.z, acc[3] - h.w);
                            } else {
                                // x tile row: bytes [0,64) bf16(sample) of the 32 channels, [64,128) bf16(sample - hi); this lane owns
                                // channels 4j..4j+3 = 8 bytes at 8j in each half; swizzle acts on 16-byte chunks
                                const float4 h = make_float4(tf32_hi_d(acc[0]), tf32_hi_d(acc[1]), tf32_hi_d(acc[2]), tf32_hi_d(acc[3]));
                                *reinterpret_cast<float4*>(st + off) = h;
                                uint8_t* xt = st + DT_A_BYTES + Cfg::B_BYTES + (uint32_t)m * 128u + (uint32_t)((j & 1) * 8);
                                __nv_bfloat162 a0 = __floats2bfloat162_rn(acc[0], acc[1]), a1 = __floats2bfloat162_rn(acc[2], acc[3]);
                                __nv_bfloat162 l0 = __floats2bfloat162_rn(acc[0] - h.x, acc[1] - h.y),
                                               l1 = __floats2bfloat162_rn(acc[2] - h.z, acc[3] - h.w);
                                uint2 u, w2;
                                u.x = *reinterpret_cast<uint32_t*>(&a0); u.y = *reinterpret_cast<uint32_t*>(&a1);
                                w2.x = *reinterpret_cast<uint32_t*>(&l0); w2.y = *reinterpret_cast<uint32_t*>(&l1);
                                *reinterpret_cast<uint2*>(xt + ((((j >> 1)) ^ (m & 7)) << 4)) = u;
                                *reinterpret_cast<uint2*>(xt + (((4 + (j >> 1)) ^ (m & 7)) << 4)) = w2;
                            }
                        }
                    }
                    fence_proxy_async();              // generic-proxy smem writes -> visible to the tensor-core (async) proxy
                    mbar_arrive(&full_bar[s]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFnD)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_w_map_d(CUtensorMap* m, const void* ptr, bool bf16, int Cout, int K, int BN) {
    static EncodeTiledFnD enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return GLARE_ERR_UNSUPPORTED;
        enc = reinterpret_cast<EncodeTiledFnD>(p);
    }
    const cuuint64_t es = bf16 ? 2 : 4;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Cout, 1};
    cuuint64_t strides[2] = {(cuuint64_t)K * es, (cuuint64_t)K * Cout * es};
    cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GLARE_OK : GLARE_ERR_BAD_ARG;
}

template <int MODE, int BN, bool OMS, int SW>
static int launch_dcn_tc_w(const CUtensorMap& tB, const CUtensorMap& tBl, const DcnTcArgs& a, int smem, cudaStream_t stream) {
    using Cfg = DcnCfg<MODE, BN, OMS>;
    GLARE_CUDA(cudaFuncSetAttribute(dcn_tc_kernel<MODE, BN, OMS, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_DYN));
    const int grid = a.total_tiles < kNumSMs ? a.total_tiles : kNumSMs;
    dcn_tc_kernel<MODE, BN, OMS, SW><<<grid, DT_FIXED_THREADS + 64 * SW, smem, stream>>>(tB, tBl, a);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

template <int MODE, int BN, bool OMS>
static int launch_dcn_tc_v(const CUtensorMap& tB, const CUtensorMap& tBl, DcnTcArgs a, cudaStream_t stream) {
    using Cfg = DcnCfg<MODE, BN, OMS>;
    // ring depth: GLARE_DCN_STAGES overrides (A/B switch for profiling; measured flat between 2 and 6 stages, profiles/r30)
    static const int want = getenv("GLARE_DCN_STAGES") ? atoi(getenv("GLARE_DCN_STAGES")) : Cfg::STAGES;
    a.stages = want < 2 ? 2 : want;
    if (a.stages > Cfg::STAGES) a.stages = Cfg::STAGES;
    const int smem = a.stages * Cfg::STAGE_BYTES + Cfg::OM_BYTES + 1024;
    // sampler warps per stage group: 8 (16 warps in flight, 2 x 4 corner loads per lane per batch) for the default mode, the profile of the
    // 4-warp version showed 3.5 warps per scheduler half of whose cycles waited on the corner loads (profiles/r31); GLARE_DCN_SW=4 selects it
    if constexpr (MODE == 4) {
        static const bool sw4 = getenv("GLARE_DCN_SW") != nullptr && atoi(getenv("GLARE_DCN_SW")) == 4;
        if (!sw4) return launch_dcn_tc_w<MODE, BN, OMS, 8>(tB, tBl, a, smem, stream);
    }
    return launch_dcn_tc_w<MODE, BN, OMS, 4>(tB, tBl, a, smem, stream);
}

template <int MODE, int BN>
static int launch_dcn_tc(const CUtensorMap& tB, const CUtensorMap& tBl, const DcnTcArgs& a, cudaStream_t stream) {
    static_assert(DcnCfg<MODE, BN, false>::STAGES >= 2, "pipeline needs at least two stages");
    // offsets/mask staged in shared memory when that still leaves a two-stage ring (not for the 96 KB stages of modes 2/3 at BN = 256)
    if constexpr (DcnCfg<MODE, BN, true>::STAGES_RAW >= 2) {
        static const bool no_oms = getenv("GLARE_DCN_NO_OM_SMEM") != nullptr;      // A/B switch for profiling only
        if (27 * a.dg <= DT_OM_MAX && !no_oms) return launch_dcn_tc_v<MODE, BN, true>(tB, tBl, a, stream);
    }
    return launch_dcn_tc_v<MODE, BN, false>(tB, tBl, a, stream);
}

}  // namespace glare

using namespace glare;

// x NHWC fp32 [B,H,W,C]; offmask = raw conv_offset output NHWC fp32 [B,H,W,27*dg]; w / w_lo packed by
// glare_conv_pack_weight(mode, weight[Cout,C,3,3]); y NHWC fp32 [B,H,W,Cout].  3x3, stride 1, pad 1, dilation 1, groups 1.
static int dcn_tc_launch(int mode, const float* x, const float* offmask, int mask_prob, const void* w, const void* w_lo, const float* bias_or_null,
                         float* y, int B, int H, int W, int C, int Cout, int deformable_groups, cudaStream_t stream);

GLARE_API int glare_dcnv2_pack_fwd_nhwc_tc(int mode, const float* x, const float* offmask, const void* w, const void* w_lo,
                                           const float* bias_or_null, float* y, int B, int H, int W, int C, int Cout,
                                           int deformable_groups, cudaStream_t stream) {
    return dcn_tc_launch(mode, x, offmask, 0, w, w_lo, bias_or_null, y, B, H, W, C, Cout, deformable_groups, stream);
}

// The reference OPERATOR's inputs on the same kernel (ModulatedDeformConvFunction.forward, ops/dcn/deform_conv.py:124-153, for 3x3 /
// stride 1 / pad 1 / dilation 1 / groups 1): offmask NHWC [B,H,W,27*dg] = the op's offset tensor (channels [0,18dg)) and its mask tensor
// (channels [18dg,27dg), ALREADY through the sigmoid) side by side.
GLARE_API int glare_dcnv2_fwd_nhwc_tc(int mode, const float* x, const float* offset_mask, const void* w, const void* w_lo,
                                      const float* bias_or_null, float* y, int B, int H, int W, int C, int Cout, int deformable_groups,
                                      cudaStream_t stream) {
    return dcn_tc_launch(mode, x, offset_mask, 1, w, w_lo, bias_or_null, y, B, H, W, C, Cout, deformable_groups, stream);
}

static int dcn_tc_launch(int mode, const float* x, const float* offmask, int mask_prob, const void* w, const void* w_lo, const float* bias_or_null,
                         float* y, int B, int H, int W, int C, int Cout, int deformable_groups, cudaStream_t stream) {
    if (mode < 0 || mode > 4 || B < 0 || H <= 0 || W <= 0 || C <= 0 || Cout <= 0 || deformable_groups <= 0) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !offmask || !w || !y || ((mode == 2 || mode == 3) && !w_lo)) return GLARE_ERR_BAD_ARG;
    const int bke = mode == 0 ? 64 : 32;
    if (C % deformable_groups != 0 || C % bke != 0 || (C / deformable_groups) % 8 != 0 || Cout % 4 != 0) return GLARE_ERR_UNSUPPORTED;
    if ((long long)H * W * C >= 0x7fffffffLL) return GLARE_ERR_UNSUPPORTED;        // the sampler addresses one image with 32-bit offsets
    DcnTcArgs a{};
    a.x = x; a.om = offmask; a.bias = bias_or_null; a.y = y; a.mask_prob = mask_prob;
    a.B = B; a.H = H; a.W = W; a.C = C; a.Cout = Cout; a.dg = deformable_groups; a.cpg = C / deformable_groups;
    a.TH = DT_TH; a.TW = DT_TW;
    a.tiles_x = (W + a.TW - 1) / a.TW; a.tiles_y = (H + a.TH - 1) / a.TH;
    a.kchunks = C / bke;
    a.cpg_shift = -1;
    for (int sft = 0; sft < 16; ++sft)
        if ((1 << sft) == a.cpg) a.cpg_shift = sft;
    int BN = Cout >= 256 ? 256 : (Cout > 64 ? 128 : 64);
    const long long m_tiles = (long long)B * a.tiles_y * a.tiles_x;
    while (BN > 64 && m_tiles * ((Cout + BN - 1) / BN) < kNumSMs) BN >>= 1;
    a.n_blocks = (Cout + BN - 1) / BN;
    const long long total = m_tiles * a.n_blocks;
    if (total > 0x7fffffff) return GLARE_ERR_UNSUPPORTED;
    a.total_tiles = (int)total;
    CUtensorMap tB, tBl;
    int rc;
    if ((rc = make_w_map_d(&tB, w, mode == 0 || mode == 4, Cout, (mode == 4 ? 2 : 1) * 9 * C, BN)) != GLARE_OK) return rc;
    tBl = tB;
    if (mode == 2 && (rc = make_w_map_d(&tBl, w_lo, false, Cout, 9 * C, BN)) != GLARE_OK) return rc;
    if (mode == 3 && (rc = make_w_map_d(&tBl, w_lo, true, Cout, 2 * 9 * C, BN)) != GLARE_OK) return rc;
#define GLARE_DCN_DISPATCH(M)                                                 \
    do {                                                                      \
        if (BN == 256) return launch_dcn_tc<M, 256>(tB, tBl, a, stream);      \
        if (BN == 128) return launch_dcn_tc<M, 128>(tB, tBl, a, stream);      \
        return launch_dcn_tc<M, 64>(tB, tBl, a, stream);                      \
    } while (0)
    if (mode == 0) GLARE_DCN_DISPATCH(0);
    if (mode == 1) GLARE_DCN_DISPATCH(1);
    if (mode == 2) GLARE_DCN_DISPATCH(2);
    if (mode == 3) GLARE_DCN_DISPATCH(3);
    GLARE_DCN_DISPATCH(4);
#undef GLARE_DCN_DISPATCH
}
