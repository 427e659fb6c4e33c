// Dense 3x3 / 1x1 convolution (stride 1, "same" padding) as an implicit GEMM on the 5th-gen tensor cores.
//
// Replaces the cuDNN calls behind the reference's nn.Conv2d on the hot path: ResnetBlock.conv1/conv2
// (encoder_decoder.py:88-115), Upsample.conv (:38-53), AttnBlock q/k/v/proj_out (:146-165), nin_shortcut,
// WarpBlock.offset / DCNv2Pack.conv_offset (deformableDecoder_arch.py:282, deform_conv.py:352-371) and the
// hoisted first layers of the flow coupling nets (flow.py:13-52).
//
//   y[n,h,w,co] = bias[co] + residual[n,h,w,co] + sum_{dy,dx,ci} x[n,h+dy-p,w+dx-p,ci] * W[co,dy,dx,ci]
//
// Layout: activations NHWC (channels innermost = GEMM K contiguous), weights [Cout][kh*kw][Cin] (K-major).
// Mapping (one CTA per SM, persistent over output tiles):
//   M = 128 output pixels (TH x TW patch of one image), N = BN output channels, K = taps x Cin.
//   warp 0   TMA producer: per (tap, 128-byte K chunk) one 4-D tiled load of the shifted activation patch --
//            out-of-image rows/columns are zero-filled by the TMA unit, which IS the conv zero padding -- and
//            one 2-D load of the weight slab, both SWIZZLE_128B, landing on an mbarrier (full[stage]).
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16 bf16 / 8 tf32) from shared-memory
//            descriptors into a TMEM accumulator; tcgen05.commit frees the stage (empty[stage]) and publishes
//            the finished accumulator (tmem_full[acc]).  Two accumulators (2 x BN TMEM columns) let the
//            epilogue of tile i overlap the MMAs of tile i+1.
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns -> registers, + bias (+ residual), 128-byte NHWC stores.
// Precision modes: 0 = bf16 operands, 1 = tf32 operands (fp32 storage), 2 = 3xTF32 (x = hi + lo split of both
// operands, D += A_hi B_hi + A_lo B_hi + A_hi B_lo) which restores ~fp32 accuracy for the fp32 configuration.
// Accumulation is fp32 in TMEM in every mode.
#include <cuda.h>
#include <stdlib.h>

#include "tc.cuh"

namespace glare {

constexpr int CT_THREADS = 192;
constexpr int CT_A_BYTES = 128 * 128;     // 128 pixels x 128-byte K chunk

struct ConvTcArgs {
    const float* bias;       // [Cout] or null
    const float* residual;   // NHWC [B,H,W,Cout] or null
    float* y;                // NHWC [B,H,W,Cout]
    int B, H, W, Cin, Cout, ksize, pad, TH, TW, tiles_x, tiles_y, n_blocks, kchunks;   // H, W: OUTPUT size
    int stride;              // 1, or 2 (Downsample: zero pad right/bottom, pad = 0; the TMA box walks the input with stride 2)
    int ntaps, tap_w, tap_dy0, tap_dx0;   // filter taps: tap t reads input offset (t / tap_w + tap_dy0, t % tap_w + tap_dx0)
    int oscale, oa, ob;      // output pixel of tile pixel (y, x) = (y * oscale + oa, x * oscale + ob): 2 for the sub-pixel phases of Upsample+conv
    int total_tiles;         // cluster work items: ceil(m_tiles / CL) * n_blocks
    int m_tiles;             // B * tiles_y * tiles_x
    int w_batched;           // weights differ per sample (3rd tensor-map coordinate = n): attention S = Q K^T, O = P V
    long long ldy;           // output row (pixel) stride in elements, >= Cout
    float* gn_part;          // optional [B][gn_slots][32][4]: per (pixel tile, epilogue warp) partial {S1, S2, shift, count} of the OUTPUT per
                             // GroupNorm group -- every entry written exactly once (no atomics), reduced in fp64 by conv_gn_finish_kernel
    int gn_cpg;              // channels per group (Cout / 32): 4, 8 or 16
    int gn_slots;            // tiles_y * tiles_x * 4
    int tma_store;           // epilogue: registers -> swizzled smem staging -> TMA tiled store (else per-thread float4 stores)
    // attention epilogues (mode 4 only; attn.cu documents the scheme).  Scores GEMM: the accumulator s becomes
    // p = exp(exp_scale * s - ref(row)), written as the bf16x3 operand of the P V GEMM; per-(row, output block) partial row sums go to
    // row_sum_part.  ref(row) = row_norm[row] when key_norm_max is null (the caller's reference: sampled row maximum + margin), else the
    // Cauchy-Schwarz form row_norm[row] * exp_scale * max|k| - exp_margin >= every logit of the row - margin.
    const float* row_norm;        // [rows] ref(row) or |q_row|, or null
    const unsigned* key_norm_max; // bits of max_j |k_j| (non-negative float), or null
    float* row_sum_part;          // [n_blocks][part_stride]
    long long part_stride;
    float exp_scale, exp_margin;
    const float* row_scale;       // P V GEMM: y[row] *= row_scale[row] (1 / row sum), or null
    int pack_out;                 // y is the bf16x3 operand of the next GEMM (same geometry as the fp32 output: 4 bytes per element), not fp32
    float* row_sq_part;           // [n_blocks][part_stride] partial sums of squares of the output rows (|q_i|, |k_j| of the attention reference), or null
};

template <int MODE, int BN, bool PAIR = false>
struct ConvCfg {
    static constexpr bool XB = MODE == 3;                      // tf32 main term + two bf16 cross terms
    static constexpr bool B3 = MODE == 4;                      // bf16x3: one interleaved bf16 tile per operand, three bf16 passes
    static constexpr bool X3 = MODE == 2 || XB;                // second operand pair (lo / interleaved bf16 x) per stage
    static constexpr bool TF32 = MODE >= 1 && MODE <= 3;
    static constexpr int BKE = (TF32 || B3) ? 32 : 64;         // K elements per 128-byte chunk (mode 4: 32 elements x 2 bf16 pieces)
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * 128;  // CTA pair: each CTA stages half of the weight rows
    static constexpr int STAGE_BYTES = (CT_A_BYTES + B_BYTES) * (X3 ? 2 : 1);
    static constexpr int STAGING_BYTES = 2 * 128 * 128;        // epilogue staging: two 128-pixel x 32-fp32 chunks
    static constexpr int SMEM_BUDGET = 227 * 1024 - 1024 /*align slack*/ - 1024 /*static*/ - STAGING_BYTES;
    static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = 2 * BN;                   // 128 / 256 / 512: powers of two >= 32
    static constexpr int SMEM_DYN = STAGES * STAGE_BYTES + STAGING_BYTES + 1024;
    // HALO variant (3x3 stride-1 convs, mode 4, CTA pair, BN <= 128): a stage holds the activation patch of one filter ROW -- an
    // (8 + 2) x 16 pixel box, 160 rows of 128 bytes -- and the weight slabs of its three taps.  Tap dx reads the same patch shifted by
    // dx rows (the descriptor's start address moves by dx * 128 bytes, 8-row groups stay 1280 bytes apart), so the patch is fetched
    // once instead of three times: 44 KB instead of 72 KB per three taps at BN = 128, where the L2 -> SM path (not the tensor pipe)
    // was the limit (profiles/r01o: 57 % tensor-pipe activity).
    static constexpr int HALO_TW = 8, HALO_TH = 16;
    static constexpr int HALO_A_BYTES = (HALO_TW + 2) * HALO_TH * 128;
    static constexpr int HALO_STAGE_BYTES = HALO_A_BYTES + 3 * B_BYTES;
    static constexpr int HALO_STAGES_RAW = SMEM_BUDGET / HALO_STAGE_BYTES;
    static constexpr int HALO_STAGES = HALO_STAGES_RAW > 8 ? 8 : HALO_STAGES_RAW;
    static constexpr int HALO_SMEM_DYN = HALO_STAGES * HALO_STAGE_BYTES + STAGING_BYTES + 1024;
    // RING2 variant (default for 3x3 stride-1 convs with the 256-wide N tile; GLARE_CONV_NO_RING2=1 disables): the same filter-row patch staging, where one
    // stage per filter row (patch + three 16 KB weight slabs = 68 KB) would leave only two stages.  Activation patches and weight slabs get their
    // own rings instead: R2_A_SLOTS patches (one per filter row and K chunk) and R2_B_SLOTS slabs (one per tap and K chunk).
    static constexpr int R2_A_SLOTS = 3;
    static constexpr int R2_B_RAW = (SMEM_BUDGET - R2_A_SLOTS * HALO_A_BYTES - 64) / B_BYTES;
    static constexpr int R2_B_SLOTS = R2_B_RAW > 8 ? 8 : R2_B_RAW;
    static constexpr int R2_RING_BYTES = R2_A_SLOTS * HALO_A_BYTES + R2_B_SLOTS * B_BYTES;
    static constexpr int R2_SMEM_DYN = R2_RING_BYTES + STAGING_BYTES + 64 + 1024;      // + 64: the patch ring's mbarriers live behind the staging tiles
};

template <bool TF32, bool PAIR>
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (PAIR) umma_ss_2sm<TF32>(d, da, db, idesc, acc);
    else umma_ss<TF32>(d, da, db, idesc, acc);
}

// CM = 1: thread-block cluster of two CTAs working on two pixel tiles of the SAME output-channel block; each CTA loads half of
// the weight tile and TMA-multicasts it into both shared memories, halving the L2 -> SM weight traffic (the fp32-grade modes are
// L2-bandwidth bound otherwise: 96 KB of operands per 1 M MACs).  A stage is released to the producers of both CTAs by a
// multicast tcgen05.commit.
// CM = 2: CTA pair (cta_group::2).  The two CTAs of a cluster form ONE 256-pixel x BN tile: each stages its own 128 pixel rows
// and HALF of the weight rows, all TMA loads signal the leader's barrier, the leader's MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both halves and accumulates each CTA's 128 rows into that CTA's TMEM.
// Weight staging traffic and weight shared-memory reads per CTA are halved -- the binding resources of the fp32-grade modes.
template <int MODE, int BN, int CM, bool HALO = false, bool RING2 = false>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
               const __grid_constant__ CUtensorMap tmY, const ConvTcArgs a) {
    constexpr int CL = CM == 0 ? 1 : 2;
    constexpr bool PAIR = CM == 2, MCAST = CM == 1;
    using Cfg = ConvCfg<MODE, BN, PAIR>;
    static_assert(!HALO || ((MODE == 4 || MODE == 0) && CM == 2 && BN <= 128), "halo staging: 128-byte bf16 rows (modes 0 / 4), CTA pair, BN <= 128");
    static_assert(!RING2 || ((MODE == 4 || MODE == 0) && CM == 2 && !HALO && Cfg::R2_B_SLOTS >= 4), "two-ring staging: modes 0 / 4, CTA pair");
    constexpr int STAGES = HALO ? Cfg::HALO_STAGES : Cfg::STAGES;
    constexpr int STAGE_BYTES = HALO ? Cfg::HALO_STAGE_BYTES : Cfg::STAGE_BYTES;
    constexpr int RING_BYTES = RING2 ? Cfg::R2_RING_BYTES : STAGES * STAGE_BYTES;        // operand rings; the epilogue staging tiles follow
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;      // SWIZZLE_128B atoms need 1024-byte alignment
    uint8_t* const smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmY);
        if (Cfg::X3) {
            tma_prefetch_desc(&tmAlo);
            tma_prefetch_desc(&tmBlo);
        }
#pragma unroll
        for (int s = 0; s < (RING2 ? Cfg::R2_B_SLOTS : STAGES); ++s) {
            mbar_init(&full_bar[s], PAIR ? 2 : 1);                   // pair: one arrival per CTA's producer, on the leader's barrier
            mbar_init(&empty_bar[s], MCAST ? 2 : 1);                 // multicast: one tcgen05.commit per CTA of the cluster
        }
        if (RING2) {
            // full_bar / empty_bar above serve the weight-slab ring; the patch ring's barriers sit behind the epilogue staging tiles
            uint64_t* a_bars = reinterpret_cast<uint64_t*>(smem_al + RING_BYTES + Cfg::STAGING_BYTES);
#pragma unroll
            for (int s = 0; s < Cfg::R2_A_SLOTS; ++s) {
                mbar_init(&a_bars[s], 2);                                // full: one arrival per CTA's producer
                mbar_init(&a_bars[4 + s], 1);                            // empty
            }
        }
        mbar_init(&tmem_full_bar[0], 1);
        mbar_init(&tmem_full_bar[1], 1);
        mbar_init(&tmem_empty_bar[0], PAIR ? 256 : 128);             // pair: the epilogue threads of both CTAs release the leader
        mbar_init(&tmem_empty_bar[1], PAIR ? 256 : 128);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_2sm(&s_tmem_base, Cfg::TMEM_COLS);
        else tmem_alloc(&s_tmem_base, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                  // peer barriers are initialised before any multicast targets them
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    const int k_iters = a.ntaps * a.kchunks;
    uint64_t* const a_full = reinterpret_cast<uint64_t*>(smem_al + RING_BYTES + Cfg::STAGING_BYTES);   // RING2 only
    uint64_t* const a_empty = a_full + 4;
    const int cl_rank = CL > 1 ? (int)cluster_cta_rank() : 0;
    const int cl_id = blockIdx.x / CL, n_cl = gridDim.x / CL;
    const int tiles_xy = a.tiles_y * a.tiles_x;
    // work item w -> (output-channel block, pixel tile of this CTA); a pixel tile past the end (odd tile count) has n == B:
    // its loads are out of bounds (zero fill), its stores are clipped
#define GLARE_DECODE_WORK(w)                                   \
    const int nb = (w) % a.n_blocks;                           \
    const int mt = ((w) / a.n_blocks) * CL + cl_rank;          \
    const int n = mt / tiles_xy;                               \
    const int r_ = mt - n * tiles_xy;                          \
    const int ty = r_ / a.tiles_x, tx = r_ - ty * a.tiles_x;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int w = cl_id; w < a.total_tiles; w += n_cl) {
                GLARE_DECODE_WORK(w)
                const int y0 = ty * a.TH * a.stride + a.tap_dy0, x0 = tx * a.TW * a.stride + a.tap_dx0;
                if (RING2) {
                    const int half = cl_rank * (BN / 2);
                    uint8_t* const ring_b = smem_al + Cfg::R2_A_SLOTS * Cfg::HALO_A_BYTES;
                    for (int dy = 0; dy < 3; ++dy) {
                        for (int kc = 0; kc < a.kchunks; ++kc, ++it) {           // it counts patches; slabs are 3 * it + dx
                            const int sa = it % Cfg::R2_A_SLOTS;
                            mbar_wait_backoff(&a_empty[sa], ((it / Cfg::R2_A_SLOTS) & 1) ^ 1, 256);
                            if (cl_rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2 * Cfg::HALO_A_BYTES);
                            else mbar_arrive_cluster(&a_full[sa], 0);
                            tma_load_4d_2sm(smem_al + (size_t)sa * Cfg::HALO_A_BYTES, &tmA, &a_full[sa], (Cfg::B3 ? 2 : 1) * kc * Cfg::BKE, x0, y0 + dy, n);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                const uint32_t ib = 3u * it + dx;
                                const int sb = ib % Cfg::R2_B_SLOTS;
                                mbar_wait_backoff(&empty_bar[sb], ((ib / Cfg::R2_B_SLOTS) & 1) ^ 1, 256);
                                if (cl_rank == 0) mbar_arrive_expect_tx(&full_bar[sb], 2 * Cfg::B_BYTES);
                                else mbar_arrive_cluster(&full_bar[sb], 0);
                                tma_load_3d_2sm(ring_b + (size_t)sb * Cfg::B_BYTES, &tmB, &full_bar[sb],
                                                (Cfg::B3 ? 2 : 1) * ((dy * 3 + dx) * a.Cin + kc * Cfg::BKE), nb * BN + half, 0);
                            }
                        }
                    }
                    continue;
                }
                if (HALO) {
                    const int wn = a.w_batched ? (n < a.B ? n : a.B - 1) : 0;
                    const int half = cl_rank * (BN / 2);
                    for (int dy = 0; dy < 3; ++dy) {
                        for (int kc = 0; kc < a.kchunks; ++kc, ++it) {
                            const int s = it % STAGES;
                            const uint32_t ph = (it / STAGES) & 1;
                            mbar_wait_backoff(&empty_bar[s], ph ^ 1, 256);
                            uint8_t* st = smem_al + (size_t)s * STAGE_BYTES;
                            if (cl_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                            else mbar_arrive_cluster(&full_bar[s], 0);
                            tma_load_4d_2sm(st, &tmA, &full_bar[s], (Cfg::B3 ? 2 : 1) * kc * Cfg::BKE, x0, y0 + dy, n);      // (TW + 2) x TH patch of this filter row
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
                                tma_load_3d_2sm(st + Cfg::HALO_A_BYTES + dx * Cfg::B_BYTES, &tmB, &full_bar[s],
                                                (Cfg::B3 ? 2 : 1) * ((dy * 3 + dx) * a.Cin + kc * Cfg::BKE), nb * BN + half, wn);
                        }
                    }
                    continue;
                }
                for (int tap = 0; tap < a.ntaps; ++tap) {
                    const int dy = tap / a.tap_w, dx = tap - dy * a.tap_w;
                    for (int kc = 0; kc < a.kchunks; ++kc, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait_backoff(&empty_bar[s], ph ^ 1, 256);
                        uint8_t* st = smem_al + (size_t)s * STAGE_BYTES;
                        const int wn = a.w_batched ? (n < a.B ? n : a.B - 1) : 0;   // a dummy tile still feeds the peer real weights
                        const int kw = tap * a.Cin + kc * Cfg::BKE;
                        const int km = Cfg::XB ? 2 : 1;                   // the bf16 x tensors hold 64 elements per 32-element K chunk
                        const int k1 = Cfg::B3 ? 2 : 1;                   // ... and in mode 4 the (only) operand tensor is such an x tensor
                        uint8_t* lo = st + CT_A_BYTES + Cfg::B_BYTES;
                        if (PAIR) {
                            // every byte of both CTAs lands on the leader's barrier; my weight slice = rows [rank * BN/2, +BN/2)
                            if (cl_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
                            else mbar_arrive_cluster(&full_bar[s], 0);
                            const int half = cl_rank * (BN / 2);
                            tma_load_4d_2sm(st, &tmA, &full_bar[s], k1 * kc * Cfg::BKE, x0 + dx, y0 + dy, n);
                            tma_load_3d_2sm(st + CT_A_BYTES, &tmB, &full_bar[s], k1 * kw, nb * BN + half, wn);
                            if (Cfg::X3) {
                                tma_load_4d_2sm(lo, &tmAlo, &full_bar[s], km * kc * Cfg::BKE, x0 + dx, y0 + dy, n);
                                tma_load_3d_2sm(lo + CT_A_BYTES, &tmBlo, &full_bar[s], km * kw, nb * BN + half, wn);
                            }
                            continue;
                        }
                        mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                        tma_load_4d(st, &tmA, &full_bar[s], k1 * kc * Cfg::BKE, x0 + dx, y0 + dy, n);
                        if (CL == 1) {
                            tma_load_3d(st + CT_A_BYTES, &tmB, &full_bar[s], k1 * kw, nb * BN, wn);
                        } else {                                          // my half of the weight rows, delivered to both CTAs
                            const int half = cl_rank * (BN / 2);
                            tma_load_3d_mc(st + CT_A_BYTES + half * 128, &tmB, &full_bar[s], k1 * kw, nb * BN + half, wn, (uint16_t)0x3);
                            if (Cfg::X3)
                                tma_load_3d_mc(lo + CT_A_BYTES + half * 128, &tmBlo, &full_bar[s], km * kw, nb * BN + half, wn, (uint16_t)0x3);
                        }
                        if (Cfg::X3) {
                            tma_load_4d(lo, &tmAlo, &full_bar[s], km * kc * Cfg::BKE, x0 + dx, y0 + dy, n);
                            if (CL == 1) tma_load_3d(lo + CT_A_BYTES, &tmBlo, &full_bar[s], km * kw, nb * BN, wn);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0 && (!PAIR || cl_rank == 0)) {                     // pair: only the leader CTA issues
            constexpr uint32_t idesc = umma_idesc(Cfg::TF32 ? 2 : 1, PAIR ? 256 : 128, BN);
            uint32_t it = 0, tcount = 0;
            for (int w = cl_id; w < a.total_tiles; w += n_cl, ++tcount) {
                const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
                mbar_wait_bounded(&tmem_empty_bar[as], aph ^ 1);        // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                if (RING2) {
                    const uint32_t ring_b = smem_base + Cfg::R2_A_SLOTS * Cfg::HALO_A_BYTES;
                    for (int ki = 0; ki < 3 * a.kchunks; ++ki, ++it) {
                        const int sa = it % Cfg::R2_A_SLOTS;
                        mbar_wait_bounded(&a_full[sa], (it / Cfg::R2_A_SLOTS) & 1);
                        const uint32_t pa = smem_base + (uint32_t)sa * Cfg::HALO_A_BYTES;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint32_t ib = 3u * it + dx;
                            const int sb = ib % Cfg::R2_B_SLOTS;
                            mbar_wait_bounded(&full_bar[sb], (ib / Cfg::R2_B_SLOTS) & 1);
                            tc_fence_after();
                            const uint64_t da = umma_desc_sw128_sbo(pa + dx * 128, (Cfg::HALO_TW + 2) * 128);
                            const uint64_t db = umma_desc_sw128(ring_b + (uint32_t)sb * Cfg::B_BYTES);
                            if (Cfg::B3) {
#pragma unroll
                                for (int jj = 0; jj < 2; ++jj) {
                                    const uint64_t adv = (uint64_t)(jj * 2);
                                    mma<false, PAIR>(d_tmem, da + adv, db + adv, idesc, (ki | dx | jj) != 0 ? 1u : 0u);
                                    mma<false, PAIR>(d_tmem, da + adv, db + 4 + adv, idesc, 1u);
                                    mma<false, PAIR>(d_tmem, da + 4 + adv, db + adv, idesc, 1u);
                                }
                            } else {                                      // mode 0: 64 bf16 per row, four K = 16 steps
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const uint64_t adv = (uint64_t)(j * 2);
                                    mma<false, PAIR>(d_tmem, da + adv, db + adv, idesc, (ki | dx | j) != 0 ? 1u : 0u);
                                }
                            }
                            umma_commit_2sm_mc(&empty_bar[sb], (uint16_t)0x3);       // this tap's weight slab is free in both CTAs
                        }
                        umma_commit_2sm_mc(&a_empty[sa], (uint16_t)0x3);             // ... and after the third tap, the patch
                    }
                    umma_commit_2sm_mc(&tmem_full_bar[as], (uint16_t)0x3);
                    continue;
                }
                if (HALO) {
                    for (int ki = 0; ki < 3 * a.kchunks; ++ki, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait_bounded(&full_bar[s], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_base + (uint32_t)s * STAGE_BYTES;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            // pixel (y, x + dx) of the patch is row y * (TW + 2) + x + dx: start dx rows in, 8-row groups one patch row apart
                            const uint64_t da = umma_desc_sw128_sbo(sa + dx * 128, (Cfg::HALO_TW + 2) * 128);
                            const uint64_t db = umma_desc_sw128(sa + Cfg::HALO_A_BYTES + dx * Cfg::B_BYTES);
                            if (Cfg::B3) {
#pragma unroll
                                for (int jj = 0; jj < 2; ++jj) {
                                    const uint64_t adv = (uint64_t)(jj * 2);
                                    mma<false, PAIR>(d_tmem, da + adv, db + adv, idesc, (ki | dx | jj) != 0 ? 1u : 0u);
                                    mma<false, PAIR>(d_tmem, da + adv, db + 4 + adv, idesc, 1u);
                                    mma<false, PAIR>(d_tmem, da + 4 + adv, db + adv, idesc, 1u);
                                }
                            } else {                                      // mode 0: 64 bf16 per row, four K = 16 steps
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const uint64_t adv = (uint64_t)(j * 2);
                                    mma<false, PAIR>(d_tmem, da + adv, db + adv, idesc, (ki | dx | j) != 0 ? 1u : 0u);
                                }
                            }
                        }
                        umma_commit_2sm_mc(&empty_bar[s], (uint16_t)0x3);
                    }
                    umma_commit_2sm_mc(&tmem_full_bar[as], (uint16_t)0x3);
                    continue;
                }
                for (int ki = 0; ki < k_iters; ++ki, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait_bounded(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_base + (uint32_t)s * STAGE_BYTES;
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + CT_A_BYTES);
                    const uint64_t dal = umma_desc_sw128(sa + CT_A_BYTES + Cfg::B_BYTES);
                    const uint64_t dbl = umma_desc_sw128(sa + 2 * CT_A_BYTES + Cfg::B_BYTES);
                    if (Cfg::B3) {
                        // rows = [32 x a1 | 32 x a2]: A1 B1 + A1 B2 + A2 B1, K = 16 per MMA, two K steps per half row
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            const uint64_t adv = (uint64_t)(jj * 2);
                            mma<false, PAIR>(d_tmem, da + adv, db + adv, idesc, (ki | jj) != 0 ? 1u : 0u);
                            mma<false, PAIR>(d_tmem, da + adv, db + 4 + adv, idesc, 1u);
                            mma<false, PAIR>(d_tmem, da + 4 + adv, db + adv, idesc, 1u);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < (Cfg::B3 ? 0 : 4); ++j) {       // 4 x 32-byte K steps inside the 128-byte swizzle row
                        const uint64_t adv = (uint64_t)(j * 2);        // +32 bytes in 16-byte units
                        mma<Cfg::TF32, PAIR>(d_tmem, da + adv, db + adv, idesc, (ki | j) != 0 ? 1u : 0u);
                        if (Cfg::X3 && !Cfg::XB) {
                            mma<true, PAIR>(d_tmem, dal + adv, db + adv, idesc, 1u);
                            mma<true, PAIR>(d_tmem, da + adv, dbl + adv, idesc, 1u);
                        }
                    }
                    if (Cfg::XB) {
                        // x tiles: bytes [0,64) of a row = bf16(x) for the chunk's 32 k, bytes [64,128) = bf16(lo); K = 16 per MMA
                        constexpr uint32_t idesc_bf = umma_idesc(1, PAIR ? 256 : 128, BN);
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            const uint64_t adv = (uint64_t)(jj * 2);
                            mma<false, PAIR>(d_tmem, dal + 4 + adv, dbl + adv, idesc_bf, 1u);      // A_lo * B
                            mma<false, PAIR>(d_tmem, dal + adv, dbl + 4 + adv, idesc_bf, 1u);      // A * B_lo
                        }
                    }
                    if (PAIR) umma_commit_2sm_mc(&empty_bar[s], (uint16_t)0x3);   // frees the stage in both CTAs
                    else if (MCAST) umma_commit_mc(&empty_bar[s], (uint16_t)0x3); // the peer multicasts into this stage too
                    else umma_commit(&empty_bar[s]);                    // stage reusable once these MMAs retire
                }
                if (PAIR) umma_commit_2sm_mc(&tmem_full_bar[as], (uint16_t)0x3);   // accumulator halves complete in both CTAs
                else umma_commit(&tmem_full_bar[as]);                   // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue warps (TMEM -> registers -> [smem -> TMA store |] NHWC global) =====================
        const int q = warp & 3;                                          // TMEM lane quadrant this warp may access
        const int m = q * 32 + lane;                                     // accumulator row = pixel inside the tile
        const int py = m / a.TW, px = m - py * a.TW;
        // each epilogue warp owns its 32 accumulator rows (= 2 image rows x 16 pixels of the tile) end to end: its own two 4 KB
        // staging buffers and its own TMA stores -- no barrier between the four warps
        uint8_t* const staging = smem_al + (size_t)RING_BYTES + (size_t)q * (2 * 32 * 128);
        uint32_t tcount = 0, chunk_id = 0;
        for (int w = cl_id; w < a.total_tiles; w += n_cl, ++tcount) {
            GLARE_DECODE_WORK(w)
            const int gy = ty * a.TH + py, gx = tx * a.TW + px;
            const bool valid = gy < a.H && gx < a.W && n < a.B;
            const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait_backoff(&tmem_full_bar[as], aph, 512);      // (producers / epilogue are ahead by design: poll with back-off, tc.cuh)
            tc_fence_after();
            const long long pix = ((long long)n * a.H * a.oscale + (gy * a.oscale + a.oa)) * (a.W * a.oscale) + (gx * a.oscale + a.ob);
            float* yrow = a.y + pix * a.ldy;
            const float* rrow = (a.residual && valid) ? a.residual + pix * a.Cout : nullptr;
            const bool epi_exp = (MODE == 4 || MODE == 0) && a.row_norm != nullptr;     // (mode 0: handled at the top of `process`)
            const bool epi_pack = MODE == 4 && (a.row_norm != nullptr || a.pack_out != 0);
            const bool epi_sq = MODE == 4 && a.row_sq_part != nullptr;
            float e_ref = 0.f, e_sum = 0.f, r_scale = 1.f, e_sq = 0.f;
            if (epi_exp && valid)                                        // key_norm_max == null: row_norm holds the reference itself (attn.cu, sampled maximum)
                e_ref = a.key_norm_max ? fmaf(__ldg(a.row_norm + pix), a.exp_scale * __uint_as_float(__ldg(a.key_norm_max)), -a.exp_margin)
                                       : __ldg(a.row_norm + pix);
            if (a.row_scale != nullptr) r_scale = valid ? __ldg(a.row_scale + pix) : 0.f;
            // one 32-column chunk of this thread's accumulator row: registers -> bias / residual / attention epilogues -> store
            auto process = [&](const uint32_t (&v)[32], const int c0) {
                const int co = nb * BN + c0;
                if constexpr (MODE == 0) {
                    if (a.row_norm != nullptr) {
                        // bf16 operands (BASELINE config 3): the scores GEMM's exp epilogue writes P~ as the single-piece bf16 operand of the
                        // P V GEMM.  Two 32-column chunks share one 128-byte staging row (64 bf16): the left chunk fills 16-byte chunks 0-3, the
                        // right one 4-7 and issues the TMA store (box = 64 keys x TW pixels x TH / 4 rows).  Keys >= Cout get 0.
                        const int right = (c0 >> 5) & 1;
                        const int co_pair = co - 32 * right;
                        if (co_pair >= a.Cout) return;                   // uniform over the CTA
                        const float c1 = a.exp_scale * 1.4426950408889634f, r1 = e_ref * 1.4426950408889634f;
                        float p[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) p[j] = ex2_approx((co + j < a.Cout) ? fmaf(__uint_as_float(v[j]), c1, -r1) : -INFINITY);
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // fixed order: repeatable row sums
#pragma unroll
                        for (int j = 0; j < 32; j += 4) { s0 += p[j]; s1 += p[j + 1]; s2 += p[j + 2]; s3 += p[j + 3]; }
                        e_sum += (s0 + s1) + (s2 + s3);
                        uint8_t* buf = staging + (chunk_id & 1) * (32 * 128);
                        if (!right) __syncwarp();                        // lane 0 has seen the previous store of this buffer drain
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            uint32_t w4[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const __nv_bfloat162 h = __floats2bfloat162_rn(p[8 * jj + 2 * e], p[8 * jj + 2 * e + 1]);
                                w4[e] = *reinterpret_cast<const uint32_t*>(&h);
                            }
                            *reinterpret_cast<uint4*>(buf + lane * 128 + (((4 * right + jj) ^ (lane & 7)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                        }
                        if (right) {
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
                                if (n < a.B) tma_store_4d(&tmY, buf, co_pair, tx * a.TW, ty * a.TH + (a.TH >> 2) * q, n);
                                tma_store_commit();
                                tma_store_wait_read<1>();
                            }
                            ++chunk_id;
                        }
                        return;
                    }
                }
                if (co >= a.Cout) return;                                // uniform over the CTA
                const bool full = co + 32 <= a.Cout;                     // uniform: no ragged tail inside this chunk
                float o[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
                if (full) {
                    if (a.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + co + j));
                            o[j] += bv.x; o[j + 1] += bv.y; o[j + 2] += bv.z; o[j + 3] += bv.w;
                        }
                    }
                    if (rrow) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 rv = __ldg(reinterpret_cast<const float4*>(rrow + co + j));
                            o[j] += rv.x; o[j + 1] += rv.y; o[j + 2] += rv.z; o[j + 3] += rv.w;
                        }
                    }
                } else {                                                 // last chunk of a Cout % 32 != 0 output (3-channel heads, attention keys)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (co + j < a.Cout) {
                            if (a.bias) o[j] += __ldg(a.bias + co + j);
                            if (rrow) o[j] += __ldg(rrow + co + j);
                        }
                    }
                }
                if (a.row_scale != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] *= r_scale;
                }
                if (epi_sq) {
                    float q0 = 0.f, q1 = 0.f;                           // columns past Cout hold exact zeros (zero-filled weight rows, no bias)
#pragma unroll
                    for (int j = 0; j < 32; j += 2) { q0 = fmaf(o[j], o[j], q0); q1 = fmaf(o[j + 1], o[j + 1], q1); }
                    e_sq += q0 + q1;
                }
                if (epi_pack) {
                    // columns [co, co + 32) of this row = one 128-byte operand chunk [16 words of a1 pairs | 16 words of a2 pairs]
                    float p[32];
                    if (epi_exp) {
                        // p = 2^(s * scale*log2(e) - ref*log2(e)): one FFMA + one MUFU.EX2 per element, branch-free (a single epilogue warp per
                        // scheduler has no other latency hiding).  The exponent's rounding (2^-24 |t|, |t| <~ 90) and ex2.approx (2^-22) stay
                        // below the 2^-17 resolution of the two-piece bf16 operand p is stored as.
                        const float c1 = a.exp_scale * 1.4426950408889634f, r1 = e_ref * 1.4426950408889634f;
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) p[j] = ex2_approx(fmaf(o[j], c1, -r1));
                        } else {                                        // ragged last chunk: keys past the end get exp2(-inf) = 0
#pragma unroll
                            for (int j = 0; j < 32; ++j) p[j] = ex2_approx((co + j < a.Cout) ? fmaf(o[j], c1, -r1) : -INFINITY);
                        }
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // fixed order: repeatable row sums
#pragma unroll
                        for (int j = 0; j < 32; j += 4) { s0 += p[j]; s1 += p[j + 1]; s2 += p[j + 2]; s3 += p[j + 3]; }
                        e_sum += (s0 + s1) + (s2 + s3);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) p[j] = o[j];
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(p[j], p[j + 1]);               // a1 pair (one F2FP)
                        const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h);
                        const float r0 = p[j] - __uint_as_float(hw << 16), r1 = p[j + 1] - __uint_as_float(hw & 0xffff0000u);
                        const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);                      // a2 pair
                        o[j >> 1] = __uint_as_float(hw);
                        o[16 + (j >> 1)] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
                    }
                }
                if (a.tma_store) {
                    uint8_t* buf = staging + (chunk_id & 1) * (32 * 128);
                    __syncwarp();                                        // lane 0 has seen the previous store of this buffer drain
                    if (a.gn_part != nullptr && !valid) {                // pixels outside the image are clipped by the store; keep them out of the sums
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)                          // SWIZZLE_128B: 16-byte chunk j of row r at chunk j ^ (r & 7)
                        *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (a.gn_part != nullptr) {
                        // GroupNorm statistics of what this conv produces (the next layer's Normalize, encoder_decoder.py:34-35), from the
                        // staged tile: lane l sums channel co + l over the warp's valid pixels -- column reads of the swizzled tile are
                        // conflict-free (32 distinct words of one 128-byte row per step) --, cpg adjacent lanes fold into their group, and the
                        // group's first lane writes the (tile, warp) partial.  No shuffles over pixels, no atomics.  Shifted-data sums: values
                        // are taken relative to c = the group's first channel at the warp's first pixel, so the fp32 partial keeps the
                        // variance when |mean| >> std; the finish kernel undoes the shift in fp64.  Partial = {S1, S2, c, n}.
                        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                        const float v0 = *reinterpret_cast<const float*>(buf + (((lane >> 2) ^ 0) << 4 | ((lane & 3) << 2)));      // row 0
                        const float cshift = __shfl_sync(0xffffffffu, v0, lane & ~(a.gn_cpg - 1));
                        float cs = 0.f, cq = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const float v = *reinterpret_cast<const float*>(buf + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                            const float dv = (vmask >> r) & 1u ? v - cshift : 0.f;
                            cs += dv;
                            cq = fmaf(dv, dv, cq);
                        }
                        for (int sh = 1; sh < a.gn_cpg; sh <<= 1) {
                            cs += __shfl_xor_sync(0xffffffffu, cs, sh);
                            cq += __shfl_xor_sync(0xffffffffu, cq, sh);
                        }
                        if ((lane & (a.gn_cpg - 1)) == 0 && n < a.B) {
                            float4* dst = reinterpret_cast<float4*>(a.gn_part) +
                                          ((long long)n * a.gn_slots + (long long)r_ * 4 + q) * 32 + (co + lane) / a.gn_cpg;
                            *dst = make_float4(cs, cq, cshift, (float)(__popc(vmask) * a.gn_cpg));
                        }
                    }
                    if (lane == 0) {
                        // box = 32 channels x TW pixels x TH / 4 rows; clips pixels / channels outside the tensor
                        if (n < a.B) tma_store_4d(&tmY, buf, co, tx * a.TW, ty * a.TH + (a.TH >> 2) * q, n);
                        tma_store_commit();
                        tma_store_wait_read<1>();                        // this warp's other staging buffer is free again
                    }
                    ++chunk_id;
                } else if (valid && !epi_pack) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (co + j + 4 <= a.Cout) *reinterpret_cast<float4*>(yrow + co + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                        else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (co + j + e < a.Cout) yrow[co + j + e] = o[j + e];
                        }
                    }
                }
            };
            // TMEM -> registers, double buffered: the load of chunk c + 1 is in flight while chunk c is processed
            uint32_t va[32], vb[32];
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
            tmem_ld_32x32(tacc, va);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 64) {
                tmem_ld_wait();
                tmem_ld_32x32(tacc + c0 + 32, vb);
                process(va, c0);
                tmem_ld_wait();
                if (c0 + 64 < BN) {
                    tmem_ld_32x32(tacc + c0 + 64, va);
                } else {                                                 // accumulator fully read: hand it back to the MMA thread
                    tc_fence_before();
                    if (PAIR) mbar_arrive_cluster(&tmem_empty_bar[as], 0);   // ... which lives in the leader CTA
                    else mbar_arrive(&tmem_empty_bar[as]);
                }
                process(vb, c0 + 32);
            }
            if (epi_exp && valid) a.row_sum_part[(long long)nb * a.part_stride + pix] = e_sum;
            if (epi_sq && valid) a.row_sq_part[(long long)nb * a.part_stride + pix] = e_sq;
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
#undef GLARE_DECODE_WORK
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                  // no CTA leaves while its peer can still signal its barriers
    if (warp == 1) {
        __syncwarp();
        if (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// GroupNorm statistics from the conv epilogue's partials {S1, S2, c, n} (shifted sums about c over n values): the true sums
// sum x = S1 + n c, sum x^2 = S2 + 2 c S1 + n c^2 per partial in fp64, reduced over the sample's slots in a fixed order
__global__ void __launch_bounds__(128) conv_gn_finish_kernel(const float4* __restrict__ part, int slots, double* __restrict__ stats) {
    __shared__ double s_s[128], s_q[128];
    const int b = blockIdx.x >> 5, g = blockIdx.x & 31;
    const float4* p = part + (long long)b * slots * 32 + g;
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < slots; i += 128) {
        const float4 v = __ldg(p + (long long)i * 32);
        const double S1 = v.x, S2 = v.y, c = v.z, n = v.w;
        s += S1 + n * c;
        q += S2 + 2.0 * c * S1 + n * c * c;
    }
    s_s[threadIdx.x] = s;
    s_q[threadIdx.x] = q;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_s[threadIdx.x] += s_s[threadIdx.x + o]; s_q[threadIdx.x] += s_q[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[(long long)blockIdx.x * 2] = s_s[0];
        stats[(long long)blockIdx.x * 2 + 1] = s_q[0];
    }
}

// ---- weight / activation packing ---------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) { return tf32_round(x); }

// OIHW fp32 -> [Cout][kh*kw][Cin]; mode 0: bf16; mode 1: fp32 (tensor core truncates to tf32); mode 2: hi (low 13
// mantissa bits cleared) + lo = w - hi (exact in fp32)
__global__ void conv_pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int kk, int mode,
                                        void* __restrict__ out_hi, float* __restrict__ out_lo) {
    const long long n = (long long)Cout * Cin * kk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        const long long r = i / Cin;
        const int t = (int)(r % kk), co = (int)(r / kk);
        const float v = w[((long long)co * Cin + ci) * kk + t];
        if (mode == 0) {
            reinterpret_cast<__nv_bfloat16*>(out_hi)[i] = __float2bfloat16_rn(v);
        } else if (mode == 1) {
            reinterpret_cast<float*>(out_hi)[i] = v;
        } else if (mode == 2) {
            const float h = tf32_hi(v);
            reinterpret_cast<float*>(out_hi)[i] = h;
            out_lo[i] = v - h;
        } else if (mode == 3) {
            const float h = tf32_hi(v);
            reinterpret_cast<float*>(out_hi)[i] = h;
            __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(out_lo);
            const long long xi = (i >> 5) * 64 + (i & 31);
            xb[xi] = __float2bfloat16_rn(v);
            xb[xi + 32] = __float2bfloat16_rn(v - h);
        } else {
            __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(out_hi);
            const long long xi = (i >> 5) * 64 + (i & 31);
            split_b3(v, xb[xi], xb[xi + 32]);
        }
    }
}

// elementwise operand preparation for activations that do not come out of the GroupNorm kernel
__global__ void conv_prep_act_kernel(const float4* __restrict__ x, long long n4, int mode, void* __restrict__ out_hi,
                                     float4* __restrict__ out_lo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        if (mode == 0) {
            __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&a);
            o.y = *reinterpret_cast<uint32_t*>(&b);
            reinterpret_cast<uint2*>(out_hi)[i] = o;
        } else if (mode == 2) {
            const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            reinterpret_cast<float4*>(out_hi)[i] = h;
            out_lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        } else if (mode == 3) {
            const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            reinterpret_cast<float4*>(out_hi)[i] = h;
            store_x4(reinterpret_cast<__nv_bfloat16*>(out_lo), i * 4, v.x, v.y, v.z, v.w, h.x, h.y, h.z, h.w);
        } else {
            store_b3_4(reinterpret_cast<__nv_bfloat16*>(out_hi), i * 4, v.x, v.y, v.z, v.w);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_act_map(CUtensorMap* m, const void* ptr, bool bf16, int B, int H, int W, int C, int TH, int TW, int stride, long long ldc = 0) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GLARE_ERR_UNSUPPORTED;
    const cuuint64_t es = bf16 ? 2 : 4;
    const cuuint64_t ld = ldc > 0 ? (cuuint64_t)ldc : (cuuint64_t)C;      // pixel stride in elements (> C: a column slice of a wider matrix)
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {ld * es, (cuuint64_t)W * ld * es, (cuuint64_t)H * W * ld * es};
    // traversal stride s: the box spans TW*s x TH*s input pixels and delivers every s-th one -> TW x TH rows in smem
    cuuint32_t box[4] = {(cuuint32_t)(128 / es), (cuuint32_t)(TW * stride), (cuuint32_t)(TH * stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GLARE_OK : GLARE_ERR_BAD_ARG;
}

static int make_w_map(CUtensorMap* m, const void* ptr, bool bf16, int Cout, int K, int BN, int n_w, long long w_batch_stride, long long ldk = 0) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GLARE_ERR_UNSUPPORTED;
    const cuuint64_t es = bf16 ? 2 : 4;
    const cuuint64_t ld = ldk > 0 ? (cuuint64_t)ldk : (cuuint64_t)K;      // row stride in elements (> K: a column slice of a wider matrix)
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Cout, (cuuint64_t)n_w};
    cuuint64_t strides[2] = {ld * es, (cuuint64_t)(n_w > 1 ? w_batch_stride : (long long)ld * Cout) * es};
    cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GLARE_OK : GLARE_ERR_BAD_ARG;
}

static int make_out_map(CUtensorMap* m, const float* ptr, int B, int H, int W, int Cout, long long ldy, int TH, int TW) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GLARE_ERR_UNSUPPORTED;
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ldy * 4, (cuuint64_t)W * ldy * 4, (cuuint64_t)H * W * ldy * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)TW, (cuuint32_t)(TH / 4), 1};       // one epilogue warp: 32 tile rows = TH/4 image rows x TW pixels
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GLARE_OK : GLARE_ERR_BAD_ARG;
}

// P~ operand of the bf16 mode: bf16 [B,H,W,ldy], one epilogue warp stores 64 keys x TW pixels x TH / 4 rows per instruction
static int make_out_map_bf16(CUtensorMap* m, const void* ptr, int B, int H, int W, int n_cols, long long ldy, int TH, int TW) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GLARE_ERR_UNSUPPORTED;
    cuuint64_t dims[4] = {(cuuint64_t)n_cols, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ldy * 2, (cuuint64_t)W * ldy * 2, (cuuint64_t)H * W * ldy * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)(TH / 4), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GLARE_OK : GLARE_ERR_BAD_ARG;
}

template <int MODE, int BN, int CM, bool HALO = false, bool RING2 = false>
static int launch_conv_cl(const CUtensorMap& tA, const CUtensorMap& tAl, const CUtensorMap& tB, const CUtensorMap& tBl,
                          const CUtensorMap& tY, const ConvTcArgs& a, cudaStream_t stream) {
    constexpr int CL = CM == 0 ? 1 : 2;
    using Cfg = ConvCfg<MODE, BN, CM == 2>;
    static_assert(Cfg::STAGES >= 2, "pipeline needs at least two stages");
    static_assert(!HALO || Cfg::HALO_STAGES >= 3, "halo pipeline needs at least three stages");
    constexpr int SMEM = RING2 ? Cfg::R2_SMEM_DYN : (HALO ? Cfg::HALO_SMEM_DYN : Cfg::SMEM_DYN);
    auto kern = conv_tc_kernel<MODE, BN, CM, HALO, RING2>;
    GLARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int n_cl = kNumSMs / CL;
    const int clusters = a.total_tiles < n_cl ? a.total_tiles : n_cl;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * CL));
    cfg.blockDim = dim3(CT_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GLARE_CUDA(cudaLaunchKernelEx(&cfg, kern, tA, tAl, tB, tBl, tY, a));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

template <int MODE, int BN>
static int launch_conv(int cl, const CUtensorMap& tA, const CUtensorMap& tAl, const CUtensorMap& tB, const CUtensorMap& tBl,
                       const CUtensorMap& tY, const ConvTcArgs& a, cudaStream_t stream) {
    // cl: 1 = single CTA, 2 = multicast cluster, 3 = CTA pair (cta_group::2), 4 = CTA pair with filter-row halo staging
    if (cl == 4) {
        if constexpr ((MODE == 4 || MODE == 0) && BN <= 128) return launch_conv_cl<MODE, BN, 2, true>(tA, tAl, tB, tBl, tY, a, stream);
        else return GLARE_ERR_UNSUPPORTED;
    }
    if (cl == 5) {                                                   // two-ring patch staging for the 256-wide N tile
        if constexpr ((MODE == 4 || MODE == 0) && BN == 256) return launch_conv_cl<MODE, BN, 2, false, true>(tA, tAl, tB, tBl, tY, a, stream);
        else return GLARE_ERR_UNSUPPORTED;
    }
    if (cl == 3) return launch_conv_cl<MODE, BN, 2>(tA, tAl, tB, tBl, tY, a, stream);
    return cl == 2 ? launch_conv_cl<MODE, BN, 1>(tA, tAl, tB, tBl, tY, a, stream) : launch_conv_cl<MODE, BN, 0>(tA, tAl, tB, tBl, tY, a, stream);
}

}  // namespace glare

using namespace glare;

// bytes per element of the packed operand for a precision mode (0 bf16, 1 tf32, 2 3xtf32)
GLARE_API int glare_conv_tc_elem_bytes(int mode) { return mode == 0 ? 2 : 4; }   // hi operand; the mode-3 x tensor is bf16 x 2

GLARE_API int glare_conv_pack_weight(int mode, const float* w_oihw, int Cout, int Cin, int ksize, void* out_hi, void* out_lo,
                                     cudaStream_t stream) {
    if (!w_oihw || !out_hi || ((mode == 2 || mode == 3) && !out_lo) || mode < 0 || mode > 4 || Cout <= 0 || Cin <= 0 || ksize < 1 || ksize > 3)
        return GLARE_ERR_BAD_ARG;
    if (mode >= 3 && Cin % 32 != 0) return GLARE_ERR_UNSUPPORTED;
    const long long n = (long long)Cout * Cin * ksize * ksize;
    const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    conv_pack_weight_kernel<<<grid, 256, 0, stream>>>(w_oihw, Cout, Cin, ksize * ksize, mode, out_hi, reinterpret_cast<float*>(out_lo));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// fp32 activations -> tensor-core operand(s): mode 0 bf16 copy, mode 2 tf32 hi/lo split (mode 1 needs no preparation)
GLARE_API int glare_conv_prep_act(int mode, const float* x, long long n, void* out_hi, void* out_lo, cudaStream_t stream) {
    if (n < 0 || (mode != 0 && mode != 2 && mode != 3 && mode != 4) || (n & 3) || (mode >= 3 && (n & 31))) return GLARE_ERR_BAD_ARG;
    if (n == 0) return GLARE_OK;
    if (!x || !out_hi || ((mode == 2 || mode == 3) && !out_lo)) return GLARE_ERR_BAD_ARG;
    const long long n4 = n / 4;
    const int grid = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    conv_prep_act_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), n4, mode, out_hi, reinterpret_cast<float4*>(out_lo));
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// x / x_lo: NHWC [B,H,W,Cin] (bf16 for mode 0, fp32 otherwise; x_lo only for mode 2); w / w_lo: packed by
// glare_conv_pack_weight; bias [Cout] / residual NHWC [B,H,W,Cout] fp32 or NULL; y NHWC fp32.
struct TapSpec { int ntaps, tap_w, dy0, dx0, oscale, oa, ob; };
struct AttnEpi {                       // attention epilogues of the two GEMMs (ConvTcArgs documents the fields)
    const float* row_norm; const unsigned* key_norm_max; float* row_sum_part; long long part_stride; float exp_scale, exp_margin;
    const float* row_scale; int* n_blocks_out; int pack_out; float* row_sq_part;
    long long ldx, ldw;                // row strides (logical elements) of x / w when they are column slices of wider matrices (0: dense)
};
static int conv_tc_launch(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo, const float* bias,
                          const float* residual, float* y, int B, int Hin, int Win, int H, int W, int Cin, int Cout, TapSpec ts,
                          int stride, long long ldy, long long w_batch_stride, cudaStream_t stream, double* gn_stats = nullptr,
                          float* gn_scratch = nullptr, long long gn_scratch_floats = 0, const AttnEpi* ae = nullptr);
static inline TapSpec std_taps(int ksize) { return TapSpec{ksize * ksize, ksize, -(ksize / 2), -(ksize / 2), 1, 0, 0}; }

GLARE_API int glare_conv2d_nhwc_tc(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo, const float* bias,
                                   const float* residual, float* y, int B, int H, int W, int Cin, int Cout, int ksize,
                                   cudaStream_t stream) {
    if (ksize != 1 && ksize != 3) return GLARE_ERR_BAD_ARG;
    return conv_tc_launch(mode, x, x_lo, w, w_lo, bias, residual, y, B, H, W, H, W, Cin, Cout, std_taps(ksize), 1, Cout, 0, stream);
}

// Downsample.forward (encoder_decoder.py:68-72): zero pad (0,1,0,1) then 3x3 stride-2 conv, pad 0.  x NHWC [B,Hin,Win,Cin];
// y NHWC [B,Hout,Wout,Cout] with Hout = (Hin + 1 - 3) / 2 + 1.  The padding is the TMA unit's out-of-bounds zero fill.
GLARE_API int glare_conv2d_nhwc_tc_down2(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo,
                                         const float* bias, float* y, int B, int Hin, int Win, int Cin, int Cout, cudaStream_t stream) {
    if (Hin < 2 || Win < 2) return GLARE_ERR_BAD_ARG;
    const int Ho = (Hin + 1 - 3) / 2 + 1, Wo = (Win + 1 - 3) / 2 + 1;
    return conv_tc_launch(mode, x, x_lo, w, w_lo, bias, nullptr, y, B, Hin, Win, Ho, Wo, Cin, Cout, TapSpec{9, 3, 0, 0, 1, 0, 0}, 2, Cout, 0,
                          stream);
}

// One sub-pixel phase (a, b) in {0,1}^2 of Upsample.forward (encoder_decoder.py:49-53: nearest x2, then 3x3 conv, pad 1):
// output pixel (2i + a, 2j + b) only sees a 2x2 neighbourhood of the LOW-resolution input, rows {i + a - 1, i + a}, columns
// {j + b - 1, j + b}, with the 3x3 taps that land on the same source pixel pre-summed into a 2x2 filter (host side, once per
// weight).  Four launches (one per phase) replace the upsample copy + 3x3 conv at 4/9 of the FLOPs.  x NHWC [B,H,W,Cin] (low
// resolution), w packed [Cout][4][Cin] for this phase, y NHWC [B,2H,2W,Cout] (full tensor; this call writes one phase).
GLARE_API int glare_conv2d_nhwc_tc_up2_phase(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo,
                                             const float* bias, float* y, int B, int H, int W, int Cin, int Cout, int a, int b,
                                             cudaStream_t stream) {
    if ((a | b) & ~1) return GLARE_ERR_BAD_ARG;
    return conv_tc_launch(mode, x, x_lo, w, w_lo, bias, nullptr, y, B, H, W, H, W, Cin, Cout, TapSpec{4, 2, a - 1, b - 1, 2, a, b}, 1, Cout, 0,
                          stream);
}

// General form used by the Python host: kind 0 = 3x3 / 1x1 stride-1 conv, 1 = Downsample conv (stride 2), 2 = one sub-pixel phase
// (pa, pb) of Upsample+conv.  gn_stats (optional, [B][32][2] fp64, kinds 0 and 1, Cout % 128 == 0): GroupNorm(32) sum / sum of squares of
// the OUTPUT, taken from the epilogue's staged tiles as per-(tile, warp) partials in gn_scratch (>= glare_conv_gn_scratch_floats(B, H, W)
// floats for the OUTPUT size) and reduced in fp64 by a second small launch -- deterministic, no atomics.
GLARE_API int glare_conv2d_nhwc_tc_g(int mode, int kind, const void* x, const void* x_lo, const void* w, const void* w_lo,
                                     const float* bias, const float* residual, float* y, int B, int Hin, int Win, int Cin, int Cout,
                                     int ksize, int pa, int pb, double* gn_stats, float* gn_scratch, long long gn_scratch_floats,
                                     cudaStream_t stream) {
    if (kind == 0) {
        if (ksize != 1 && ksize != 3) return GLARE_ERR_BAD_ARG;
        return conv_tc_launch(mode, x, x_lo, w, w_lo, bias, residual, y, B, Hin, Win, Hin, Win, Cin, Cout, std_taps(ksize), 1, Cout, 0, stream,
                              gn_stats, gn_scratch, gn_scratch_floats);
    }
    if (kind == 1) {
        if (Hin < 2 || Win < 2 || residual) return GLARE_ERR_BAD_ARG;
        const int Ho = (Hin + 1 - 3) / 2 + 1, Wo = (Win + 1 - 3) / 2 + 1;
        return conv_tc_launch(mode, x, x_lo, w, w_lo, bias, nullptr, y, B, Hin, Win, Ho, Wo, Cin, Cout, TapSpec{9, 3, 0, 0, 1, 0, 0}, 2, Cout, 0,
                              stream, gn_stats, gn_scratch, gn_scratch_floats);
    }
    if (kind == 2) {
        if (((pa | pb) & ~1) || residual) return GLARE_ERR_BAD_ARG;
        if (gn_stats) return GLARE_ERR_UNSUPPORTED;          // the sub-pixel phases store directly (no staged tile to take the sums from)
        return conv_tc_launch(mode, x, x_lo, w, w_lo, bias, nullptr, y, B, Hin, Win, Hin, Win, Cin, Cout,
                              TapSpec{4, 2, pa - 1, pb - 1, 2, pa, pb}, 1, Cout, 0, stream);
    }
    return GLARE_ERR_BAD_ARG;
}

// Extended form: ldy = output pixel stride in elements (>= Cout, multiple of 4); w_batch_stride != 0 selects per-sample
// weights w[n] = w + n * w_batch_stride elements (the two attention GEMMs: S = Q K^T with W = K[n], O = P V with
// W = V[n]^T -- encoder_decoder.py:176-187).  residual must be null when ldy != Cout.
GLARE_API int glare_conv2d_nhwc_tc_ex(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo,
                                      const float* bias, const float* residual, float* y, int B, int H, int W, int Cin, int Cout,
                                      int ksize, long long ldy, long long w_batch_stride, cudaStream_t stream) {
    if (ksize != 1 && ksize != 3) return GLARE_ERR_BAD_ARG;
    return conv_tc_launch(mode, x, x_lo, w, w_lo, bias, residual, y, B, H, W, H, W, Cin, Cout, std_taps(ksize), 1, ldy, w_batch_stride,
                          stream);
}

// Attention scores with the softmax numerator fused into the epilogue (mode 4 operands; attn.cu has the scheme and the kernels around it).
// q: operand [rows_h * rows_w][C] of the query rows, k: operand [n_keys][C] of this sample's keys; writes the bf16x3 operand
// P~ [rows][n_pad] = exp(scale * q k^T - ref(row)) (zero for keys in [n_keys, roundup32(n_keys)); chunks past that are not touched) and
// row_sum_part[nb][row], nb < *n_blocks_host, the partial row sums per output block (part_stride >= rows elements apart).
GLARE_API int glare_attn_scores_exp_tc(int mode, const void* q, const void* k, int rows_h, int rows_w, int C, int n_keys, int n_pad,
                                       float scale, float margin, const float* q_row_norm, const unsigned* key_norm_max, void* p_out,
                                       float* row_sum_part, long long part_stride, int* n_blocks_host, cudaStream_t stream) {
    if (mode != 4 && mode != 0) return GLARE_ERR_UNSUPPORTED;
    if (!q_row_norm || !row_sum_part || !n_blocks_host || n_pad < n_keys || (n_pad & 31) || !(scale > 0.f) || !(margin >= 0.f))
        return GLARE_ERR_BAD_ARG;                            // key_norm_max may be null: q_row_norm then holds the per-row reference itself
    AttnEpi ae{q_row_norm, key_norm_max, row_sum_part, part_stride, scale, margin, nullptr, n_blocks_host, 0, nullptr, 0, 0};
    return conv_tc_launch(mode, q, nullptr, k, nullptr, nullptr, nullptr, reinterpret_cast<float*>(p_out), 1, rows_h, rows_w, rows_h, rows_w, C,
                          n_keys, std_taps(1), 1, n_pad, 0, stream, nullptr, nullptr, 0, &ae);
}

// O = diag(row_scale) (P~ V + residual): p operand [rows] x n_keys (row stride ldp elements), vt operand [C] x n_keys (V^T of the sample,
// row stride ldvt); y [rows][ldy] fp32, or with pack_out != 0 (ldy == C, C % 32 == 0) the bf16x3 operand of the following proj_out conv.
// Key bands: the tensor core's fp32 accumulator truncates, a bias that grows with the contraction length (1080p: 131 648 keys -> 1.3e-3 of
// the output scale, tests/test_fullsize_parity_gpu.py); the host therefore contracts at most a few 10^4 keys per launch and chains the
// launches through `residual` (fp32 [rows][C], added round-to-nearest in the epilogue BEFORE the row scale): bands 1..n-1 pass
// row_scale = NULL and write the running sum, the last band passes row_scale (and pack_out).  ldp / ldvt = 0: dense (= n_keys).
GLARE_API int glare_attn_pv_tc(int mode, const void* p, long long ldp, const void* vt, long long ldvt, const float* row_scale, const float* residual,
                               void* y, int rows_h, int rows_w, int n_keys, int C, long long ldy, int pack_out, cudaStream_t stream) {
    if (mode != 4 && !(mode == 0 && !pack_out)) return GLARE_ERR_UNSUPPORTED;
    if ((pack_out && !row_scale) || (ldp && ldp < n_keys) || (ldvt && ldvt < n_keys) || (residual && ldy != C)) return GLARE_ERR_BAD_ARG;
    AttnEpi ae{nullptr, nullptr, nullptr, 0, 0.f, 0.f, row_scale, nullptr, pack_out ? 1 : 0, nullptr, ldp, ldvt};
    return conv_tc_launch(mode, p, nullptr, vt, nullptr, nullptr, residual, reinterpret_cast<float*>(y), 1, rows_h, rows_w, rows_h, rows_w, n_keys, C,
                          std_taps(1), 1, ldy, 0, stream, nullptr, nullptr, 0, &ae);
}

// Stride-1 conv (ksize 1 or 3) whose output goes straight to another tensor-core GEMM: y is written as that GEMM's bf16x3 operand (NHWC
// [B,H,W,Cout], 4 bytes per element, Cout % 32 == 0) instead of fp32 -- the q / k projections of AttnBlock (encoder_decoder.py:172-174), the
// WarpBlock offset conv feeding conv_offset (deformableDecoder_arch.py:285-287).  row_sq_part (optional, [n_blocks][part_stride], part_stride >=
// B*H*W): per output row the partial sums of squares of the fp32 values per output block, *n_blocks_host planes (the attention row reference).
GLARE_API int glare_conv2d_nhwc_tc_pack(int mode, const void* x, const void* w, const float* bias, void* y_operand, int B, int H, int W, int Cin,
                                        int Cout, int ksize, float* row_sq_part, long long part_stride, int* n_blocks_host, cudaStream_t stream) {
    if (mode != 4) return GLARE_ERR_UNSUPPORTED;
    if ((ksize != 1 && ksize != 3) || (row_sq_part && !n_blocks_host)) return GLARE_ERR_BAD_ARG;
    AttnEpi ae{nullptr, nullptr, nullptr, part_stride, 0.f, 0.f, nullptr, n_blocks_host, 1, row_sq_part, 0, 0};
    return conv_tc_launch(mode, x, nullptr, w, nullptr, bias, nullptr, reinterpret_cast<float*>(y_operand), B, H, W, H, W, Cin, Cout, std_taps(ksize), 1,
                          Cout, 0, stream, nullptr, nullptr, 0, &ae);
}

static int conv_tc_launch(int mode, const void* x, const void* x_lo, const void* w, const void* w_lo, const float* bias,
                          const float* residual, float* y, int B, int Hin, int Win, int H, int W, int Cin, int Cout, TapSpec ts,
                          int stride, long long ldy, long long w_batch_stride, cudaStream_t stream, double* gn_stats, float* gn_scratch,
                          long long gn_scratch_floats, const AttnEpi* ae) {
    const int ksize = ts.tap_w;
    const bool epi_exp = ae && ae->row_norm;
    const bool epi_pack = ae && (ae->row_norm || ae->pack_out);          // the output is an operand tensor
    // attention epilogues: mode 4 (all of them) and mode 0 (exp with a bf16 P~ operand; row scale) -- not the operand-packing / row-square ones
    if (ae && mode != 4 && !(mode == 0 && !ae->pack_out && !ae->row_sq_part)) return GLARE_ERR_UNSUPPORTED;
    if (epi_exp && mode == 0 && (ldy & 63)) return GLARE_ERR_BAD_ARG;
    if (ae && (ae->row_norm || ae->row_scale) && (bias || gn_stats || (residual && ae->row_norm))) return GLARE_ERR_UNSUPPORTED;
    if (ae && ae->pack_out && !ae->row_norm && ((Cout & 31) || ldy != Cout || gn_stats || ts.oscale != 1)) return GLARE_ERR_UNSUPPORTED;
    if (ae && ae->row_sq_part && ae->part_stride < (long long)B * H * W) return GLARE_ERR_BAD_ARG;
    if (epi_exp && (!ae->row_sum_part || ae->part_stride < (long long)B * H * W || (ldy & 31) || Cout < 32))
        return GLARE_ERR_BAD_ARG;
    if (gn_stats && (Cout % 128 != 0 || Cout > 512 || B <= 0 || !gn_scratch || ldy != Cout || ts.oscale != 1)) return GLARE_ERR_UNSUPPORTED;
    if (ldy < Cout || (ldy & 3) || w_batch_stride < 0 || (residual && ldy != Cout)) return GLARE_ERR_BAD_ARG;
    if (mode < 0 || mode > 4 || B < 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (ksize < 1 || ksize > 3)) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!x || !w || !y || ((mode == 2 || mode == 3) && (!x_lo || !w_lo))) return GLARE_ERR_BAD_ARG;
    const int bke = mode == 0 ? 64 : 32;
    if (Cin % bke != 0 || (Cout % 4 != 0 && ldy == Cout)) return GLARE_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15) return GLARE_ERR_BAD_ARG;
    ConvTcArgs a{};
    a.bias = bias; a.residual = residual; a.y = y;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize; a.pad = -ts.dy0; a.stride = stride;
    a.gn_part = nullptr; a.gn_cpg = Cout / 32;
    a.ntaps = ts.ntaps; a.tap_w = ts.tap_w; a.tap_dy0 = ts.dy0; a.tap_dx0 = ts.dx0;
    a.oscale = ts.oscale; a.oa = ts.oa; a.ob = ts.ob;
    if (residual && ts.oscale != 1) return GLARE_ERR_BAD_ARG;
    a.TH = 8; a.TW = 16;
    a.tiles_x = (W + a.TW - 1) / a.TW; a.tiles_y = (H + a.TH - 1) / a.TH;
    a.kchunks = Cin / bke;
    // N tile: as wide as Cout allows, narrowed while the launch would leave SMs without a tile (short-M GEMMs: O = P V)
    int BN = Cout >= 256 ? 256 : (Cout > 64 ? 128 : 64);
    long long m_tiles = (long long)B * a.tiles_y * a.tiles_x;
    while (BN > 64 && m_tiles * ((Cout + BN - 1) / BN) < kNumSMs) BN >>= 1;
    // filter-row halo staging (ConvCfg): 3x3 stride-1 convs with narrow N tiles, 8 x 16 pixel tiles
    static const bool no_halo = getenv("GLARE_CONV_NO_HALO") != nullptr;            // A/B switch for profiling only
    const bool halo = !no_halo && (mode == 4 || mode == 0) && BN <= 128 && ts.ntaps == 9 && ts.tap_w == 3 && stride == 1 && ts.oscale == 1 &&
                      w_batch_stride == 0 && getenv("GLARE_CONV_NO_CLUSTER") == nullptr && getenv("GLARE_CONV_MCAST") == nullptr;
    // two-ring patch staging for the 256-wide N tile: default since round 2 (green on hardware, conv_tc 451.4 -> 447.7 ms/step,
    // profiles/r45_bench_ring2.json); GLARE_CONV_NO_RING2 selects the one-ring kernel (A/B switch)
    static const bool want_ring2 = getenv("GLARE_CONV_NO_RING2") == nullptr;
    const bool ring2 = want_ring2 && !no_halo && (mode == 4 || mode == 0) && BN == 256 && ts.ntaps == 9 && ts.tap_w == 3 && stride == 1 && ts.oscale == 1 &&
                       w_batch_stride == 0 && !ae && getenv("GLARE_CONV_NO_CLUSTER") == nullptr && getenv("GLARE_CONV_MCAST") == nullptr;
    if (halo || ring2) {
        a.TH = 16; a.TW = 8;
        a.tiles_x = (W + a.TW - 1) / a.TW; a.tiles_y = (H + a.TH - 1) / a.TH;
        m_tiles = (long long)B * a.tiles_y * a.tiles_x;
    }
    a.n_blocks = (Cout + BN - 1) / BN;
    if (ae) {
        a.row_norm = ae->row_norm; a.key_norm_max = ae->key_norm_max; a.row_sum_part = ae->row_sum_part; a.part_stride = ae->part_stride;
        a.exp_scale = ae->exp_scale; a.exp_margin = ae->exp_margin; a.row_scale = ae->row_scale;
        a.pack_out = ae->pack_out; a.row_sq_part = ae->row_sq_part;
        if (ae->n_blocks_out) *ae->n_blocks_out = a.n_blocks;
    }
    // clusters of two CTAs share the weight tile by multicast; per-sample weights over a batch cannot be shared across samples
    // default: CTA pair (cta_group::2).  A/B switches for profiling only: GLARE_CONV_MCAST -> two independent CTAs sharing the weight
    // tile by TMA multicast, GLARE_CONV_NO_CLUSTER -> single CTAs
    static const bool no_cluster = getenv("GLARE_CONV_NO_CLUSTER") != nullptr;
    static const bool use_mcast = getenv("GLARE_CONV_MCAST") != nullptr;
    int cl = (no_cluster || (w_batch_stride != 0 && B > 1) || m_tiles < 2) ? 1 : (use_mcast ? 2 : 3);
    const bool use_halo = (halo || ring2) && cl == 3;
    if ((halo || ring2) && !use_halo) {           // single tile: back to the standard geometry
        a.TH = 8; a.TW = 16;
        a.tiles_x = (W + a.TW - 1) / a.TW; a.tiles_y = (H + a.TH - 1) / a.TH;
        m_tiles = (long long)B * a.tiles_y * a.tiles_x;
    }
    if (use_halo) cl = ring2 ? 5 : 4;
    const int csz = cl == 1 ? 1 : 2;              // CTAs per cluster (box rows of the weight maps = BN / csz in both cluster modes)
    const long long total = (long long)a.n_blocks * ((m_tiles + csz - 1) / csz);
    if (total > 0x7fffffff || m_tiles > 0x7fffffff) return GLARE_ERR_UNSUPPORTED;
    a.total_tiles = (int)total;
    a.m_tiles = (int)m_tiles;
    a.ldy = ldy;
    a.w_batched = w_batch_stride != 0 ? 1 : 0;
    const int n_w = a.w_batched ? B : 1;
    CUtensorMap tA, tAl, tB, tBl, tY;
    int rc;
    {
        static const bool direct = getenv("GLARE_CONV_DIRECT_STORE") != nullptr;   // A/B switch for profiling only
        a.tma_store = ((!direct || epi_pack) && Cout >= 32 && ts.oscale == 1) ? 1 : 0;
    }
    // exp epilogue: the output is an operand tensor of 128-byte chunks (32 keys each), written whole up to the padded row length ldy
    if (epi_exp && mode == 0) {
        if ((rc = make_out_map_bf16(&tY, y, B, H, W, (int)ldy, ldy, a.TH, a.TW)) != GLARE_OK) return rc;
    } else if ((rc = make_out_map(&tY, y, B, H, W, epi_exp ? (int)ldy : Cout, ldy, a.TH, a.TW)) != GLARE_OK) return rc;
    const bool bf = mode == 0 || mode == 4;
    const int e2 = mode == 4 ? 2 : 1;              // mode 4: the operand tensors are interleaved bf16 pairs, 2 per element
    if ((rc = make_act_map(&tA, x, bf, B, Hin, Win, e2 * Cin, a.TH, use_halo ? a.TW + 2 : a.TW, stride, ae ? e2 * ae->ldx : 0)) != GLARE_OK) return rc;
    if ((rc = make_w_map(&tB, w, bf, Cout, e2 * ts.ntaps * Cin, BN / csz, n_w, e2 * w_batch_stride, ae ? e2 * ae->ldw : 0)) != GLARE_OK) return rc;
    tAl = tA; tBl = tB;
    if (mode == 2) {
        if ((rc = make_act_map(&tAl, x_lo, false, B, Hin, Win, Cin, a.TH, a.TW, stride)) != GLARE_OK) return rc;
        if ((rc = make_w_map(&tBl, w_lo, false, Cout, ts.ntaps * Cin, BN / csz, n_w, w_batch_stride)) != GLARE_OK) return rc;
    } else if (mode == 3) {                        // interleaved bf16 x tensors: 2 bf16 per element, same bytes per row as fp32
        if ((rc = make_act_map(&tAl, x_lo, true, B, Hin, Win, 2 * Cin, a.TH, a.TW, stride)) != GLARE_OK) return rc;
        if ((rc = make_w_map(&tBl, w_lo, true, Cout, 2 * ts.ntaps * Cin, BN / csz, n_w, 2 * w_batch_stride)) != GLARE_OK) return rc;
    }
    if (gn_stats) {
        if (!a.tma_store) return GLARE_ERR_UNSUPPORTED;
        a.gn_slots = a.tiles_y * a.tiles_x * 4;
        if ((long long)B * a.gn_slots * 128 > gn_scratch_floats) return GLARE_ERR_BAD_ARG;
        a.gn_part = gn_scratch;
    }
#define GLARE_CONV_DISPATCH(M)                                                            \
    do {                                                                                  \
        if (BN == 256) rc = launch_conv<M, 256>(cl, tA, tAl, tB, tBl, tY, a, stream);     \
        else if (BN == 128) rc = launch_conv<M, 128>(cl, tA, tAl, tB, tBl, tY, a, stream);\
        else rc = launch_conv<M, 64>(cl, tA, tAl, tB, tBl, tY, a, stream);                \
    } while (0)
    if (mode == 0) GLARE_CONV_DISPATCH(0);
    else if (mode == 1) GLARE_CONV_DISPATCH(1);
    else if (mode == 2) GLARE_CONV_DISPATCH(2);
    else if (mode == 3) GLARE_CONV_DISPATCH(3);
    else GLARE_CONV_DISPATCH(4);
#undef GLARE_CONV_DISPATCH
    if (rc != GLARE_OK || !gn_stats) return rc;
    conv_gn_finish_kernel<<<(unsigned)(B * 32), 128, 0, stream>>>(reinterpret_cast<const float4*>(gn_scratch), a.gn_slots, gn_stats);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// floats of scratch glare_conv2d_nhwc_tc_g needs for the fused GroupNorm statistics of an output of B x H x W pixels (both tile geometries)
GLARE_API long long glare_conv_gn_scratch_floats(int B, int H, int W) {
    const long long t1 = (long long)((H + 7) / 8) * ((W + 15) / 16), t2 = (long long)((H + 15) / 16) * ((W + 7) / 8);
    return (long long)B * (t1 > t2 ? t1 : t2) * 4 * 128;
}
