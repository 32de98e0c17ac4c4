// az_nn_tc.cu — the residual tower of the policy/value network on 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM, operands streamed into shared memory by bulk async copies).
//
// The network is a 3x3 conv ResNet on a 7x6 board (python/src/build_graph.py:54-90); each tower
// convolution is the only dense contraction on the self-play path: an implicit GEMM with
// M = positions, N = 256 output channels, K = 9 taps x 256 input channels.
//
// Layout that makes the implicit GEMM a set of shifted views (no im2col, no gather):
//   * activations are bf16, stored [channel chunk of 8][row][8 channels] (16-byte row cells), i.e. the canonical K-major
//     SWIZZLE_NONE UMMA operand layout with an 8-row core-matrix stride of 128 B: a row shift is just a different
//     16-byte-aligned start address in the shared-memory descriptor, so tap (ky, kx) is the SAME operand tile, moved.
//   * 48-row board layout (AZ_TC_RPB, az_nn.cuh): cell (y, x) -> row y*6 + x, then six zero rows.  The zero rows are the bottom border of the
//     board and the top border of the next one; the left / right border comes from two masked copies of the operand built in
//     shared memory (kx = 0 taps read the copy with the x = 5 cells zeroed, kx = 2 taps the copy with the x = 0 cells zeroed).
//     42 of 48 rows carry data.
//   * weights are pre-packed per layer into contiguous pipeline stages of K = 32 input channels of one tap:
//     [half of the output channels][chunk][128 out][8] = 8 KB per CTA of a pair (tower), [chunk][256 out][8] (stem), each fetched
//     with one cp.async.bulk.
//   (The first tower kernel — one CTA per tile on a 56-row layout with a zero column instead of masked copies — and the
//   layer-to-layer pipelining experiment are history: numbers in profiles/README.md, code in the round-1 tree.)
//
// Kernels: k_nn_conv_tc3 — the tower on CTA PAIRS (tcgen05 cta_group::2), operand ring over K groups;
// k_nn_conv_tc — one CTA per tile: the stem (13 -> 256 channels, 2 input chunks, BatchNorm indexed by board row; its three operand
// copies come from global memory).  Both are persistent, one CTA per SM: warp 0 streams
// operands, warp 1 issues tcgen05.mma (one elected lane) into one of two 256-column TMEM accumulators and owns the TMEM
// allocation, warps 2-5 run the epilogue of the PREVIOUS tile concurrently: tcgen05.ld -> folded BatchNorm -> (+ residual) ->
// ReLU -> zero the padding rows -> bf16 -> 16-byte coalesced stores.
#include <cstring>
#include <vector>
#include <cuda_bf16.h>

#include "az_common.cuh"
#include "az_game.cuh"
#include "az_nn.cuh"

#define TC_TILE_ROWS 128
#define TC_HALO AZ_TC_HALO
#define TC_A_ROWS (TC_TILE_ROWS + 2 * TC_HALO)              // 144
#define TC_CHUNKS 32                                         // 256 channels / 8
#ifndef TC_STAGES
#define TC_STAGES 4
#endif
#define TC_THREADS 192
#define TC_LAYER_BYTES (9 * 256 * 256 * 2)                   // 1179648 packed bf16 weights of one tower conv
#define TC_STEM_BYTES (9 * 16 * 256 * 2)                     // 73728: stem weights, 13 input channels padded to 16

// compile-time shape of one conv flavour: KCH = input channel chunks (32 for the tower, 2 for the stem)
template <int KCH> struct TcShape {
    static constexpr int CH_PER_STAGE = KCH >= 4 ? 4 : 2;            // chunks per weight stage
    static constexpr int KSTEPS = CH_PER_STAGE / 2;                   // K=16 MMAs per stage
    static constexpr int KBLOCKS = KCH / CH_PER_STAGE;
    static constexpr int ITERS = 9 * KBLOCKS;                         // weight stages per output tile
    static constexpr int STAGE_BYTES = CH_PER_STAGE * 256 * 16;
    static constexpr int A_BYTES = KCH * TC_A_ROWS * 16;              // one A buffer
    static constexpr int SMEM_BYTES = 2 * A_BYTES + TC_STAGES * STAGE_BYTES + 2 * 256 * 4 + 32 * 8 + 16;
    static constexpr int smem_bytes(int nv) { return 2 * nv * A_BYTES + TC_STAGES * STAGE_BYTES + 2 * 256 * 4 + 32 * 8 + 16; }
};

// work buffers of one forward in flight.  Two sets exist so that two forwards of the same network (the two game cohorts of a
// search, az_mcts.cu) can be in flight on two streams at once; everything else in AzTcState is read-only during a forward.
struct AzTcScratch {
    int cap_boards = 0, n_tiles = 0, r_alloc = 0;
    __nv_bfloat16* d_act[3] = { nullptr, nullptr, nullptr };   // [32][r_alloc][8]
    __nv_bfloat16* d_in = nullptr;                             // [3][2][r_alloc][8]: encoded input, 13 channels padded to 16, three copies
    size_t in_var_stride = 0;                                  // elements between the three copies of the encoded input
    float4* d_head_z = nullptr;                                // [cap_boards * 42] (pi0, pi1, v, -) per board cell, written by the last tower layer
};
#define AZ_TC_SCRATCH_SETS 2

struct AzTcState {
    int n_sm = 148;
    AzTcScratch scratch[AZ_TC_SCRATCH_SETS];
    uint8_t* d_wpacked = nullptr;                              // the 72 KB stem weight stages
    int max_pairs = 74;                                        // CTA pairs of k_nn_conv_tc3 the device holds at once
    uint8_t* d_wpacked3 = nullptr;                             // tower weights for k_nn_conv_tc3: [K group][tap][half]
    float* d_scale = nullptr; float* d_shift = nullptr;        // [2*blocks][256] folded BN, then [256] (7 used) for the stem's row BN
    float4* d_head_w = nullptr;                                // [256] (pi0, pi1, v, 0): the heads' 1x1 convolutions, per input channel
};

#include "az_tc_ptx.cuh"

// Board layout: cell (y, x) -> row y*6 + x, six zero rows behind the 42 cells (48 rows per board).  The zero rows serve as top /
// bottom border; the left / right border comes from two masked copies of the operand: taps with kx = 0 read the copy whose x = 5
// cells are zero, taps with kx = 2 the copy whose x = 0 cells are zero (what they would wrongly pick up from the neighbouring
// board row — or, for the diagonal taps at a board's first / last cell, from the neighbouring BOARD — is exactly such a cell).
#define TC_RPB AZ_TC_RPB
__device__ __forceinline__ bool tc_row_valid(int r, int n_boards) { int b = r / TC_RPB, p = r - b * TC_RPB; return b < n_boards && p < 42; }
__device__ __forceinline__ int tc_row_y(int r) { return (r % TC_RPB) / 6; }
__device__ __forceinline__ int tc_tap_shift(int tap) { return (tap / 3 - 1) * 6 + (tap % 3 - 1); }
__device__ __forceinline__ int tc_tap_variant(int tap) { return tap % 3 == 0 ? 1 : (tap % 3 == 2 ? 2 : 0); }

// ---------------------------------------------------------------- conv3x3 on tensor cores (tower and stem)
// (the stem: KCH = 2).  NV = 3 operand copies per A buffer (plain, x=5 zeroed, x=0 zeroed) read from three global arrays
// var_stride_bytes apart, written by the pack kernels; the tower builds its copies in shared memory, k_nn_conv_tc3
template <int KCH, bool ROW_BN, int NV>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_nn_conv_tc(const __nv_bfloat16* __restrict__ in, const uint8_t* __restrict__ wpacked, const float* __restrict__ scale,
             const float* __restrict__ shift, const __nv_bfloat16* __restrict__ skip, __nv_bfloat16* __restrict__ out,
             int n_boards, int r_alloc, int n_tiles, size_t var_stride_bytes)
{
    using S = TcShape<KCH>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                   // 2 x NV x A_BYTES
    uint8_t* sB = smem + 2 * NV * S::A_BYTES;             // TC_STAGES x STAGE_BYTES
    float* s_scale = reinterpret_cast<float*>(sB + TC_STAGES * S::STAGE_BYTES);
    float* s_shift = s_scale + 256;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
    uint64_t* bar_a_full = bars;                          // [2] A operand of a tile landed
    uint64_t* bar_a_empty = bars + 2;                     // [2] all MMAs reading that A buffer retired
    uint64_t* bar_w_full = bars + 4;                      // [TC_STAGES]
    uint64_t* bar_w_empty = bars + 4 + TC_STAGES;         // [TC_STAGES]
    uint64_t* bar_acc_full = bars + 4 + 2 * TC_STAGES;    // [2] accumulator complete
    uint64_t* bar_acc_empty = bars + 6 + 2 * TC_STAGES;   // [2] accumulator drained by the 4 epilogue warps
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += TC_THREADS) { s_scale[i] = scale[i]; s_shift[i] = shift[i]; }
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(bar_a_full + b, 1); mbar_init(bar_a_empty + b, 1); mbar_init(bar_acc_full + b, 1); mbar_init(bar_acc_empty + b, 4); }
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_w_full + s, 1); mbar_init(bar_w_empty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            // ---- operand streamer
            const uint8_t* src = reinterpret_cast<const uint8_t*>(in);
            auto load_a = [&](int j, int tile) {
                const int b = j & 1;
                if (j >= 2) mbar_wait(bar_a_empty + b, (uint32_t)(((j >> 1) - 1) & 1));
                mbar_expect_tx(bar_a_full + b, NV * S::A_BYTES);
                for (int v = 0; v < NV; ++v)
                    for (int c = 0; c < KCH; ++c)
                        bulk_g2s(sA + (size_t)(b * NV + v) * S::A_BYTES + (size_t)c * TC_A_ROWS * 16,
                                 src + (size_t)v * var_stride_bytes + ((size_t)c * r_alloc + (size_t)tile * TC_TILE_ROWS) * 16, TC_A_ROWS * 16, bar_a_full + b);
            };
            uint32_t wit = 0;
            int j = 0;
            if ((int)blockIdx.x < n_tiles) load_a(0, blockIdx.x);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
                for (int it = 0; it < S::ITERS; ++it, ++wit) {
                    const uint32_t s = wit % TC_STAGES, k = wit / TC_STAGES;
                    if (k > 0) mbar_wait(bar_w_empty + s, (k - 1) & 1u);
                    mbar_expect_tx(bar_w_full + s, S::STAGE_BYTES);
                    bulk_g2s(sB + (size_t)s * S::STAGE_BYTES, wpacked + (size_t)it * S::STAGE_BYTES, S::STAGE_BYTES, bar_w_full + s);
                    // prefetch the next tile's A operand early in this tile (its buffer is released by the previous tile's MMAs)
                    if (it == (S::ITERS > 8 ? 8 : S::ITERS - 1) && tile + (int)gridDim.x < n_tiles) load_a(j + 1, tile + gridDim.x);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            uint32_t wit = 0;
            int j = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
                const int b = j & 1;
                mbar_wait(bar_a_full + b, (uint32_t)((j >> 1) & 1));
                if (j >= 2) mbar_wait(bar_acc_empty + b, (uint32_t)(((j >> 1) - 1) & 1));
                tc_fence_after();
                for (int it = 0; it < S::ITERS; ++it, ++wit) {
                    const uint32_t s = wit % TC_STAGES, k = wit / TC_STAGES;
                    mbar_wait(bar_w_full + s, k & 1u);
                    tc_fence_after();
                    const int tap = it / S::KBLOCKS, kb = it - tap * S::KBLOCKS;
                    const int sh = tc_tap_shift(tap), var = NV > 1 ? tc_tap_variant(tap) : 0;
#pragma unroll
                    for (int kk = 0; kk < S::KSTEPS; ++kk) {
                        const uint64_t bdesc = umma_desc(b_base + (uint32_t)(s * S::STAGE_BYTES + kk * 2 * 256 * 16), 256 * 16, 128);
                        const int chunk0 = kb * S::CH_PER_STAGE + kk * 2;
                        const uint64_t adesc = umma_desc(a_base + (uint32_t)((b * NV + var) * S::A_BYTES + (chunk0 * TC_A_ROWS + TC_HALO + sh) * 16), TC_A_ROWS * 16, 128);
                        tc_mma_bf16(tmem_base + (uint32_t)(b * 256), adesc, bdesc, TC_IDESC, (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    tc_commit(bar_w_empty + s);          // frees the weight stage when these MMAs retire
                }
                tc_commit(bar_a_empty + b);
                tc_commit(bar_acc_full + b);
            }
        }
    } else {
        // ---- epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31
        const int q = warp & 3;
        int j = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            const int b = j & 1;
            mbar_wait(bar_acc_full + b, (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            {
            const int r = tile * TC_TILE_ROWS + q * 32 + lane;                 // padded row of this thread
            const bool valid = tc_row_valid(r, n_boards);
            const int yrow = tc_row_y(r);                                      // board row (stem BatchNorm index)
            // (ncu, round 1: with a load-then-use residual read and one TMEM load + wait per 8 columns the branch2b layers ran 285 us
            // against 220 us for the branch2a layers — the epilogue, not the MMA, set the tile time)
            const size_t cell0 = ((size_t)TC_HALO + r) * 8;
            const size_t cstride = (size_t)r_alloc * 8;
            // four channel chunks (32 accumulator columns) per TMEM load; the residual cells of the NEXT four are in flight meanwhile
            uint4 skq[4];
            if (skip) {
#pragma unroll
                for (int p = 0; p < 4; ++p) skq[p] = __ldg(reinterpret_cast<const uint4*>(skip + cell0 + (size_t)p * cstride));
            }
#pragma unroll 1
            for (int c4 = 0; c4 < TC_CHUNKS / 4; ++c4) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + c4 * 32), v);
                uint4 sk[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) sk[p] = skq[p];
                if (skip && c4 + 1 < TC_CHUNKS / 4) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) skq[p] = __ldg(reinterpret_cast<const uint4*>(skip + cell0 + (size_t)((c4 + 1) * 4 + p) * cstride));
                }
                tc_ld_wait();
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int c = c4 * 4 + p;
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int bi = ROW_BN ? (yrow < 7 ? yrow : 0) : c * 8 + e;
                        f[e] = fmaf(__uint_as_float(v[p * 8 + e]), s_scale[bi], s_shift[bi]);
                    }
                    if (skip) {
                        const __nv_bfloat162* s2 = reinterpret_cast<const __nv_bfloat162*>(&sk[p]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) { float2 x = __bfloat1622float2(s2[e]); f[2 * e] += x.x; f[2 * e + 1] += x.y; }
                    }
                    uint4 o;
                    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x0 = valid ? fmaxf(f[2 * e], 0.0f) : 0.0f, x1 = valid ? fmaxf(f[2 * e + 1], 0.0f) : 0.0f;
                        o2[e] = __floats2bfloat162_rn(x0, x1);
                    }
                    *reinterpret_cast<uint4*>(out + cell0 + (size_t)c * cstride) = o;
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + b);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// ---------------------------------------------------------------- CTA pairs (tcgen05 cta_group::2): shared constants
// ncu on the one-CTA kernel: every M128 N256 K16 MMA reads 4 KB of A and 8 KB of B from shared memory; at 128 B/clk that is
// 96 clk for an instruction the tensor pipe finishes in 64, and the kernel ran exactly at that bound (65 % pipe active).
// On a CTA pair (two SMs of a TPC, M = 256) each CTA keeps its own 128 rows of A and only HALF of the weights (128 output
// channels, 8 KB per K=32 stage); the pair's tensor cores exchange the B halves, so every SM reads 8 KB per MMA and streams half
// the weight bytes from L2.
#define TC2_STAGES 8
#define TC2_STAGE_BYTES (4 * 128 * 16)                       // K = 32 (4 chunks) x 128 output channels

// ---------------------------------------------------------------- tower conv3x3 on a CTA pair, 48-row board layout
// tcgen05 cta_group::2, half of the weights per CTA (above), on the 48-row layout: the A operand
// exists in three copies in shared memory (plain / x=5 cells zeroed / x=0 cells zeroed, see tc_tap_variant), so it no longer fits
// twice.  The K loop is therefore turned inside out — K group (32 channels) outer, the nine taps inner — and the operand lives in a
// RING of K groups: a group's slot is refilled for the next tile as soon as its nine taps have retired.
//   warp 0      : streams the plain copy of K groups T3_LA groups ahead (4 x 2.3 KB bulk copies) and the weight half-stages
//   warp 6      : "masker" — when a group has landed, writes its two masked copies with ordinary shared-memory stores, makes them
//                 visible to the tensor core (fence.proxy.async) and signals this CTA's and the leader's "group ready" barrier
//   warp 1      : leader: MMA issue (kb outer, tap inner); peer: relays "weight half landed"
//   warps 2..5  : epilogue, as before
#define T3_THREADS 224
#define T3_G 5                                               // ring slots
#define T3_LA 3                                              // operand prefetch distance in K groups (<= T3_G - 2, see the streamer)
#define T3_VAR_BYTES (4 * TC_A_ROWS * 16)                    // one copy of one K group: 4 chunks x 144 rows x 16 B = 9216
#define T3_SLOT_BYTES (3 * T3_VAR_BYTES)
#define T3_SMEM_BYTES (T3_G * T3_SLOT_BYTES + TC2_STAGES * TC2_STAGE_BYTES + 2 * 256 * 4 + 64 * 8 + 16)
#define T3_HEAD_BYTES (256 * 16)                             // T3_HEADS: (pi0, pi1, v, -) head weights per input channel

// MODE = T3_RAW (training step, az_tc_conv_raw): the epilogue stores the fp32 accumulators of the real board cells to out32
// [board * 42 + cell][256] row-major — no BatchNorm fold, residual, ReLU or bf16 rounding; scale / shift / skip / out are unused.
// MODE = T3_HEADS (the LAST tower layer of a forward): the activation is not stored at all.  Each epilogue thread holds one board
// cell's 256 output channels anyway, so it applies the two heads' 1x1 convolutions (pi: 2 channels, v: 1 channel,
// build_graph.py:76-90) on the spot — on the bf16-rounded activation, in the channel order the separate head kernel used, so the
// sums are bit-identical to reading the stored activation back — and writes (pi0, pi1, v) per cell to out32 (as float4
// [board * 42 + cell]).  Saves the layer's 57 MB store and the head kernel's 103 MB read per 4096 boards.
#define T3_ACT 0
#define T3_RAW 1
#define T3_HEADS 2
template <int MODE>
__global__ void __launch_bounds__(T3_THREADS, 1)
k_nn_conv_tc3(const __nv_bfloat16* __restrict__ in, const uint8_t* __restrict__ wpacked3, const float* __restrict__ scale,
              const float* __restrict__ shift, const __nv_bfloat16* __restrict__ skip, __nv_bfloat16* __restrict__ out,
              int n_boards, int r_alloc, int n_tiles, float* __restrict__ out32, const float4* __restrict__ head_w)
{
    constexpr bool RAW = MODE == T3_RAW;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                   // T3_G x T3_SLOT_BYTES
    uint8_t* sB = smem + T3_G * T3_SLOT_BYTES;            // TC2_STAGES x TC2_STAGE_BYTES
    float* s_scale = reinterpret_cast<float*>(sB + TC2_STAGES * TC2_STAGE_BYTES);
    float* s_shift = s_scale + 256;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
    uint64_t* bar_a_full = bars;                          // [T3_G] plain copy of the group landed
    uint64_t* bar_a_ready = bars + 8;                     // [T3_G] masked copies written (this CTA)
    uint64_t* bar_pa_ready = bars + 16;                   // [T3_G] leader only: the peer's group is ready
    uint64_t* bar_a_empty = bars + 24;                    // [T3_G] the group's MMAs retired (leader commit, both CTAs)
    uint64_t* bar_acc_full = bars + 32;                   // [2]
    uint64_t* bar_acc_empty = bars + 34;                  // [2] leader only, 8 epilogue warps of the pair
    uint64_t* bar_w_full = bars + 36;                     // [TC2_STAGES]
    uint64_t* bar_w_empty = bars + 36 + TC2_STAGES;
    uint64_t* bar_pw_full = bars + 36 + 2 * TC2_STAGES;   // leader only
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 64);
    float4* s_hw = reinterpret_cast<float4*>(smem + T3_SMEM_BYTES);      // T3_HEADS only (the launch adds T3_HEAD_BYTES)

    // let the next tower layer (launched with the programmatic-serialization attribute) be scheduled onto SMs as this grid's
    // CTAs exit: its barrier / TMEM set-up then overlaps this layer's last wave; it reads nothing of ours before its pdl_wait()
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const bool leader = crank == 0;
    const int n_pairs = (int)(gridDim.x >> 1), pid = (int)(blockIdx.x >> 1);
    const int n_items = (n_tiles + 1) >> 1;
    const int my_items = pid < n_items ? (n_items - pid + n_pairs - 1) / n_pairs : 0;
    const int total_groups = my_items * 8;
    auto tile_of_group = [&](int q) {                     // the 128-row tile this CTA works on while consuming K group q
        int tile = 2 * (pid + (q >> 3) * n_pairs) + (int)crank;
        return tile < n_tiles ? tile : n_tiles - 1;       // odd tile count: the peer's last tile is multiplied, never stored
    };

    if (!RAW) for (int i = threadIdx.x; i < 256; i += T3_THREADS) { s_scale[i] = scale[i]; s_shift[i] = shift[i]; }
    if (MODE == T3_HEADS) for (int i = threadIdx.x; i < 256; i += T3_THREADS) s_hw[i] = head_w[i];
    if (threadIdx.x == 0) {
        for (int g = 0; g < T3_G; ++g) { mbar_init(bar_a_full + g, 1); mbar_init(bar_a_ready + g, 1); mbar_init(bar_pa_ready + g, 1); mbar_init(bar_a_empty + g, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full + b, 1); mbar_init(bar_acc_empty + b, 8); }
        for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(bar_w_full + s, 1); mbar_init(bar_w_empty + s, 1); mbar_init(bar_pw_full + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();                                           // the previous kernel's activations are complete and visible from here on

    if (warp == 0) {
        if (lane == 0) {
            // ---- operand streamer.  Group q + T3_LA is requested while group q's weights are being issued; its slot was last used
            // by group q + T3_LA - T3_G <= q - 2, whose MMAs have retired by then (the weight ring keeps this thread less than
            // one group ahead of the MMA issuer), so the wait below never stalls the weight stream.
            const uint8_t* src = reinterpret_cast<const uint8_t*>(in);
            auto load_group = [&](int q) {
                const int slot = q % T3_G, rnd = q / T3_G, kb = q & 7, tile = tile_of_group(q);
                if (rnd > 0) mbar_wait_cluster(bar_a_empty + slot, (uint32_t)((rnd - 1) & 1));
                mbar_expect_tx(bar_a_full + slot, T3_VAR_BYTES);
                for (int c = 0; c < 4; ++c)
                    bulk_g2s(sA + (size_t)slot * T3_SLOT_BYTES + (size_t)c * TC_A_ROWS * 16,
                             src + ((size_t)(kb * 4 + c) * r_alloc + (size_t)tile * TC_TILE_ROWS) * 16, TC_A_ROWS * 16, bar_a_full + slot);
            };
            for (int q = 0; q < T3_LA && q < total_groups; ++q) load_group(q);
            uint32_t wit = 0;
            for (int q = 0; q < total_groups; ++q) {
                if (q + T3_LA < total_groups) load_group(q + T3_LA);
                const int kb = q & 7;
                for (int tap = 0; tap < 9; ++tap, ++wit) {
                    const uint32_t s = wit % TC2_STAGES, k = wit / TC2_STAGES;
                    if (k > 0) mbar_wait_cluster(bar_w_empty + s, (k - 1) & 1u);
                    mbar_expect_tx(bar_w_full + s, TC2_STAGE_BYTES);
                    bulk_g2s(sB + (size_t)s * TC2_STAGE_BYTES, wpacked3 + ((size_t)(kb * 9 + tap) * 2 + crank) * TC2_STAGE_BYTES, TC2_STAGE_BYTES, bar_w_full + s);
                }
            }
        }
    } else if (warp == 6) {
        // ---- masker: A_L (x = 5 cells zero) and A_R (x = 0 cells zero) of every landed group
        for (int q = 0; q < total_groups; ++q) {
            const int slot = q % T3_G, rnd = q / T3_G;
            const int row0 = tile_of_group(q) * TC_TILE_ROWS - TC_HALO;          // board-layout row of the slot's first operand row (may be < 0: halo of tile 0)
            mbar_wait(bar_a_full + slot, (uint32_t)(rnd & 1));
            uint8_t* base = sA + (size_t)slot * T3_SLOT_BYTES;
            for (int idx = lane; idx < 4 * TC_A_ROWS; idx += 32) {
                const int row = idx % TC_A_ROWS;
                const int p = (row0 + row + TC_RPB) % TC_RPB;                     // row0 + row >= -8
                const int x = p < 42 ? p % 6 : -1;
                const uint4 cell = *reinterpret_cast<const uint4*>(base + (size_t)idx * 16);
                const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(base + T3_VAR_BYTES + (size_t)idx * 16) = x == 5 ? zero : cell;
                *reinterpret_cast<uint4*>(base + 2 * T3_VAR_BYTES + (size_t)idx * 16) = x == 0 ? zero : cell;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_a_ready + slot); if (!leader) mbar_arrive_remote(bar_pa_ready + slot, 0u); }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ---- MMA issuer (leader CTA): K group outer, taps inner
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            uint32_t wit = 0;
            for (int j = 0; j < my_items; ++j) {
                const int b = j & 1;
                if (j >= 2) mbar_wait_cluster(bar_acc_empty + b, (uint32_t)(((j >> 1) - 1) & 1));
                for (int kb = 0; kb < 8; ++kb) {
                    const int q = j * 8 + kb, slot = q % T3_G, rnd = q / T3_G;
                    mbar_wait(bar_a_ready + slot, (uint32_t)(rnd & 1));
                    mbar_wait_cluster(bar_pa_ready + slot, (uint32_t)(rnd & 1));
                    tc_fence_after();
                    for (int tap = 0; tap < 9; ++tap, ++wit) {
                        const uint32_t s = wit % TC2_STAGES, k = wit / TC2_STAGES;
                        mbar_wait(bar_w_full + s, k & 1u);
                        mbar_wait_cluster(bar_pw_full + s, k & 1u);
                        tc_fence_after();
                        const int sh = tc_tap_shift(tap), var = tc_tap_variant(tap);
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint64_t bdesc = umma_desc(b_base + (uint32_t)(s * TC2_STAGE_BYTES + kk * 2 * 128 * 16), 128 * 16, 128);
                            const uint64_t adesc = umma_desc(a_base + (uint32_t)(slot * T3_SLOT_BYTES + var * T3_VAR_BYTES + (kk * 2 * TC_A_ROWS + TC_HALO + sh) * 16), TC_A_ROWS * 16, 128);
                            tc2_mma_bf16(tmem_base + (uint32_t)(b * 256), adesc, bdesc, TC2_IDESC, (kb > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                        }
                        tc2_commit(bar_w_empty + s);
                    }
                    tc2_commit(bar_a_empty + slot);
                }
                tc2_commit(bar_acc_full + b);
            }
        } else if (lane == 0) {
            // ---- relay (peer CTA): this CTA's weight halves have landed
            const uint32_t total_stages = (uint32_t)total_groups * 9u;
            for (uint32_t wit = 0; wit < total_stages; ++wit) {
                const uint32_t s = wit % TC2_STAGES, k = wit / TC2_STAGES;
                mbar_wait(bar_w_full + s, k & 1u);
                mbar_arrive_remote(bar_pw_full + s, 0u);
            }
        }
    } else {
        // ---- epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 of this CTA's half of the pair's accumulator
        const int q4 = warp & 3;
        for (int j = 0; j < my_items; ++j) {
            const int b = j & 1;
            const int tile = 2 * (pid + j * n_pairs) + (int)crank;
            mbar_wait_cluster(bar_acc_full + b, (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (RAW) {
                if (tile < n_tiles) {
                    const int r = tile * TC_TILE_ROWS + q4 * 32 + lane;
                    const bool valid = tc_row_valid(r, n_boards);
                    const int bb = r / TC_RPB;
                    float* orow = out32 + ((size_t)bb * 42 + (size_t)(r - bb * TC_RPB)) * 256;
#pragma unroll 1
                    for (int c4 = 0; c4 < TC_CHUNKS / 4; ++c4) {
                        uint32_t v[32];
                        tc_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(b * 256 + c4 * 32), v);
                        tc_ld_wait();
                        if (valid) {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                *reinterpret_cast<float4*>(orow + c4 * 32 + e * 4) = make_float4(__uint_as_float(v[e * 4]), __uint_as_float(v[e * 4 + 1]),
                                                                                               __uint_as_float(v[e * 4 + 2]), __uint_as_float(v[e * 4 + 3]));
                        }
                    }
                }
            } else if (tile < n_tiles) {
                const int r = tile * TC_TILE_ROWS + q4 * 32 + lane;
                const bool valid = tc_row_valid(r, n_boards);
                const size_t cell0 = ((size_t)TC_HALO + r) * 8;
                const size_t cstride = (size_t)r_alloc * 8;
                float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f;             // T3_HEADS: the cell's three head sums
                // four channel chunks (32 accumulator columns) per TMEM load; the residual cells of the NEXT four are in flight meanwhile
                uint4 skq[4];
                if (skip) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) skq[p] = __ldg(reinterpret_cast<const uint4*>(skip + cell0 + (size_t)p * cstride));
                }
#pragma unroll 1
                for (int c4 = 0; c4 < TC_CHUNKS / 4; ++c4) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(b * 256 + c4 * 32), v);
                    uint4 sk[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) sk[p] = skq[p];
                    if (skip && c4 + 1 < TC_CHUNKS / 4) {
#pragma unroll
                        for (int p = 0; p < 4; ++p) skq[p] = __ldg(reinterpret_cast<const uint4*>(skip + cell0 + (size_t)((c4 + 1) * 4 + p) * cstride));
                    }
                    tc_ld_wait();
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int c = c4 * 4 + p;
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = fmaf(__uint_as_float(v[p * 8 + e]), s_scale[c * 8 + e], s_shift[c * 8 + e]);
                        if (skip) {
                            const __nv_bfloat162* s2 = reinterpret_cast<const __nv_bfloat162*>(&sk[p]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) { float2 x = __bfloat1622float2(s2[e]); f[2 * e] += x.x; f[2 * e + 1] += x.y; }
                        }
                        uint4 o;
                        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float x0 = valid ? fmaxf(f[2 * e], 0.0f) : 0.0f, x1 = valid ? fmaxf(f[2 * e + 1], 0.0f) : 0.0f;
                            o2[e] = __floats2bfloat162_rn(x0, x1);
                            if (MODE == T3_HEADS) {                // same values, same order as a head kernel reading the stored bf16 cell
                                const float2 xv = __bfloat1622float2(o2[e]);
                                const float4 wa = s_hw[c * 8 + 2 * e], wb = s_hw[c * 8 + 2 * e + 1];
                                h0 = fmaf(xv.x, wa.x, h0); h1 = fmaf(xv.x, wa.y, h1); h2 = fmaf(xv.x, wa.z, h2);
                                h0 = fmaf(xv.y, wb.x, h0); h1 = fmaf(xv.y, wb.y, h1); h2 = fmaf(xv.y, wb.z, h2);
                            }
                        }
                        if (MODE != T3_HEADS) *reinterpret_cast<uint4*>(out + cell0 + (size_t)c * cstride) = o;
                    }
                }
                if (MODE == T3_HEADS && valid) {
                    const int bb = r / TC_RPB;
                    reinterpret_cast<float4*>(out32)[(size_t)bb * 42 + (size_t)(r - bb * TC_RPB)] = make_float4(h0, h1, h2, 0.0f);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(bar_acc_empty + b); else mbar_arrive_remote(bar_acc_empty + b, 0u);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------- input packing + heads on the padded bf16 layout
// fp32 [n][42][13] -> bf16 [2 chunks][r_alloc][8] (channels 13..15 = 0); thread = board cell
// (three copies var_stride elements apart — plain, x = 5 cells zeroed, x = 0 cells zeroed — for the stem's taps)
__global__ void __launch_bounds__(256) k_nn_pack_input_tc(const float* __restrict__ x, int n, __nv_bfloat16* __restrict__ out, int r_alloc,
                                                           size_t var_stride)
{
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n * 42) return;
    int b = i / 42, p = i - b * 42;
    const float* src = x + (size_t)i * AZ_NN_IN_CH;
    int row = b * TC_RPB + p;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint4 o;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int c0 = c * 8 + 2 * e;
            float a = c0 < AZ_NN_IN_CH ? src[c0] : 0.0f, bb = c0 + 1 < AZ_NN_IN_CH ? src[c0 + 1] : 0.0f;
            o2[e] = __floats2bfloat162_rn(a, bb);
        }
        const size_t at = ((size_t)c * r_alloc + TC_HALO + row) * 8;
        *reinterpret_cast<uint4*>(out + at) = o;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(out + var_stride + at) = (p % 6) == 5 ? zero : o;
        *reinterpret_cast<uint4*>(out + 2 * var_stride + at) = (p % 6) == 0 ? zero : o;
    }
}

// The same packed input straight from the game states (the leaf batch of a search): NNInputData + setInStateTensor
// (alphazero_nn_data.cpp:165-196, alphazero_nn.cpp:31-67) as k_env_encode computes them (az_env.cu — same expressions, same
// roundings), rounded to bf16 and written in the layout above without the fp32 [n][42][13] round trip through HBM.
// One warp per position; lane l packs board cells l and l + 32.
// (st = SoA game states, word w of position i at st[w * st_stride + i]; st_stride >= n lets a caller evaluate the first n of more slots)
__global__ void __launch_bounds__(128) k_nn_pack_state_tc(const uint32_t* __restrict__ st, int n, int st_stride, __nv_bfloat16* __restrict__ out,
                                                           int r_alloc, size_t var_stride)
{
    __shared__ uint32_t s_words[4][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * 4 + warp;
    if (gi >= n) return;
    if (lane < 14) s_words[warp][lane] = st[(size_t)lane * st_stride + gi];
    __syncwarp();
    const uint8_t* land = (const uint8_t*)s_words[warp];
    AzGame g;
    az_unpack_scalars(g, s_words[warp][10], s_words[warp][11], s_words[warp][12], s_words[warp][13]);
    const uint32_t b0 = land[lane], b1 = lane < 10 ? land[32 + lane] : (3u << 6);
    const uint32_t m0a = __ballot_sync(0xffffffffu, (b0 >> 6) == 0), m0b = __ballot_sync(0xffffffffu, (b1 >> 6) == 0);
    const uint32_t m1a = __ballot_sync(0xffffffffu, (b0 >> 6) == 1), m1b = __ballot_sync(0xffffffffu, (b1 >> 6) == 1);
    g.own0 = (uint64_t)m0a | ((uint64_t)m0b << 32); g.own1 = (uint64_t)m1a | ((uint64_t)m1b << 32);
    int t0 = ((b0 >> 6) == 0 ? (int)(b0 & 63u) : 0) + ((b1 >> 6) == 0 ? (int)(b1 & 63u) : 0);
    int t1 = ((b0 >> 6) == 1 ? (int)(b0 & 63u) : 0) + ((b1 >> 6) == 1 ? (int)(b1 & 63u) : 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); }
    const uint32_t cur = g.cur, enemy = cur ^ 1u;
    const float ref = (float)az_reinforcement_value(g.own(cur)), eref = (float)az_reinforcement_value(g.own(enemy));
    const float reinf_share = __fdiv_rn(ref, __fadd_rn(ref, eref));
    float att = __fdiv_rn((float)g.attacks, 8.0f); att = att < 1.0f ? att : 1.0f;
    const float ta = (float)(cur ? t1 : t0), eta = (float)(cur ? t0 : t1);
    const float army_share = __fdiv_rn(ta, __fadd_rn(ta, eta));
    // channels 3..12 are the same for every cell of the board: chunk 0 = {own, enemy, neutral, 3, 4, 5, 6, 7}, chunk 1 = {8..12, 0, 0, 0}
    float sc[10];
    sc[0] = army_share; sc[1] = reinf_share; sc[2] = att; sc[3] = g.allow_draw ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < 6; ++k) sc[4 + k] = g.phase == (uint32_t)k ? 1.0f : 0.0f;
    uint4 c1;
    {
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&c1);
        o2[0] = __floats2bfloat162_rn(sc[5], sc[6]); o2[1] = __floats2bfloat162_rn(sc[7], sc[8]);
        o2[2] = __floats2bfloat162_rn(sc[9], 0.0f); o2[3] = __floats2bfloat162_rn(0.0f, 0.0f);
    }
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int p = lane + 32 * h;
        if (p >= 42) break;
        const uint32_t v = h ? b1 : b0, o = v >> 6;
        const float fa = __fdiv_rn((float)(v & 63u), 32.0f);
        uint4 c0;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&c0);
        o2[0] = __floats2bfloat162_rn(o == cur ? fa : 0.0f, o == enemy ? fa : 0.0f);
        o2[1] = __floats2bfloat162_rn(o == AZ_NEUTRAL ? fa : 0.0f, sc[0]);
        o2[2] = __floats2bfloat162_rn(sc[1], sc[2]); o2[3] = __floats2bfloat162_rn(sc[3], sc[4]);
        const int row = gi * TC_RPB + p;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint4 val = c ? c1 : c0;
            const size_t at = ((size_t)c * r_alloc + TC_HALO + row) * 8;
            *reinterpret_cast<uint4*>(out + at) = val;
            *reinterpret_cast<uint4*>(out + var_stride + at) = (p % 6) == 5 ? zero : val;
            *reinterpret_cast<uint4*>(out + 2 * var_stride + at) = (p % 6) == 0 ? zero : val;
        }
    }
}

__device__ __forceinline__ float warp_sum_tc(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// policy + value heads for HB boards per block, from the per-cell 1x1-convolution sums the last tower layer's epilogue wrote
// (k_nn_conv_tc3<T3_HEADS>): BatchNorm + ReLU per cell, the dense layers, softmax / tanh (build_graph.py:76-90)
#define HB 6     // 6 boards = 252 cells: one pass of the 256 threads
__global__ void __launch_bounds__(256) k_nn_heads_tail(const float4* __restrict__ head_z, int n, AzHeadParams hp,
                                                        float* __restrict__ policy, float* __restrict__ value)
{
    __shared__ float s_pi[HB][84], s_v[HB][42], s_logit[HB][44], s_red[HB][8];
    const int b0 = blockIdx.x * HB, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = n - b0 < HB ? n - b0 : HB;
    const float sc0 = hp.bn_pi[0] * rsqrtf(hp.bn_pi[6] + AZ_NN_BN_EPS), sc1 = hp.bn_pi[1] * rsqrtf(hp.bn_pi[7] + AZ_NN_BN_EPS);
    const float scv = hp.bn_v[0] * rsqrtf(hp.bn_v[3] + AZ_NN_BN_EPS);
    for (int i = threadIdx.x; i < nb * 42; i += 256) {
        const int bl = i / 42, p = i - bl * 42;
        const float4 z = __ldg(head_z + (size_t)(b0 + bl) * 42 + p);
        s_pi[bl][p * 2 + 0] = fmaxf((z.x - hp.bn_pi[4]) * sc0 + hp.bn_pi[2], 0.0f);
        s_pi[bl][p * 2 + 1] = fmaxf((z.y - hp.bn_pi[5]) * sc1 + hp.bn_pi[3], 0.0f);
        s_v[bl][p] = fmaxf((z.z - hp.bn_v[2]) * scv + hp.bn_v[1], 0.0f);
    }
    __syncthreads();
    // dense 84 -> 43 (policy logits): thread = output, weights reused over the block's boards
    if (threadIdx.x < 43) {
        float acc[HB];
#pragma unroll
        for (int bl = 0; bl < HB; ++bl) acc[bl] = hp.dense_b[threadIdx.x];
        for (int k = 0; k < 84; ++k) {
            const float w = hp.dense_w[k * 43 + threadIdx.x];
#pragma unroll
            for (int bl = 0; bl < HB; ++bl) acc[bl] = fmaf(s_pi[bl][k], w, acc[bl]);
        }
#pragma unroll
        for (int bl = 0; bl < HB; ++bl) s_logit[bl][threadIdx.x] = acc[bl];
    }
    // dense 42 -> 256 + ReLU, then 256 -> 1 (value)
    {
        float acc[HB];
#pragma unroll
        for (int bl = 0; bl < HB; ++bl) acc[bl] = hp.dense1_b[threadIdx.x];
        for (int k = 0; k < 42; ++k) {
            const float w = hp.dense1_w[k * 256 + threadIdx.x];
#pragma unroll
            for (int bl = 0; bl < HB; ++bl) acc[bl] = fmaf(s_v[bl][k], w, acc[bl]);
        }
        const float w2 = hp.dense2_w[threadIdx.x];
#pragma unroll
        for (int bl = 0; bl < HB; ++bl) {
            const float part = warp_sum_tc(fmaxf(acc[bl], 0.0f) * w2);
            if (lane == 0) s_red[bl][warp] = part;
        }
    }
    __syncthreads();
    if (threadIdx.x < nb) {
        float sum = hp.dense2_b[0];
        for (int i = 0; i < 8; ++i) sum += s_red[threadIdx.x][i];
        value[b0 + threadIdx.x] = tanhf(sum);
    }
    // softmax: one warp per board
    for (int bl = warp; bl < nb; bl += 8) {
        const float l0 = s_logit[bl][lane], l1 = lane < 11 ? s_logit[bl][32 + lane] : -INFINITY;
        float mx = fmaxf(l0, l1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e0 = expf(l0 - mx), e1 = lane < 11 ? expf(l1 - mx) : 0.0f;
        const float sum = warp_sum_tc(e0 + e1);
        policy[(size_t)(b0 + bl) * 43 + lane] = e0 / sum;
        if (lane < 11) policy[(size_t)(b0 + bl) * 43 + 32 + lane] = e1 / sum;
    }
}

// ---------------------------------------------------------------- host side

static std::string tc_block_name(int i) { return std::to_string(i) + std::string(1, (char)('a' + i)); }

// HWIO fp32 [3][3][cin][256] -> bf16 [tap][K block][chunk][256 out][8], cin padded to a multiple of 16
static void pack_conv(const float* w, int cin, int kch, __nv_bfloat16* dst)
{
    const int ch_per_stage = kch >= 4 ? 4 : 2, kblocks = kch / ch_per_stage;
    for (int tap = 0; tap < 9; ++tap)
        for (int kb = 0; kb < kblocks; ++kb)
            for (int ch = 0; ch < ch_per_stage; ++ch)
                for (int n = 0; n < 256; ++n)
                    for (int e = 0; e < 8; ++e) {
                        int ci = (kb * ch_per_stage + ch) * 8 + e;
                        float v = ci < cin ? w[((size_t)tap * cin + ci) * 256 + n] : 0.0f;
                        dst[((((size_t)tap * kblocks + kb) * ch_per_stage + ch) * 256 + n) * 8 + e] = __float2bfloat16_rn(v);
                    }
}

// HWIO fp32 [3][3][256][256] -> bf16 [tap][K block of 32][half of the output channels][chunk 4][128 out][8]: one 8 KB stage per CTA of a pair
// (kb_outer: stages ordered [K block][tap] for k_nn_conv_tc3 instead of [tap][K block])
static void pack_conv_pair(const float* w, __nv_bfloat16* dst, bool kb_outer = true)
{
    for (int tap = 0; tap < 9; ++tap)
        for (int kb = 0; kb < 8; ++kb)
            for (int h = 0; h < 2; ++h)
                for (int ch = 0; ch < 4; ++ch)
                    for (int n = 0; n < 128; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int ci = (kb * 4 + ch) * 8 + e, co = h * 128 + n;
                            const size_t stage = kb_outer ? (size_t)kb * 9 + tap : (size_t)tap * 8 + kb;
                            dst[((((stage * 2 + h) * 4 + ch) * 128 + n) * 8) + e] = __float2bfloat16_rn(w[((size_t)tap * 256 + ci) * 256 + co]);
                        }
}

// heads != NULL: the last layer — no activation out, per-cell head sums to head_z instead (T3_HEADS)
static cudaError_t launch_conv_pair3(int grid, cudaStream_t s, const __nv_bfloat16* in, const uint8_t* w3, const float* scale, const float* shift,
                                     const __nv_bfloat16* skip, __nv_bfloat16* out, int n_boards, int r_alloc, int n_tiles,
                                     const float4* head_w = nullptr, float4* head_z = nullptr, bool pdl = true)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(T3_THREADS); cfg.stream = s;
    cfg.dynamicSmemBytes = T3_SMEM_BYTES + (head_w ? T3_HEAD_BYTES : 0);
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    // programmatic dependent launch between consecutive layers: the next layer's barrier / TMEM set-up overlaps this layer's last wave
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    // (off when two forwards share the device: a successor parked on an SM pair in griddepcontrol.wait would keep the other
    // forward's layer from filling the pairs this layer's last wave leaves idle)
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
    if (head_w)
        return cudaLaunchKernelEx(&cfg, k_nn_conv_tc3<T3_HEADS>, in, w3, scale, shift, skip, out, n_boards, r_alloc, n_tiles,
                                  reinterpret_cast<float*>(head_z), head_w);
    return cudaLaunchKernelEx(&cfg, k_nn_conv_tc3<T3_ACT>, in, w3, scale, shift, skip, out, n_boards, r_alloc, n_tiles, (float*)nullptr,
                              (const float4*)nullptr);
}

// persistent launch of the stem convolution
template <int KCH, bool ROW_BN, int NV>
static cudaError_t launch_conv(int grid, cudaStream_t s, const __nv_bfloat16* in, const uint8_t* w, const float* scale, const float* shift,
                               const __nv_bfloat16* skip, __nv_bfloat16* out, int n_boards, int r_alloc, int n_tiles, size_t var_stride_bytes)
{
    k_nn_conv_tc<KCH, ROW_BN, NV><<<grid, TC_THREADS, TcShape<KCH>::smem_bytes(NV), s>>>(in, w, scale, shift, skip, out, n_boards, r_alloc, n_tiles,
                                                                                        var_stride_bytes);
    return cudaGetLastError();
}

int az_nn_tc_prepare(az_nn* nn)
{
    if (!nn->tc) nn->tc = new AzTcState();
    AzTcState* tc = nn->tc;
    const int layers = 2 * nn->blocks;
    std::vector<uint8_t> packed((size_t)TC_STEM_BYTES);                   // stem stages
    std::vector<uint8_t> packed3((size_t)layers * TC_LAYER_BYTES);         // tower stages, one half per CTA of a pair
    std::vector<float> scale((size_t)(layers + 1) * 256, 0.0f), shift((size_t)(layers + 1) * 256, 0.0f);
    for (int L = 0; L < layers; ++L) {
        std::string sfx = tc_block_name(L / 2) + ((L & 1) ? "_branch2b" : "_branch2a");
        const float* w = az_nn_host_var(nn, "res" + sfx + "/kernel");           // HWIO [3][3][256][256]
        const float* g = az_nn_host_var(nn, "bn" + sfx + "/gamma");
        const float* be = az_nn_host_var(nn, "bn" + sfx + "/beta");
        const float* mu = az_nn_host_var(nn, "bn" + sfx + "/moving_mean");
        const float* var = az_nn_host_var(nn, "bn" + sfx + "/moving_variance");
        if (!w || !g || !be || !mu || !var) { az_set_error("missing tower variable for layer %d", L); return AZ_ERR_INVALID_ARG; }
        for (int c = 0; c < 256; ++c) {
            float sc = g[c] / sqrtf(var[c] + AZ_NN_BN_EPS);
            scale[(size_t)L * 256 + c] = sc; shift[(size_t)L * 256 + c] = be[c] - mu[c] * sc;
        }
        pack_conv_pair(w, reinterpret_cast<__nv_bfloat16*>(packed3.data() + (size_t)L * TC_LAYER_BYTES));
    }
    {   // stem: conv/kernel [3][3][13][256], BatchNorm over the 7 board rows
        const float* w = az_nn_host_var(nn, "conv/kernel");
        const float* g = az_nn_host_var(nn, "conv_bn/gamma");
        const float* be = az_nn_host_var(nn, "conv_bn/beta");
        const float* mu = az_nn_host_var(nn, "conv_bn/moving_mean");
        const float* var = az_nn_host_var(nn, "conv_bn/moving_variance");
        if (!w || !g || !be || !mu || !var) { az_set_error("missing stem variable"); return AZ_ERR_INVALID_ARG; }
        for (int y = 0; y < 7; ++y) {
            float sc = g[y] / sqrtf(var[y] + AZ_NN_BN_EPS);
            scale[(size_t)layers * 256 + y] = sc; shift[(size_t)layers * 256 + y] = be[y] - mu[y] * sc;
        }
        pack_conv(w, AZ_NN_IN_CH, 2, reinterpret_cast<__nv_bfloat16*>(packed.data()));
    }
    if (!tc->d_wpacked) {
        AZ_CUDA(cudaMalloc(&tc->d_wpacked, packed.size()));
        AZ_CUDA(cudaMalloc(&tc->d_wpacked3, packed3.size()));
        AZ_CUDA(cudaMalloc(&tc->d_scale, scale.size() * sizeof(float)));
        AZ_CUDA(cudaMalloc(&tc->d_shift, shift.size() * sizeof(float)));
        AZ_CUDA((cudaFuncSetAttribute(k_nn_conv_tc<2, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcShape<2>::smem_bytes(3))));
        AZ_CUDA(cudaFuncSetAttribute(k_nn_conv_tc3<T3_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES));
        AZ_CUDA(cudaFuncSetAttribute(k_nn_conv_tc3<T3_HEADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES + T3_HEAD_BYTES));
        AZ_CUDA(cudaMalloc(&tc->d_head_w, 256 * sizeof(float4)));
        int dev = 0, sms = 148;
        AZ_CUDA(cudaGetDevice(&dev));
        AZ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        tc->n_sm = sms;
        {   // CTA pairs the device can hold at once
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(sms / 2 * 2)); cfg.blockDim = dim3(T3_THREADS); cfg.dynamicSmemBytes = T3_SMEM_BYTES + T3_HEAD_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nc = 0;
            AZ_CUDA(cudaOccupancyMaxActiveClusters(&nc, k_nn_conv_tc3<T3_HEADS>, &cfg));
            tc->max_pairs = nc < sms / 2 ? nc : sms / 2;
            if (tc->max_pairs < 1) { az_set_error("the device cannot hold one CTA pair of the tower kernel (sm_100a thread-block clusters needed)"); return AZ_ERR_CUDA; }
        }
    }
    AZ_CUDA(cudaMemcpy(tc->d_wpacked, packed.data(), packed.size(), cudaMemcpyHostToDevice));
    AZ_CUDA(cudaMemcpy(tc->d_wpacked3, packed3.data(), packed3.size(), cudaMemcpyHostToDevice));
    AZ_CUDA(cudaMemcpy(tc->d_scale, scale.data(), scale.size() * sizeof(float), cudaMemcpyHostToDevice));
    AZ_CUDA(cudaMemcpy(tc->d_shift, shift.data(), shift.size() * sizeof(float), cudaMemcpyHostToDevice));
    {   // the heads' 1x1 convolutions: pi/kernel [1][1][256][2], v/kernel [1][1][256][1] -> (pi0, pi1, v, 0) per input channel
        const float* pw = az_nn_host_var(nn, "pi/kernel");
        const float* vw = az_nn_host_var(nn, "v/kernel");
        if (!pw || !vw) { az_set_error("missing head variable"); return AZ_ERR_INVALID_ARG; }
        std::vector<float4> hw(256);
        for (int c = 0; c < 256; ++c) hw[(size_t)c] = make_float4(pw[c * 2], pw[c * 2 + 1], vw[c], 0.0f);
        AZ_CUDA(cudaMemcpy(tc->d_head_w, hw.data(), hw.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    return AZ_OK;
}

// ---------------------------------------------------------------- raw 3x3 convolution for the training step (az_nn_train.cu)
// out[r][co] = sum_{t,ci} in[nb(r,t)][ci] * w[t][ci][co]           (flip = 0: forward)
// out[r][ci] = sum_{t,co} in[nb(r,t)][co] * w[8-t][ci][co]         (flip = 1: data gradient)
// on the tower kernel: bf16 operands, fp32 accumulation, fp32 result [n * 42][256].  The fp32 source goes into the 48-row chunked
// bf16 layout, the fp32 HWIO weights (which change every step) are packed into the kernel's stage order on the device.

// fp32 [n * 42][256] -> bf16 [32 chunks][r_alloc][8] at rows TC_HALO + board * TC_RPB + cell; padding rows are never written (zero from the
// allocation).  Block = 32 source rows x 32 chunks through shared memory: contiguous reads along a row, contiguous writes along a chunk.
// variants = 3: two more copies var_stride_u4 apart — cells with x = 0 zeroed, cells with x = 5 zeroed (the weight gradient masks the
// OUTPUT cell of a tap that would reach over the left / right board border).
__global__ void __launch_bounds__(256) k_tc_chunk_rpb(const float* __restrict__ src, int rows, int r_alloc, __nv_bfloat16* __restrict__ out,
                                                     int variants, size_t var_stride_u4)
{
    __shared__ uint4 tile[32][33];
    const int r0 = blockIdx.x * 32;
    for (int j = threadIdx.x; j < 1024; j += 256) {
        const int rl = j >> 5, cc = j & 31, r = r0 + rl;
        if (r < rows) {
            const float* sp = src + (size_t)r * 256 + cc * 8;
            const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
            uint4 o;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
            o2[0] = __floats2bfloat162_rn(a.x, a.y); o2[1] = __floats2bfloat162_rn(a.z, a.w);
            o2[2] = __floats2bfloat162_rn(b.x, b.y); o2[3] = __floats2bfloat162_rn(b.z, b.w);
            tile[rl][cc] = o;
        }
    }
    __syncthreads();
    uint4* o = reinterpret_cast<uint4*>(out);
    for (int j = threadIdx.x; j < 1024; j += 256) {
        const int cl = j >> 5, rl = j & 31, r = r0 + rl;
        if (r < rows) {
            const int bb = r / 42, p = r - bb * 42, x = p % 6;
            const size_t at = (size_t)cl * r_alloc + TC_HALO + (size_t)bb * TC_RPB + p;
            const uint4 v = tile[rl][cl], zero = make_uint4(0u, 0u, 0u, 0u);
            o[at] = v;
            if (variants == 3) { o[var_stride_u4 + at] = x == 0 ? zero : v; o[2 * var_stride_u4 + at] = x == 5 ? zero : v; }
        }
    }
}

// ---------------------------------------------------------------- weight gradient of a 256 -> 256 convolution, no unrolled operand
//   dW[t][ci][co] = sum over board cells r of a[nb(r, t)][ci] * dz[r][co]
// as nine GEMMs D_t[ci][co] = A_t^T . dZ with K = rows of the 48-row layout.  Both operands are read straight from the chunked bf16
// buffers [chunk][row][8]: seen with the row index as K they are MN-MAJOR UMMA operands (8 K-rows x 16 bytes of MN per core matrix,
// K groups 128 B apart = LBO, chunks = SBO apart), so the transposition costs nothing, and tap t is the A operand read `shift(t)` rows
// further on — a 16-byte-aligned address offset.  Top / bottom borders come from the zero rows between boards, left / right borders
// from the x-masked copies of dz (k_tc_chunk_rpb variants).
// One CTA = one tap x one K split: M = 2 x 128 input channels (two accumulators, all 512 TMEM columns), N = 256, 64 K-rows per stage.
// The kernel is bound by L2 -> SM delivery (116 FLOP per operand byte; measured 255 MB into the SMs in 64 us = 4 TB/s, tensor pipe
// active 21 %), so the three taps of one kernel COLUMN (same dx: same dz copy, activation rows 6 apart) run as a CLUSTER of three
// CTAs that share every operand slab by multicast: each CTA requests a third of the slabs for all three (a: 64 + 12 rows, every CTA
// reads its own 6-row offset; dz: 64 rows), a stage is refilled once the MMAs of all three have retired (tcgen05.commit multicast to
// the three "empty" barriers).  Measured gain of the multicast: 4 us per launch (the SMs still ingest the same bytes); one contiguous
// copy per operand per stage instead of 64 slabs would give another 7 us (timing experiment) — the lever that is left is more
// FLOPs per delivered byte (two taps per CTA pair sharing dz and overlapping activation rows).
#define WG_STAGES 3
#define WG_KROWS 64
#define WG_A_ROWS (WG_KROWS + 12)
#define WG_A_STAGE (32 * WG_A_ROWS * 16)                      // 38912
#define WG_B_STAGE (32 * WG_KROWS * 16)                       // 32768
#define WG_SMEM (WG_STAGES * (WG_A_STAGE + WG_B_STAGE) + 16 * 8 + 16)
#define WG_IDESC (TC_IDESC | (1u << 15) | (1u << 16))         // A and B MN-major

__global__ void __launch_bounds__(192, 1)
k_tc_wgrad(const __nv_bfloat16* __restrict__ a_rpb, const __nv_bfloat16* __restrict__ dz_rpb, size_t var_stride_bytes, int r_alloc, int kb_total,
           int kb_per_split, float* __restrict__ part)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + WG_STAGES * WG_A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + WG_STAGES * WG_B_STAGE);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + WG_STAGES;
    uint64_t* bar_acc = bars + 2 * WG_STAGES;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int kx = (int)(blockIdx.x / 3), ky = (int)crank, tap = ky * 3 + kx;          // cluster = the three ky of one kx
    const int split = blockIdx.y;
    const int kb0 = split * kb_per_split;
    const int nkb = min(kb_per_split, kb_total - kb0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 3); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ---- loaders: lane l owns chunk l of both operands; the CTA with rank l % 3 requests it for the whole cluster
        const int var = kx == 0 ? 1 : (kx == 2 ? 2 : 0);
        const bool mine = (uint32_t)(lane % 3) == crank;
        const uint8_t* a8 = reinterpret_cast<const uint8_t*>(a_rpb) + ((size_t)lane * r_alloc + TC_HALO - 6 + (kx - 1)) * 16;
        const uint8_t* b8 = reinterpret_cast<const uint8_t*>(dz_rpb) + (size_t)var * var_stride_bytes + ((size_t)lane * r_alloc + TC_HALO) * 16;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % WG_STAGES, k = i / WG_STAGES;
            if (k > 0) mbar_wait_cluster(bar_empty + s, (uint32_t)((k - 1) & 1));
            if (lane == 0) mbar_expect_tx(bar_full + s, WG_A_STAGE + WG_B_STAGE);
            __syncwarp();
            if (mine) {
                const size_t r0 = (size_t)(kb0 + i) * WG_KROWS * 16;
                bulk_g2s_multicast(sA + (size_t)s * WG_A_STAGE + (size_t)lane * (WG_A_ROWS * 16), a8 + r0, WG_A_ROWS * 16, bar_full + s, (uint16_t)7);
                bulk_g2s_multicast(sB + (size_t)s * WG_B_STAGE + (size_t)lane * (WG_KROWS * 16), b8 + r0, WG_KROWS * 16, bar_full + s, (uint16_t)7);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t a_base = smem_u32(sA) + (uint32_t)(ky * 6 * 16), b_base = smem_u32(sB);      // this CTA's row offset inside the shared slab
            for (int i = 0; i < nkb; ++i) {
                const int s = i % WG_STAGES, k = i / WG_STAGES;
                mbar_wait(bar_full + s, (uint32_t)(k & 1));
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < WG_KROWS / 16; ++kk) {
                    const uint64_t bdesc = umma_desc(b_base + (uint32_t)(s * WG_B_STAGE + kk * 256), 128, WG_KROWS * 16);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint64_t adesc = umma_desc(a_base + (uint32_t)(s * WG_A_STAGE + h * 16 * WG_A_ROWS * 16 + kk * 256), 128, WG_A_ROWS * 16);
                        tc_mma_bf16(tmem_base + (uint32_t)(h * 256), adesc, bdesc, WG_IDESC, (i > 0 || kk > 0) ? 1u : 0u);
                    }
                }
                tc_commit_multicast(bar_empty + s, (uint16_t)7);
            }
            tc_commit(bar_acc);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(bar_acc, 0u);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            float* crow = part + ((size_t)split * (9 * 256) + (size_t)(tap * 256 + h * 128 + q * 32 + lane)) * 256;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 256 + c * 32), v);
                tc_ld_wait();
                if (nkb > 0) {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        *reinterpret_cast<float4*>(crow + c * 32 + e * 4) = make_float4(__uint_as_float(v[e * 4]), __uint_as_float(v[e * 4 + 1]),
                                                                                       __uint_as_float(v[e * 4 + 2]), __uint_as_float(v[e * 4 + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                       // nobody leaves while a peer may still write into its shared memory or barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// fp32 HWIO [9][256][256] (device) -> the stage order of pack_conv_pair(kb_outer): [K group][tap][half][chunk 4][128 out][8].
// flip: the data gradient's kernel, tap 8 - t with input and output channels exchanged.  Thread = one 16-byte group (8 k values).
__global__ void __launch_bounds__(256) k_tc_pack_pair(const float* __restrict__ w, int flip, __nv_bfloat16* __restrict__ dst)
{
    const int i = blockIdx.x * 256 + threadIdx.x;                      // ((((kb * 9 + tap) * 2 + h) * 4 + ch) * 128 + n)
    if (i >= 8 * 9 * 2 * 4 * 128) return;
    const int n = i & 127, ch = (i >> 7) & 3, h = (i >> 9) & 1, st = i >> 10, tap = st % 9, kb = st / 9;
    const int co = h * 128 + n, ci0 = (kb * 4 + ch) * 8;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e)
        f[e] = flip ? w[((size_t)(8 - tap) * 256 + co) * 256 + ci0 + e] : w[((size_t)tap * 256 + ci0 + e) * 256 + co];
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    reinterpret_cast<uint4*>(dst)[i] = o;
}

int az_tc_conv_raw_reserve(AzTcConvScratch* sc, int n)
{
    if (sc->max_pairs == 0) {
        AZ_CUDA(cudaFuncSetAttribute(k_nn_conv_tc3<T3_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES));
        AZ_CUDA(cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        int dev = 0, sms = 148;
        AZ_CUDA(cudaGetDevice(&dev));
        AZ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(sms / 2 * 2)); cfg.blockDim = dim3(T3_THREADS); cfg.dynamicSmemBytes = T3_SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        AZ_CUDA(cudaOccupancyMaxActiveClusters(&nc, k_nn_conv_tc3<T3_RAW>, &cfg));
        sc->max_pairs = nc < sms / 2 ? nc : sms / 2;
        sc->n_sm = sms;
        if (sc->max_pairs < 1) { az_set_error("the device cannot hold one CTA pair of the tower kernel"); return AZ_ERR_CUDA; }
        AZ_CUDA(cudaMalloc(&sc->d_w, TC_LAYER_BYTES));
    }
    if (n <= sc->cap_boards) return AZ_OK;
    cudaFree(sc->d_in); cudaFree(sc->d_in3); sc->d_in = nullptr; sc->d_in3 = nullptr; sc->cap_boards = 0;
    for (__nv_bfloat16* p : sc->a_rpb) cudaFree(p);
    sc->a_rpb.clear();
    int tiles = (n * TC_RPB + TC_TILE_ROWS - 1) / TC_TILE_ROWS;
    tiles += tiles & 1;
    const int r_alloc = tiles * TC_TILE_ROWS + 2 * TC_HALO;
    const size_t bytes = (size_t)TC_CHUNKS * r_alloc * 16;
    AZ_CUDA(cudaMalloc(&sc->d_in, bytes));
    AZ_CUDA(cudaMemset(sc->d_in, 0, bytes));                  // padding rows and halos must read as zero
    AZ_CUDA(cudaMalloc(&sc->d_in3, 3 * bytes));
    AZ_CUDA(cudaMemset(sc->d_in3, 0, 3 * bytes));
    sc->cap_boards = n; sc->r_alloc = r_alloc; sc->dz_boards = 0;
    return AZ_OK;
}

static int conv_raw_launch(AzTcConvScratch* sc, const __nv_bfloat16* in_rpb, int n, const float* d_w, int flip, float* d_out, cudaStream_t s)
{
    k_tc_pack_pair<<<(8 * 9 * 2 * 4 * 128) / 256, 256, 0, s>>>(d_w, flip, reinterpret_cast<__nv_bfloat16*>(sc->d_w));
    AZ_CUDA(cudaGetLastError());
    const int tiles = (n * TC_RPB + TC_TILE_ROWS - 1) / TC_TILE_ROWS, pitems = (tiles + 1) / 2;
    const int pgrid = 2 * (pitems < sc->max_pairs ? pitems : sc->max_pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pgrid); cfg.blockDim = dim3(T3_THREADS); cfg.dynamicSmemBytes = T3_SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AZ_CUDA(cudaLaunchKernelEx(&cfg, k_nn_conv_tc3<T3_RAW>, in_rpb, (const uint8_t*)sc->d_w, (const float*)nullptr, (const float*)nullptr,
                               (const __nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr, n, sc->r_alloc, tiles, d_out, (const float4*)nullptr));
    return AZ_OK;
}

int az_tc_conv_raw(AzTcConvScratch* sc, const float* d_src, int n, const float* d_w, int flip, float* d_out, cudaStream_t s)
{
    int rc = az_tc_conv_raw_reserve(sc, n); if (rc) return rc;
    const int rows = n * 42;
    // (boards a larger earlier batch left behind are harmless here: every board's own six zero rows plus the x-masked operand
    // copies separate it from its neighbours, and rows of boards >= n are computed but never stored)
    k_tc_chunk_rpb<<<(unsigned)((rows + 31) / 32), 256, 0, s>>>(d_src, rows, sc->r_alloc, sc->d_in, 1, 0);
    return conv_raw_launch(sc, sc->d_in, n, d_w, flip, d_out, s);
}

// the gradient dz of a layer, converted once (plain + the two x-masked copies) for both of its consumers
int az_tc_dz_prepare(AzTcConvScratch* sc, const float* d_dz, int n, cudaStream_t s)
{
    int rc = az_tc_conv_raw_reserve(sc, n); if (rc) return rc;
    const size_t bytes = (size_t)TC_CHUNKS * sc->r_alloc * 16;
    // the weight gradient SUMS over rows: what a larger earlier batch left behind its last board must read as zero again
    if (n < sc->dz_boards) AZ_CUDA(cudaMemsetAsync(sc->d_in3, 0, 3 * bytes, s));
    sc->dz_boards = n;
    const int rows = n * 42;
    k_tc_chunk_rpb<<<(unsigned)((rows + 31) / 32), 256, 0, s>>>(d_dz, rows, sc->r_alloc, sc->d_in3, 3, bytes / 16);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

// data gradient from the prepared dz: out[r][ci] = sum_{t,co} dz[nb(r,t)][co] * w[8-t][ci][co]
int az_tc_dgrad_prepared(AzTcConvScratch* sc, int n, const float* d_w, float* d_out, cudaStream_t s)
{
    return conv_raw_launch(sc, sc->d_in3, n, d_w, 1, d_out, s);
}

int az_tc_layers_reserve(AzTcConvScratch* sc, int n, int layers)
{
    int rc = az_tc_conv_raw_reserve(sc, n); if (rc) return rc;
    const size_t bytes = (size_t)TC_CHUNKS * sc->r_alloc * 16;
    while ((int)sc->a_rpb.size() < layers) {
        __nv_bfloat16* p = nullptr;
        AZ_CUDA(cudaMalloc(&p, bytes));
        AZ_CUDA(cudaMemset(p, 0, bytes));
        sc->a_rpb.push_back(p);
    }
    return AZ_OK;
}

int az_tc_dz_target(AzTcConvScratch* sc, int n, cudaStream_t s, __nv_bfloat16** d_dz3, size_t* var_stride_u4)
{
    int rc = az_tc_conv_raw_reserve(sc, n); if (rc) return rc;
    const size_t bytes = (size_t)TC_CHUNKS * sc->r_alloc * 16;
    if (n < sc->dz_boards) AZ_CUDA(cudaMemsetAsync(sc->d_in3, 0, 3 * bytes, s));      // see az_tc_dz_prepare
    sc->dz_boards = n;
    *d_dz3 = sc->d_in3; *var_stride_u4 = bytes / 16;
    return AZ_OK;
}

int az_tc_conv_raw_rpb(AzTcConvScratch* sc, const __nv_bfloat16* in_rpb, int n, const float* d_w, int flip, float* d_out, cudaStream_t s)
{
    return conv_raw_launch(sc, in_rpb, n, d_w, flip, d_out, s);
}

// weight gradient partials [splits][9 * 256][256] from the prepared dz and the layer's fp32 input activation; returns the split count
int az_tc_wgrad_prepared(AzTcConvScratch* sc, const float* d_a, int n, float* d_part, int max_splits, int* splits_out, cudaStream_t s)
{
    const int rows = n * 42;
    k_tc_chunk_rpb<<<(unsigned)((rows + 31) / 32), 256, 0, s>>>(d_a, rows, sc->r_alloc, sc->d_in, 1, 0);
    return az_tc_wgrad_rpb(sc, sc->d_in, n, d_part, max_splits, splits_out, s);
}

int az_tc_wgrad_rpb(AzTcConvScratch* sc, const __nv_bfloat16* a_rpb, int n, float* d_part, int max_splits, int* splits_out, cudaStream_t s)
{
    const int kb_total = (n * TC_RPB + WG_KROWS - 1) / WG_KROWS;                   // 64-row K blocks that hold board rows
    int want = sc->n_sm / 9; if (want > max_splits) want = max_splits; if (want > kb_total) want = kb_total; if (want < 1) want = 1;
    const int per = (kb_total + want - 1) / want, splits = (kb_total + per - 1) / per;      // no empty split
    const size_t bytes = (size_t)TC_CHUNKS * sc->r_alloc * 16;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(9, (unsigned)splits); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = WG_SMEM; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 3; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        AZ_CUDA(cudaLaunchKernelEx(&cfg, k_tc_wgrad, a_rpb, (const __nv_bfloat16*)sc->d_in3, bytes, sc->r_alloc, kb_total, per, d_part));
    }
    *splits_out = splits;
    return AZ_OK;
}

void az_tc_conv_raw_release(AzTcConvScratch* sc)
{
    cudaFree(sc->d_in); cudaFree(sc->d_in3); cudaFree(sc->d_w);
    for (__nv_bfloat16* p : sc->a_rpb) cudaFree(p);
    *sc = AzTcConvScratch();
}

void az_nn_tc_release(az_nn* nn)
{
    if (!nn->tc) return;
    AzTcState* tc = nn->tc;
    for (AzTcScratch& sc : tc->scratch) {
        for (int i = 0; i < 3; ++i) cudaFree(sc.d_act[i]);
        cudaFree(sc.d_in); cudaFree(sc.d_head_z);
    }
    cudaFree(tc->d_wpacked); cudaFree(tc->d_wpacked3); cudaFree(tc->d_scale); cudaFree(tc->d_shift);
    cudaFree(tc->d_head_w);
    delete tc;
    nn->tc = nullptr;
}

static int tc_reserve(AzTcScratch* tc, int n)
{
    if (n <= tc->cap_boards) return AZ_OK;
    for (int i = 0; i < 3; ++i) { cudaFree(tc->d_act[i]); tc->d_act[i] = nullptr; }
    cudaFree(tc->d_in); tc->d_in = nullptr;
    cudaFree(tc->d_head_z); tc->d_head_z = nullptr;
    AZ_CUDA(cudaMalloc(&tc->d_head_z, sizeof(float4) * (size_t)n * 42));
    int tiles = (n * TC_RPB + TC_TILE_ROWS - 1) / TC_TILE_ROWS;
    tiles += tiles & 1;                                       // the pair kernel reads whole tile pairs
    int r_alloc = tiles * TC_TILE_ROWS + 2 * TC_HALO;
    size_t bytes = (size_t)TC_CHUNKS * r_alloc * 16;
    const int nv = 3;                                         // the stem's three operand copies
    for (int i = 0; i < 3; ++i) {
        AZ_CUDA(cudaMalloc(&tc->d_act[i], bytes));
        AZ_CUDA(cudaMemset(tc->d_act[i], 0, bytes));          // padding rows and halos must read as zero
    }
    AZ_CUDA(cudaMalloc(&tc->d_in, (size_t)nv * 2 * r_alloc * 16));
    AZ_CUDA(cudaMemset(tc->d_in, 0, (size_t)nv * 2 * r_alloc * 16));
    tc->in_var_stride = (size_t)2 * r_alloc * 8;
    tc->cap_boards = n; tc->n_tiles = tiles; tc->r_alloc = r_alloc;
    return AZ_OK;
}

int az_nn_tc_forward(az_nn* nn, const float* d_x, const uint32_t* d_env_state, int n, float* d_policy, float* d_value, cudaStream_t s,
                     int state_stride, int scratch_set, bool shared_device)
{
    if (state_stride < n) state_stride = n;
    AzTcState* st = nn->tc;
    if (!st || !st->d_wpacked) { az_set_error("network not finalized"); return AZ_ERR_NOT_READY; }
    if (scratch_set < 0 || scratch_set >= AZ_TC_SCRATCH_SETS) { az_set_error("scratch set out of range"); return AZ_ERR_INVALID_ARG; }
    AzTcScratch* tc = &st->scratch[scratch_set];
    int rc = tc_reserve(tc, n); if (rc) return rc;
    // from game states (d_x == NULL): one kernel packs the stem's bf16 input straight from the states; from an fp32 tensor:
    // k_nn_pack_input_tc.  Same expressions and roundings on both routes (tests/test_mcts_gpu.py compares a search fed by the
    // first with an oracle search fed by the second)
    const bool from_state = d_x == nullptr;
    // buffers are sized for cap_boards; only the tiles that hold boards of this call are computed
    const int tiles = (n * TC_RPB + TC_TILE_ROWS - 1) / TC_TILE_ROWS;
    const int grid = tiles < st->n_sm ? tiles : st->n_sm;
    const int layers = 2 * nn->blocks;
    int cur = 0, tmp = 1, nxt = 2;
    if (from_state) k_nn_pack_state_tc<<<(n + 3) / 4, 128, 0, s>>>(d_env_state, n, state_stride, tc->d_in, tc->r_alloc, tc->in_var_stride);
    else k_nn_pack_input_tc<<<(n * 42 + 255) / 256, 256, 0, s>>>(d_x, n, tc->d_in, tc->r_alloc, tc->in_var_stride);
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA((launch_conv<2, true, 3>(grid, s, tc->d_in, st->d_wpacked, st->d_scale + layers * 256, st->d_shift + layers * 256, nullptr, tc->d_act[cur], n,
                                     tc->r_alloc, tiles, tc->in_var_stride * 2)));
    const int pitems = (tiles + 1) / 2;
    const int pgrid = 2 * (pitems < st->max_pairs ? pitems : st->max_pairs);
    for (int i = 0; i < nn->blocks; ++i) {
        for (int h = 0; h < 2; ++h) {                        // branch2a: cur -> tmp; branch2b: tmp (+ cur as the residual) -> nxt
            const int L = 2 * i + h;
            const __nv_bfloat16* src = h ? tc->d_act[tmp] : tc->d_act[cur];
            const __nv_bfloat16* res = h ? tc->d_act[cur] : nullptr;
            __nv_bfloat16* dst = h ? tc->d_act[nxt] : tc->d_act[tmp];
            const bool last = L == layers - 1;              // the last layer feeds the heads directly: no activation store
            AZ_CUDA(launch_conv_pair3(pgrid, s, src, st->d_wpacked3 + (size_t)L * TC_LAYER_BYTES, st->d_scale + L * 256, st->d_shift + L * 256, res, dst, n,
                                      tc->r_alloc, tiles, last ? st->d_head_w : nullptr, last ? tc->d_head_z : nullptr, !shared_device));
        }
        int o = cur; cur = nxt; nxt = o;
    }
    k_nn_heads_tail<<<(n + HB - 1) / HB, 256, 0, s>>>(tc->d_head_z, n, az_nn_head_params(nn), d_policy, d_value);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}
