// az_mcts6.cu — lockstep MCTS for the SIX-PLAYER extension (SIXPLAYER.md; BASELINE.json configs[3]) and the az_mcts6_* C ABI.
//
// No reference parity exists for this game (PLAYER_COUNT = 2 upstream, /root/reference/src/risk_game/state/state.h:13); the
// checker is oracle/risk6_oracle.c (r6_mcts_search), compared bit for bit by tests/test_mcts6_gpu.py.  The search is the
// reference's AlphaZeroMCTS (player/alpha_zero/alphazero_mcts.cpp: PUCT :67-119, search :322-377, addValue :8-21, policy from the
// visit counts :121-148, move choice :379-412) with the two changes SIXPLAYER.md states: the table is cleared before every search
// (no trimNodes carry-over), and a leaf's scalar value belongs to the seat to move there — on the way up a node whose mover is
// that seat receives +v, every other node -v / 5.  One descent per tree per leaf batch (THREADS_PER_MCTS = 1 semantics).
//
// Mapping: one warp per game.  The game logic of a descent (status, legal mask, transition: az_game6.cuh) runs on lane 0 over the
// warp's shared-memory state row — the six-seat rules are thread code, shared with the environment kernels —, the table work
// (hash-window probing, node records, PUCT argmax) uses all 32 lanes.  The leaf batch goes through the same policy / value tower
// as the two-player game (k_env6_encode6 -> az_nn_forward_dev).  Deliberately simpler than az_mcts.cu (one pool per game, no
// virtual loss, no arena, no sample records): a throughput extension, the tower is what it spends its time in.
#include <cstring>
#include <new>

#include "az_common.cuh"
#include "az_game6.cuh"
#include "az_nn.cuh"
#include "az_pseudo_net.h"

#define M6_WARPS 4
#define M6_FULL 0xffffffffu
#define M6_KEY 27                       // state words 0..26 identify a position (word 27 = the slot's move counter)
#define M6_NODE_WORDS 192
#define M6_VALID 28
#define M6_SUMN 30
#define M6_VALUE 31
#define M6_P 32
#define M6_Q 76
#define M6_N 120
enum { M6_EVAL_NN = 0, M6_EVAL_PSEUDO = 1, M6_EVAL_UNIFORM = 2 };
enum { M6C_SIMS = 0, M6C_EVALS = 1, M6C_OVERFLOW = 2, M6C_STEPS = 3, M6C_GAMES = 4, M6C_DRAWS = 5, M6C_WIN0 = 6, M6C_N = 12 };

struct Mcts6Dev {
    int n, cap, H, dmax;
    uint32_t* nodes;        // [n][cap][M6_NODE_WORDS]
    uint32_t* index;        // [n][H]   0 = empty, else tag16 << 16 | (node + 1)
    uint32_t* count;        // [n]
    uint32_t* path;         // [n][dmax]  node | move << 16 | mover << 22
    uint32_t* path_len;     // [n]
    uint32_t* leaf_state;   // [28][n]
    uint64_t* leaf_valid;   // [n]
    int32_t* pending;       // [n]
    float* term_value;      // [n]   value of a descent that ended in a finished game
    int32_t* term_seat;     // [n]   and the seat it belongs to
    float* nn_policy;       // [n][43]
    float* nn_value;        // [n]
    uint32_t* root_state;   // the env6 state [28][n]
    uint32_t* out_visits; float* out_pi; uint8_t* out_move; int8_t* out_status; float* out_q; float* out_p; uint32_t* out_sumn; int32_t* out_table;
    unsigned long long* counters;
    float c1, c2, cpuct;
    uint64_t seed; uint32_t first_game;
    AzRulesDev rules;
    int eval_mode, temp_threshold;
};

struct M6Smem {
    uint64_t tab[AZ_TABLE_U64];
    uint32_t row[M6_WARPS][E6_WORDS];
    uint32_t scratch[M6_WARPS][12];
};

// lane-0 game context over the warp's row: the column accessors of az_game.cuh with a 4-byte stride are plain byte rows
__device__ __forceinline__ void m6_bind(E6Ctx& c, uint32_t* row, uint32_t* scratch)
{
    c.army.base = (uint8_t*)row; c.army.stride_bytes = 4;
    c.owner.base = (uint8_t*)(row + 11); c.owner.stride_bytes = 4;
    c.scratch.base = (uint8_t*)scratch; c.scratch.stride_bytes = 4;
}
__device__ __forceinline__ void m6_unpack_row(E6Ctx& c, const uint32_t* row)
{
    e6_masks(c);
    e6_unpack(c.g, row[22], row[23], row[24], row[25], row[26]);
    c.ply = row[27];
}
__device__ __forceinline__ void m6_pack_row(const E6Ctx& c, uint32_t* row)
{
    const AzGame6& g = c.g;
    row[22] = g.cards_lo;
    row[23] = (g.cards_hi & 0xffffu) | ((g.allow_draw & 0xffu) << 16) | ((g.attacks & 0xffu) << 24);
    row[24] = (g.round & 0xffffu) | (g.cur << 16) | (g.card_sets << 24);
    row[25] = (g.reinf & 0xffu) | (g.phase << 8) | (g.mob_from << 16) | (g.mob_to << 24);
    row[26] = g.pools & 0xffffffu;
}

__device__ __forceinline__ uint64_t m6_hash(uint32_t kw, int lane)
{
    uint64_t h = ((uint64_t)kw + 0x9E3779B97F4A7C15ull * (uint64_t)(lane + 1)) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    if (lane >= M6_KEY) h = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t lo = __shfl_xor_sync(M6_FULL, (uint32_t)h, o), hi = __shfl_xor_sync(M6_FULL, (uint32_t)(h >> 32), o);
        h += ((uint64_t)hi << 32) | lo;
    }
    return az_pn_mix(h);
}
__device__ __forceinline__ uint32_t* m6_node(const Mcts6Dev& m, int gi, uint32_t idx) { return m.nodes + ((size_t)gi * m.cap + idx) * M6_NODE_WORDS; }

// the node of the position whose key words sit in lanes 0..26 (kw), or -1; empty_slot = where an insert would go
__device__ __forceinline__ int m6_lookup(const Mcts6Dev& m, int gi, uint64_t h, uint32_t kw, int lane, int& empty_slot)
{
    const uint32_t* base = m.index + (size_t)gi * m.H;
    const uint32_t tag = (uint32_t)(h >> 48) & 0xffffu, start = (uint32_t)h & (uint32_t)(m.H - 1);
    empty_slot = -1;
    for (int win = 0; win < m.H; win += 32) {
        const uint32_t slot = (start + (uint32_t)win + (uint32_t)lane) & (uint32_t)(m.H - 1);
        const uint32_t e = base[slot];
        const uint32_t empties = __ballot_sync(M6_FULL, e == 0);
        uint32_t matches = __ballot_sync(M6_FULL, e != 0 && (e >> 16) == tag);
        const int first_empty = empties ? (__ffs((int)empties) - 1) : 32;
        if (first_empty < 32) matches &= (1u << first_empty) - 1u;
        while (matches) {
            const int l = __ffs((int)matches) - 1; matches &= matches - 1;
            const uint32_t idx = (__shfl_sync(M6_FULL, e, l) & 0xffffu) - 1u;
            const uint32_t* nd = m6_node(m, gi, idx);
            const uint32_t w = lane < M6_KEY ? nd[lane] : 0u;
            if (__ballot_sync(M6_FULL, w == kw) == M6_FULL) return (int)idx;
        }
        if (first_empty < 32) { empty_slot = (int)((start + (uint32_t)win + (uint32_t)first_empty) & (uint32_t)(m.H - 1)); return -1; }
    }
    return -1;
}

// getNextBestMoveAndSetVisited (alphazero_mcts.cpp:67-119), ascending move index on ties; one descent at a time: no active_N
__device__ __forceinline__ int m6_select(const Mcts6Dev& m, const uint32_t* nd, int lane)
{
    const uint64_t valid = (uint64_t)nd[M6_VALID] | ((uint64_t)nd[M6_VALID + 1] << 32);
    const float sq = __fsqrt_rn(__fadd_rn(1.0f, (float)nd[M6_SUMN]));
    float bu = -INFINITY; int bi = 64;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = lane + 32 * k;
        if (i < AZ_MOVES && ((valid >> i) & 1ull)) {
            const float P = __uint_as_float(nd[M6_P + i]), Q = __uint_as_float(nd[M6_Q + i]);
            const float noiseP = __fadd_rn(__fmul_rn(m.c1, P), m.c2);
            const float v = __fmul_rn(__fmul_rn(noiseP, m.cpuct), sq);
            const float u = __fadd_rn(Q, __fdiv_rn(v, __fadd_rn(1.0f, (float)nd[M6_N + i])));
            if (u > bu) { bu = u; bi = i; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ou = __shfl_xor_sync(M6_FULL, bu, o); const int oi = __shfl_xor_sync(M6_FULL, bi, o);
        if (ou > bu || (ou == bu && oi < bi)) { bu = ou; bi = oi; }
    }
    return bi;
}

// addValue up the recorded path: +v for the nodes of the seat the leaf value belongs to, -v / 5 for every other node
__device__ __forceinline__ void m6_backup(const Mcts6Dev& m, int gi, float v, uint32_t seat, uint32_t len, int lane)
{
    __syncwarp();
    if (lane == 0) {
        const float other = __fdiv_rn(-v, 5.0f);
        for (int d = (int)len - 1; d >= 0; --d) {
            const uint32_t e = m.path[(size_t)gi * m.dmax + d];
            const uint32_t idx = e & 0xffffu, mv = (e >> 16) & 63u, mover = (e >> 22) & 7u;
            const float val = mover == seat ? v : other;
            uint32_t* nd = m6_node(m, gi, idx);
            const uint32_t N = nd[M6_N + mv];
            const float Q = __uint_as_float(nd[M6_Q + mv]);
            const float q = N == 0 ? val : __fdiv_rn(__fadd_rn(__fmul_rn((float)N, Q), val), (float)(N + 1u));
            nd[M6_Q + mv] = __float_as_uint(q);
            nd[M6_N + mv] = N + 1u;
            nd[M6_SUMN] += 1u;
        }
        if (len > 0) atomicAdd(&m.counters[M6C_SIMS], 1ull);
        m.path_len[gi] = 0;
    }
    __syncwarp();
}

__device__ __forceinline__ uint64_t m6_pn_key(const uint32_t* row)
{
    const uint8_t* army = (const uint8_t*)row; const uint8_t* owner = (const uint8_t*)(row + 11);
    uint8_t land[AZ_LANDS];
    for (int i = 0; i < AZ_LANDS; ++i) land[i] = (uint8_t)(army[i] + 40 * owner[i]);
    return az_pn_key(land, (int)((row[24] >> 16) & 0xffu), (int)(row[24] & 0xffffu), (int)((row[25] >> 8) & 0xffu));
}

// complete the descent of the previous round: expand the queued leaf with the evaluator's output (StateSimulations ctor :26-42 after
// NNOutputData::normalize, alphazero_nn_data.cpp:3-27), or take the finished game's value, and back it up
__device__ __forceinline__ void m6_complete(const Mcts6Dev& m, int gi, uint32_t* row, int lane)
{
    const uint32_t len = m.path_len[gi];
    float v = m.term_value[gi];
    uint32_t seat = (uint32_t)m.term_seat[gi];
    if (m.pending[gi]) {
        __syncwarp();
        if (lane < E6_WORDS) row[lane] = m.leaf_state[(size_t)lane * m.n + gi];
        __syncwarp();
        const uint64_t valid = m.leaf_valid[gi];
        float p0 = 0.0f, p1 = 0.0f, value;
        if (m.eval_mode == M6_EVAL_NN) {
            p0 = m.nn_policy[(size_t)gi * AZ_MOVES + lane];
            if (lane + 32 < AZ_MOVES) p1 = m.nn_policy[(size_t)gi * AZ_MOVES + 32 + lane];
            value = m.nn_value[gi];
        } else if (m.eval_mode == M6_EVAL_PSEUDO) {
            const uint64_t key = m6_pn_key(row);
            p0 = az_pn_policy(key, lane);
            if (lane + 32 < AZ_MOVES) p1 = az_pn_policy(key, lane + 32);
            value = az_pn_value(key);
        } else { p0 = 1.0f / 43.0f; p1 = lane + 32 < AZ_MOVES ? 1.0f / 43.0f : 0.0f; value = 0.0f; }
        if (!((valid >> lane) & 1ull)) p0 = 0.0f;
        if (lane + 32 >= AZ_MOVES || !((valid >> (lane + 32)) & 1ull)) p1 = 0.0f;
        float sum = 0.0f;                                    // ascending-order sum; the masked entries are +0 and x + 0 == x exactly
#pragma unroll
        for (int i = 0; i < 32; ++i) sum = __fadd_rn(sum, __shfl_sync(M6_FULL, p0, i));
#pragma unroll
        for (int i = 0; i < AZ_MOVES - 32; ++i) sum = __fadd_rn(sum, __shfl_sync(M6_FULL, p1, i));
        if (p0 > 0.0f) p0 = __fdiv_rn(p0, sum);
        if (p1 > 0.0f) p1 = __fdiv_rn(p1, sum);
        const uint32_t kw = lane < M6_KEY ? row[lane] : 0u;
        const uint64_t h = m6_hash(kw, lane);
        int ins;
        const int found = m6_lookup(m, gi, h, kw, lane, ins);
        const uint32_t cnt = m.count[gi];
        if (found < 0 && (int)cnt < m.cap && ins >= 0) {
            uint32_t* nd = m6_node(m, gi, cnt);
            if (lane < M6_KEY) nd[lane] = kw;
            if (lane == 27) nd[M6_VALID] = (uint32_t)valid;
            if (lane == 28) nd[M6_VALID + 1] = (uint32_t)(valid >> 32);
            if (lane == 29) nd[M6_SUMN] = 0u;
            if (lane == 30) nd[M6_VALUE] = __float_as_uint(value);
            nd[M6_P + lane] = __float_as_uint(p0); nd[M6_Q + lane] = 0u; nd[M6_N + lane] = 0u;
            if (lane < AZ_MOVES - 32) { nd[M6_P + 32 + lane] = __float_as_uint(p1); nd[M6_Q + 32 + lane] = 0u; nd[M6_N + 32 + lane] = 0u; }
            __syncwarp();
            if (lane == 0) { m.count[gi] = cnt + 1; m.index[(size_t)gi * m.H + ins] = (((uint32_t)(h >> 48) & 0xffffu) << 16) | (cnt + 1u); }
            __syncwarp();
        } else if (found < 0 && lane == 0) atomicAdd(&m.counters[M6C_OVERFLOW], 1ull);
        v = value;
        seat = (row[24] >> 16) & 0xffu;                      // the seat to move at the leaf
        if (lane == 0) { m.pending[gi] = 0; atomicAdd(&m.counters[M6C_EVALS], 1ull); }
    }
    m6_backup(m, gi, v, seat, len, lane);
}

// one descent from the root (AlphaZeroMCTS::search :322-377 as a loop); sim < 0: only make sure the root is in the table
__device__ __forceinline__ void m6_descend(const Mcts6Dev& m, int gi, uint32_t* row, uint32_t* scratch, const AzTables& T, int sim, int lane)
{
    __syncwarp();
    if (lane < E6_WORDS) row[lane] = m.root_state[(size_t)lane * m.n + gi];
    __syncwarp();
    E6Ctx c; m6_bind(c, row, scratch);
    AzDicePhilox dice;
    if (lane == 0) { m6_unpack_row(c, row); dice.init(m.seed, m.first_game + (uint32_t)gi, c.ply, (uint32_t)(sim < 0 ? 0 : sim)); }
    uint32_t depth = 0;
    for (;;) {
        int st = 0; uint64_t valid = 0; uint32_t mover = 0;
        if (lane == 0) { st = e6_status(c.g, m.rules); valid = st == AZ_STATUS_RUNNING ? e6_valid(c.g, T, m.rules) : 0ull; mover = c.g.cur; }
        st = __shfl_sync(M6_FULL, st, 0);
        mover = __shfl_sync(M6_FULL, mover, 0);
        valid = ((uint64_t)__shfl_sync(M6_FULL, (uint32_t)(valid >> 32), 0) << 32) | __shfl_sync(M6_FULL, (uint32_t)valid, 0);
        if (st != AZ_STATUS_RUNNING) {                       // a finished game: its value belongs to the winner (a draw is worth 0 to everyone)
            const float tv = st == AZ_STATUS_DRAW ? 0.0f : 1.0f;
            const uint32_t seat = st == AZ_STATUS_DRAW ? mover : (uint32_t)st;
            if (lane == 0) { m.term_value[gi] = tv; m.term_seat[gi] = (int32_t)seat; m.pending[gi] = 0; }
            m6_backup(m, gi, tv, seat, depth, lane);
            depth = 0;
            break;
        }
        const uint32_t kw = lane < M6_KEY ? row[lane] : 0u;
        const uint64_t h = m6_hash(kw, lane);
        int ins;
        const int idx = m6_lookup(m, gi, h, kw, lane, ins);
        if (idx < 0) {                                       // unseen position: queue it for the evaluator
            if (lane < E6_WORDS) m.leaf_state[(size_t)lane * m.n + gi] = row[lane];
            if (lane == 0) { m.leaf_valid[gi] = valid; m.pending[gi] = 1; m.term_value[gi] = 0.0f; m.term_seat[gi] = (int32_t)mover; }
            break;
        }
        if (sim < 0) { if (lane == 0) m.pending[gi] = 0; break; }
        if ((int)depth >= m.dmax) { if (lane == 0) { atomicAdd(&m.counters[M6C_OVERFLOW], 1ull); m.pending[gi] = 0; m.term_value[gi] = 0.0f; } depth = 0; break; }
        const int mv = m6_select(m, m6_node(m, gi, (uint32_t)idx), lane);
        if (lane == 0) {
            m.path[(size_t)gi * m.dmax + depth] = (uint32_t)idx | ((uint32_t)mv << 16) | (mover << 22);
            e6_move(c, T, m.rules, mv, dice);
            m6_pack_row(c, row);
        }
        __syncwarp();
        depth++;
    }
    if (lane == 0) m.path_len[gi] = depth;
    __syncwarp();
}

__global__ void __launch_bounds__(M6_WARPS * 32) k_mcts6_begin(Mcts6Dev m)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * M6_WARPS + warp;
    if (gi >= m.n) return;
    uint32_t* ix = m.index + (size_t)gi * m.H;
    for (int i = lane; i < m.H; i += 32) ix[i] = 0u;
    if (lane == 0) { m.count[gi] = 0u; m.pending[gi] = 0; m.path_len[gi] = 0u; m.term_value[gi] = 0.0f; m.term_seat[gi] = 0; }
}

__global__ void __launch_bounds__(M6_WARPS * 32) k_mcts6_sim(Mcts6Dev m, const uint64_t* __restrict__ g_tab, int round)
{
    __shared__ M6Smem sm;
    for (int i = threadIdx.x; i < AZ_TABLE_U64; i += blockDim.x) sm.tab[i] = g_tab[i];
    __syncthreads();
    const AzTables T = az_tables_from_smem(sm.tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * M6_WARPS + warp;
    if (gi >= m.n) return;
    m6_complete(m, gi, sm.row[warp], lane);
    m6_descend(m, gi, sm.row[warp], sm.scratch[warp], T, round, lane);
}

// after the last simulation: root statistics (calculateMoveProbability :121-148), move choice (pickHigestWeightedMove :397-412 /
// pickRandomWeightedMove :379-395 under the temperature rule), optionally the real move on the env6 state (+ re-deal of finished
// games when auto_reset).  The float arithmetic runs on lane 0 in the oracle's order.
__global__ void __launch_bounds__(M6_WARPS * 32) k_mcts6_finish(Mcts6Dev m, const uint64_t* __restrict__ g_tab, int pick_mode, int apply_move, int auto_reset)
{
    __shared__ M6Smem sm;
    for (int i = threadIdx.x; i < AZ_TABLE_U64; i += blockDim.x) sm.tab[i] = g_tab[i];
    __syncthreads();
    const AzTables T = az_tables_from_smem(sm.tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * M6_WARPS + warp;
    if (gi >= m.n) return;
    uint32_t* row = sm.row[warp];
    m6_complete(m, gi, row, lane);
    __syncwarp();
    if (lane < E6_WORDS) row[lane] = m.root_state[(size_t)lane * m.n + gi];
    __syncwarp();
    const uint32_t kw = lane < M6_KEY ? row[lane] : 0u;
    const uint64_t h = m6_hash(kw, lane);
    int ins;
    const int idx = m6_lookup(m, gi, h, kw, lane, ins);
    if (lane != 0) return;
    E6Ctx c; m6_bind(c, row, sm.scratch[warp]);
    m6_unpack_row(c, row);
    int status = e6_status(c.g, m.rules);
    int move = AZ_NONE;
    const uint32_t game = m.first_game + (uint32_t)gi;
    uint32_t* ov = m.out_visits + (size_t)gi * AZ_MOVES; float* op = m.out_pi + (size_t)gi * AZ_MOVES;
    float* oq = m.out_q + (size_t)gi * AZ_MOVES; float* opp = m.out_p + (size_t)gi * AZ_MOVES;
    for (int i = 0; i < AZ_MOVES; ++i) { ov[i] = 0u; op[i] = 0.0f; oq[i] = 0.0f; opp[i] = 0.0f; }
    m.out_sumn[gi] = 0u; m.out_table[gi] = (int32_t)m.count[gi];
    if (status == AZ_STATUS_RUNNING && idx >= 0) {
        const uint32_t* nd = m6_node(m, gi, (uint32_t)idx);
        const uint64_t valid = (uint64_t)nd[M6_VALID] | ((uint64_t)nd[M6_VALID + 1] << 32);
        float sum = 0.0f;
        for (int i = 0; i < AZ_MOVES; ++i) if ((valid >> i) & 1ull) {
            ov[i] = nd[M6_N + i]; oq[i] = __uint_as_float(nd[M6_Q + i]); opp[i] = __uint_as_float(nd[M6_P + i]);
            op[i] = (float)nd[M6_N + i]; sum = __fadd_rn(sum, op[i]);
        }
        for (int i = 0; i < AZ_MOVES; ++i) op[i] = __fdiv_rn(op[i], sum);
        m.out_sumn[gi] = nd[M6_SUMN];
        const bool sample = pick_mode == 1 && (int)c.g.round <= m.temp_threshold;
        const az_u32x4 blk = az_rng_block(m.seed, game, c.ply, AZ_STREAM_REAL, 0);
        if (!sample) {
            float best = 0.0f;
            for (int i = 0; i < AZ_MOVES; ++i) if (op[i] > best) { best = op[i]; move = i; }
        } else {
            float tot = 0.0f;
            for (int i = 0; i < AZ_MOVES; ++i) tot = __fadd_rn(tot, op[i]);
            const float a = __fmul_rn(tot, az_rng_unit_float(blk.z));
            float it = 0.0f;
            for (int i = 0; i < AZ_MOVES; ++i) { it = __fadd_rn(it, op[i]); if (it >= a) { move = i; break; } }
        }
        if (apply_move && move != AZ_NONE) {
            E6DiceWord dice; dice.w = blk.x;
            e6_move(c, T, m.rules, move, dice);
            c.ply++;
            atomicAdd(&m.counters[M6C_STEPS], 1ull);
            status = e6_status(c.g, m.rules);
            if (status != AZ_STATUS_RUNNING) {
                atomicAdd(&m.counters[M6C_GAMES], 1ull);
                if (status == AZ_STATUS_DRAW) atomicAdd(&m.counters[M6C_DRAWS], 1ull); else atomicAdd(&m.counters[M6C_WIN0 + status], 1ull);
                if (auto_reset) e6_new_game(c, m.seed, game, c.ply);
            }
            m6_pack_row(c, row);
            for (int w = 0; w < 27; ++w) m.root_state[(size_t)w * m.n + gi] = row[w];
            m.root_state[(size_t)27 * m.n + gi] = c.ply;
        }
    }
    m.out_move[gi] = (uint8_t)move;
    m.out_status[gi] = (int8_t)status;
}

// the six-seat network input (SIXPLAYER.md): the reference's [7][6][13] tensor with "enemy" = the seat that moves next and
// "neutral" = every other seat; one thread per position (r6_encode of the oracle, same expressions)
__global__ void __launch_bounds__(128) k_env6_encode(const uint32_t* __restrict__ st, int n, float* __restrict__ x)
{
    const int gi = blockIdx.x * 128 + threadIdx.x;
    if (gi >= n) return;
    uint32_t row[E6_WORDS];
    for (int w = 0; w < E6_WORDS; ++w) row[w] = st[(size_t)w * n + gi];
    const uint8_t* army = (const uint8_t*)row; const uint8_t* owner = (const uint8_t*)(row + 11);
    const uint32_t cur = (row[24] >> 16) & 0xffu, phase = (row[25] >> 8) & 0xffu, attacks = (row[23] >> 24) & 0xffu, allow = (row[23] >> 16) & 0xffu;
    uint64_t own[E6_PLAYERS] = { 0, 0, 0, 0, 0, 0 };
    int own_total = 0, other_total = 0;
    for (int i = 0; i < AZ_LANDS; ++i) {
        own[owner[i]] |= 1ull << i;
        if (owner[i] == cur) own_total += army[i]; else other_total += army[i];
    }
    uint32_t nx = cur;
    for (int k = 0; k < E6_PLAYERS; ++k) { nx = nx + 1 == E6_PLAYERS ? 0u : nx + 1; if (own[nx]) break; }
    const float ref = (float)az_reinforcement_value(own[cur]);
    float oref = 0.0f;
    for (uint32_t p = 0; p < E6_PLAYERS; ++p) if (p != cur && own[p]) oref = __fadd_rn(oref, (float)az_reinforcement_value(own[p]));
    const float army_share = __fdiv_rn((float)own_total, __fadd_rn((float)own_total, (float)other_total));
    const float reinf_share = __fdiv_rn(ref, __fadd_rn(ref, oref));
    float att = __fdiv_rn((float)attacks, 8.0f); att = att > 1.0f ? 1.0f : att;
    float* out = x + (size_t)gi * AZ_INPUT_FLOATS;
    for (int i = 0; i < AZ_LANDS; ++i) {
        float* f = out + i * 13;
        const float a = __fdiv_rn((float)army[i], 32.0f);
        f[0] = owner[i] == cur ? a : 0.0f;
        f[1] = (owner[i] == nx && nx != cur) ? a : 0.0f;
        f[2] = (owner[i] != cur && owner[i] != nx) ? a : 0.0f;
        f[3] = army_share; f[4] = reinf_share; f[5] = att; f[6] = allow ? 1.0f : 0.0f;
        for (uint32_t k = 0; k < 6; ++k) f[7 + k] = phase == k ? 1.0f : 0.0f;
    }
}

// ---------------------------------------------------------------- host side
struct az_env6;
uint32_t* az_env6_state_ptr(az_env6* e); int az_env6_n(const az_env6* e); int az_env6_device(const az_env6* e);
uint64_t az_env6_seed(const az_env6* e); uint32_t az_env6_first_game(const az_env6* e); const az_rules* az_env6_rules(const az_env6* e);

struct az_mcts6 {
    az_env6* env = nullptr; az_nn* nn = nullptr;
    int evaluator = M6_EVAL_NN, precision = AZ_NN_FP32, device = 0, sims = 0;
    Mcts6Dev d;
    float* d_x = nullptr;
    std::vector<void*> allocs;
};

template <class T> static int m6_alloc(az_mcts6* mc, T** p, size_t count, bool zero = true)
{
    AZ_CUDA(cudaMalloc(p, sizeof(T) * count));
    if (zero) AZ_CUDA(cudaMemset(*p, 0, sizeof(T) * count));
    mc->allocs.push_back(*p);
    return AZ_OK;
}

extern "C" int az_mcts6_create(az_env6* env, az_nn* nn, int evaluator, int precision, az_mcts6** out)
{
    AZ_REQUIRE(env && out, "NULL argument");
    AZ_REQUIRE(evaluator >= M6_EVAL_NN && evaluator <= M6_EVAL_UNIFORM, "unknown evaluator");
    AZ_REQUIRE(evaluator != M6_EVAL_NN || nn != nullptr, "evaluator AZ_EVAL_NN needs a network");
    AZ_REQUIRE(precision == AZ_NN_FP32 || precision == AZ_NN_BF16, "unknown precision");
    const az_rules* r = az_env6_rules(env);
    const int T = r->threads_per_mcts < 1 ? 1 : r->threads_per_mcts;
    const int sims = r->mcts_simulations - (r->mcts_simulations % T);
    AZ_REQUIRE(sims >= 1 && sims <= 60000, "mcts_simulations - mcts_simulations % threads_per_mcts must be in 1..60000");
    AzDeviceGuard guard(az_env6_device(env));
    if (evaluator == M6_EVAL_NN && !nn->finalized) { int frc = az_nn_finalize(nn); if (frc) return frc; }
    az_mcts6* mc = new (std::nothrow) az_mcts6();
    AZ_REQUIRE(mc != nullptr, "out of host memory");
    mc->env = env; mc->nn = nn; mc->evaluator = evaluator; mc->precision = precision; mc->device = az_env6_device(env); mc->sims = sims;
    Mcts6Dev& d = mc->d;
    memset(&d, 0, sizeof d);
    d.n = az_env6_n(env);
    d.cap = sims + 2;                                       // the table is cleared per search: the root + one node per simulation
    d.H = 64; while (d.H < 2 * d.cap) d.H <<= 1;
    d.dmax = 192;
    const size_t n = (size_t)d.n;
    int rc = 0;
    rc |= m6_alloc(mc, &d.nodes, n * d.cap * M6_NODE_WORDS, false);
    rc |= m6_alloc(mc, &d.index, n * d.H); rc |= m6_alloc(mc, &d.count, n);
    rc |= m6_alloc(mc, &d.path, n * d.dmax); rc |= m6_alloc(mc, &d.path_len, n);
    rc |= m6_alloc(mc, &d.leaf_state, n * E6_WORDS); rc |= m6_alloc(mc, &d.leaf_valid, n); rc |= m6_alloc(mc, &d.pending, n);
    rc |= m6_alloc(mc, &d.term_value, n); rc |= m6_alloc(mc, &d.term_seat, n);
    rc |= m6_alloc(mc, &d.nn_policy, n * AZ_MOVES); rc |= m6_alloc(mc, &d.nn_value, n);
    rc |= m6_alloc(mc, &d.out_visits, n * AZ_MOVES); rc |= m6_alloc(mc, &d.out_pi, n * AZ_MOVES); rc |= m6_alloc(mc, &d.out_q, n * AZ_MOVES);
    rc |= m6_alloc(mc, &d.out_p, n * AZ_MOVES); rc |= m6_alloc(mc, &d.out_move, n); rc |= m6_alloc(mc, &d.out_status, n);
    rc |= m6_alloc(mc, &d.out_sumn, n); rc |= m6_alloc(mc, &d.out_table, n);
    rc |= m6_alloc(mc, &d.counters, (size_t)M6C_N);
    if (evaluator == M6_EVAL_NN) rc |= m6_alloc(mc, &mc->d_x, n * AZ_INPUT_FLOATS);
    if (rc) { for (void* p : mc->allocs) cudaFree(p); delete mc; return AZ_ERR_CUDA; }
    d.root_state = az_env6_state_ptr(env);
    d.c1 = 1.0f - r->dir_noise_epsi; d.c2 = r->dir_noise_epsi * r->dir_noise_value; d.cpuct = r->cpuct;
    d.rules.allow_yield = r->allow_yield; d.rules.limit_reinforcement = r->limit_reinforcement; d.rules.limit_attack = r->limit_attack;
    d.rules.max_game_rounds = r->max_game_rounds; d.rules.min_unit_move = r->min_unit_move;
    d.eval_mode = evaluator; d.temp_threshold = r->temperature_threshold;
    *out = mc;
    return AZ_OK;
}

extern "C" int az_mcts6_destroy(az_mcts6* mc)
{
    if (!mc) return AZ_OK;
    AzDeviceGuard guard(mc->device);
    for (void* p : mc->allocs) cudaFree(p);
    delete mc;
    return AZ_OK;
}

static int m6_evaluate(az_mcts6* mc, cudaStream_t s)
{
    if (mc->evaluator != M6_EVAL_NN) return AZ_OK;
    Mcts6Dev& d = mc->d;
    k_env6_encode<<<(d.n + 127) / 128, 128, 0, s>>>(d.leaf_state, d.n, mc->d_x);
    AZ_CUDA(cudaGetLastError());
    return az_nn_forward_dev(mc->nn, mc->d_x, d.n, d.nn_policy, d.nn_value, mc->precision, s);
}

static int m6_search_once(az_mcts6* mc, int pick_mode, int apply_move, int auto_reset, cudaStream_t s)
{
    Mcts6Dev& d = mc->d;
    if (mc->evaluator == M6_EVAL_NN && !mc->nn->finalized) { int frc = az_nn_finalize(mc->nn); if (frc) return frc; }
    d.seed = az_env6_seed(mc->env); d.first_game = az_env6_first_game(mc->env);
    const uint64_t* tab = az_device_tables();
    const int grid = (d.n + M6_WARPS - 1) / M6_WARPS;
    k_mcts6_begin<<<grid, M6_WARPS * 32, 0, s>>>(d);
    k_mcts6_sim<<<grid, M6_WARPS * 32, 0, s>>>(d, tab, -1);
    AZ_CUDA(cudaGetLastError());
    int rc = m6_evaluate(mc, s); if (rc) return rc;
    for (int i = 0; i < mc->sims; ++i) {
        k_mcts6_sim<<<grid, M6_WARPS * 32, 0, s>>>(d, tab, i);
        AZ_CUDA(cudaGetLastError());
        rc = m6_evaluate(mc, s); if (rc) return rc;
    }
    k_mcts6_finish<<<grid, M6_WARPS * 32, 0, s>>>(d, tab, pick_mode, apply_move, auto_reset);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_mcts6_search(az_mcts6* mc, int pick_mode, int apply_move, uint32_t* h_visits, float* h_pi, uint8_t* h_move, int8_t* h_status,
                               void* stream)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AZ_REQUIRE(pick_mode == 0 || pick_mode == 1, "pick_mode: 0 = argmax (play), 1 = self-play temperature rule");
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = m6_search_once(mc, pick_mode, apply_move, 0, s); if (rc) return rc;
    const size_t n = (size_t)mc->d.n;
    if (h_visits) AZ_CUDA(cudaMemcpyAsync(h_visits, mc->d.out_visits, sizeof(uint32_t) * n * AZ_MOVES, cudaMemcpyDeviceToHost, s));
    if (h_pi) AZ_CUDA(cudaMemcpyAsync(h_pi, mc->d.out_pi, sizeof(float) * n * AZ_MOVES, cudaMemcpyDeviceToHost, s));
    if (h_move) AZ_CUDA(cudaMemcpyAsync(h_move, mc->d.out_move, n, cudaMemcpyDeviceToHost, s));
    if (h_status) AZ_CUDA(cudaMemcpyAsync(h_status, mc->d.out_status, n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_mcts6_root_stats(az_mcts6* mc, float* h_q, float* h_p, uint32_t* h_sumn, int32_t* h_table, void* stream)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)mc->d.n;
    if (h_q) AZ_CUDA(cudaMemcpyAsync(h_q, mc->d.out_q, sizeof(float) * n * AZ_MOVES, cudaMemcpyDeviceToHost, s));
    if (h_p) AZ_CUDA(cudaMemcpyAsync(h_p, mc->d.out_p, sizeof(float) * n * AZ_MOVES, cudaMemcpyDeviceToHost, s));
    if (h_sumn) AZ_CUDA(cudaMemcpyAsync(h_sumn, mc->d.out_sumn, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
    if (h_table) AZ_CUDA(cudaMemcpyAsync(h_table, mc->d.out_table, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

// n_moves self-play moves of every game with no host synchronisation; finished games are re-dealt in place
extern "C" int az_selfplay6_run(az_mcts6* mc, int n_moves, void* stream)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AZ_REQUIRE(n_moves >= 0, "n_moves must be >= 0");
    AzDeviceGuard guard(mc->device);
    for (int i = 0; i < n_moves; ++i) { int rc = m6_search_once(mc, 1, 1, 1, (cudaStream_t)stream); if (rc) return rc; }
    return AZ_OK;
}

extern "C" int az_mcts6_counters(az_mcts6* mc, az_counters6* h_out, uint64_t* h_sims, uint64_t* h_evals, uint64_t* h_errors, int reset, void* stream)
{
    AZ_REQUIRE(mc && h_out, "NULL argument");
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long h[M6C_N];
    AZ_CUDA(cudaMemcpyAsync(h, mc->d.counters, sizeof h, cudaMemcpyDeviceToHost, s));
    if (reset) AZ_CUDA(cudaMemsetAsync(mc->d.counters, 0, sizeof h, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    h_out->steps = h[M6C_STEPS]; h_out->games = h[M6C_GAMES]; h_out->draws = h[M6C_DRAWS];
    for (int p = 0; p < E6_PLAYERS; ++p) h_out->wins[p] = h[M6C_WIN0 + p];
    if (h_sims) *h_sims = h[M6C_SIMS];
    if (h_evals) *h_evals = h[M6C_EVALS];
    if (h_errors) *h_errors = h[M6C_OVERFLOW];
    return AZ_OK;
}

extern "C" int az_env6_encode(az_env6* e, float* h_x, void* stream)
{
    AZ_REQUIRE(e && h_x, "NULL argument");
    AzDeviceGuard guard(az_env6_device(e));
    cudaStream_t s = (cudaStream_t)stream;
    const int n = az_env6_n(e);
    float* d_x = nullptr;
    AZ_CUDA(cudaMalloc(&d_x, sizeof(float) * (size_t)n * AZ_INPUT_FLOATS));
    k_env6_encode<<<(n + 127) / 128, 128, 0, s>>>(az_env6_state_ptr(e), n, d_x);
    cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_x, d_x, sizeof(float) * (size_t)n * AZ_INPUT_FLOATS, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    cudaFree(d_x);
    if (ce != cudaSuccess) { az_set_error("az_env6_encode: %s", cudaGetErrorString(ce)); return AZ_ERR_CUDA; }
    return AZ_OK;
}
