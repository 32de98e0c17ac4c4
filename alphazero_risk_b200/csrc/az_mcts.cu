// az_mcts.cu — lockstep MCTS over thousands of games for sm_100a, and the az_mcts_* / az_selfplay_* C ABI.
//
// Reference behaviour (player/alpha_zero/alphazero_mcts.cpp): a transposition table keyed by the
// FULL game state (StateSimulationsStorage, :189-245), PUCT selection with a constant "noise"
// prior mix (:67-119), recursive search with chance nodes re-sampled on every descent (:322-377),
// backup Q = (N*Q + v)/(N+1) (:8-21), trim of nodes not visited since the previous search
// (:229-245), root policy from visit counts (:121-148) and move choice (:379-412).
//
// B200 design: one WARP per game.  az_rules.concurrent_descents = K descents per tree are in flight per leaf batch
// (batch = K x games): the lockstep schedule of K search threads — descent j of a round selects after descents
// 0..j-1 and sees their active_N marks (the reference's virtual-loss rule, :91-107), a descent that ends in a
// terminal state backs up at once, the others are evaluated in one batch and then expanded + backed up in order.
// K = 1 is THREADS_PER_MCTS = 1.  Every game owns two node pools + two hash indices in HBM that
// ping-pong per search: nodes created or visited during search k live in pool k&1 (a node found in
// the previous pool is migrated on first touch), so "trimNodes" is an epoch bump + clearing one
// small index, and a node is alive exactly when the reference would still hold it.  A node is one
// 640-byte record (56-byte state key, legal mask, sumN, value, P[43], Q[43], N[43]) read and
// written by the 32 lanes as coalesced rows; PUCT argmax and hash-window probing are warp
// reductions; the float operations use explicit _rn intrinsics (and the TU is built with
// -fmad=false) so every rounding matches the reference's x86 build bit for bit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "az_common.cuh"
#include "az_game.cuh"
#include "az_samples.cuh"
#include "az_nn.cuh"
#include "az_pseudo_net.h"
#include "az_arena.cuh"

#define NODE_WORDS 160
#define NW_VALID 14
#define NW_SUMN 16
#define NW_VALUE 17
#define NW_P 20
#define NW_Q 64
#define NW_N 108
#define MCTS_WARPS 4
#define FULL 0xffffffffu

enum { EVAL_NN = 0, EVAL_PSEUDO = 1, EVAL_UNIFORM = 2 };
#define REC_BYTES AZ_SAMPLE_BYTES
enum { CNT_SIMS = 0, CNT_EVALS = 1, CNT_POOL_OVERFLOW = 2, CNT_DEPTH_OVERFLOW = 3, CNT_STEPS = 4, CNT_GAMES = 5, CNT_W0 = 6, CNT_W1 = 7, CNT_DRAW = 8, CNT_ILLEGAL = 9, CNT_PATH = 10, CNT_POOL_PEAK = 11, CNT_N = 12 };

struct MctsDev {
    int n, cap, H, dmax;
    int g0, gcount;         // the games this launch of k_mcts_sim works on: [g0, g0 + gcount) (a cohort; the whole range otherwise)
    int K;                  // descents per tree that select before any backs up (az_rules.concurrent_descents); slot = j * n + game
    uint32_t* nodes;        // [n][2][cap][NODE_WORDS]
    uint32_t* index;        // [n][2][H]     0 = empty, else tag16 << 16 | (node + 1)
    uint32_t* count;        // [n][2]
    uint32_t* epoch;        // [n]
    uint32_t* migrated;     // [n]
    uint32_t* path;         // [K*n][dmax]   node | move << 16 | flip << 22
    uint32_t* path_len;     // [K*n]
    uint32_t* leaf_state;   // [16][K*n]
    uint64_t* leaf_valid;   // [K*n]
    int32_t* pending;       // [K*n]
    float* term_value;      // [K*n]
    float* nn_policy;       // [K*n][43]
    float* nn_value;        // [K*n]
    uint32_t* root_state;   // env state [16][n]
    uint8_t* extra_trim;    // [n] trims to add before the next search (play-mode turn start / new game)
    const uint8_t* side_sel; int side;   // arena with two searchers: only games with side_sel[game] == side take part in this search (NULL = all)
    uint32_t* out_visits; float* out_pi; float* out_q; float* out_p; uint8_t* out_move; float* out_value; uint32_t* out_sumn; int32_t* out_table; int8_t* out_status;
    unsigned long long* counters;
    // self-play sample recording (NNTrainData, alphazero_nn_data.h:112-121): per-game staging until the game ends
    uint32_t* rec_state;    // [n][rec_moves][14]   root state before the move
    float* rec_pi;          // [n][rec_moves][43]   policy target
    uint32_t* rec_len;      // [n]                  staged samples of the running game (> rec_moves = overflowed)
    uint8_t* rec_out;       // [rec_cap][265]       finished games, packed records in the reference's file layout
    unsigned long long* rec_count;   // [2]         records in rec_out, samples dropped (staging or output full)
    int rec_moves; unsigned long long rec_cap;
    float c1, c2, cpuct;
    uint64_t seed; uint32_t first_game;
    AzRulesDev rules;
    int eval_mode, temp_threshold;
};

struct WarpSmem {
    uint32_t row[MCTS_WARPS][16];       // packed state of the game a warp is working on (land bytes + scalars)
    uint32_t scratch[MCTS_WARPS][12];   // fortify DFS parent bytes
};

struct WG {                // warp-uniform game context: every lane holds the same values
    AzGame g;
    AzLandRow land, scratch;
    uint32_t* row;
};

__device__ __forceinline__ void wg_bind(WG& w, WarpSmem& sm, int warp)
{
    w.row = sm.row[warp];
    w.land.base = (uint8_t*)sm.row[warp];
    w.scratch.base = (uint8_t*)sm.scratch[warp];
}

// state words (SoA [16][n]) -> warp context
__device__ __forceinline__ void wg_load(WG& w, const uint32_t* __restrict__ st, int n, int gi, int lane)
{
    __syncwarp();
    if (lane < 16) w.row[lane] = st[(size_t)lane * n + gi];
    __syncwarp();
    // the four land masks by warp votes: lane l looks at lands l and l + 32 (ncu: the per-lane loop over all 42 lands that
    // az_masks_add_word runs was 17 % of k_mcts_sim's instructions)
    const uint8_t* lb = (const uint8_t*)w.row;
    const uint32_t b0 = lb[lane], b1 = lane < AZ_LANDS - 32 ? lb[32 + lane] : 0x80u;      // 0x80: neutral, no army
    const uint32_t a0 = b0 & 63u, a1 = b1 & 63u;
    w.g.own0 = (uint64_t)__ballot_sync(FULL, (b0 >> 6) == 0u) | ((uint64_t)__ballot_sync(FULL, (b1 >> 6) == 0u) << 32);
    w.g.own1 = (uint64_t)__ballot_sync(FULL, (b0 >> 6) == 1u) | ((uint64_t)__ballot_sync(FULL, (b1 >> 6) == 1u) << 32);
    w.g.gt1 = (uint64_t)__ballot_sync(FULL, a0 > 1u) | ((uint64_t)__ballot_sync(FULL, a1 > 1u) << 32);
    w.g.full = (uint64_t)__ballot_sync(FULL, a0 == AZ_ARMY_MAX) | ((uint64_t)__ballot_sync(FULL, a1 == AZ_ARMY_MAX) << 32);
    az_unpack_scalars(w.g, w.row[10], w.row[11], w.row[12], w.row[13]);
}

// registers -> packed words 10..13 of the row (land bytes are already there)
__device__ __forceinline__ void wg_flush(WG& w, int lane)
{
    __syncwarp();
    if (lane == 0) {
        w.row[10] = (w.row[10] & 0xffffu) | (w.g.cards0 << 16) | (w.g.cards1 << 24);
        w.row[11] = az_pack_w11(w.g); w.row[12] = az_pack_w12(w.g); w.row[13] = az_pack_w13(w.g);
    }
    __syncwarp();
}

__device__ __forceinline__ uint64_t warp_hash(uint32_t kw, int lane)
{
    uint64_t h = ((uint64_t)kw + 0x9E3779B97F4A7C15ull * (uint64_t)(lane + 1)) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    if (lane >= 14) h = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)h, o), hi = __shfl_xor_sync(FULL, (uint32_t)(h >> 32), o);
        h += ((uint64_t)hi << 32) | lo;
    }
    return az_pn_mix(h);
}

__device__ __forceinline__ uint32_t* node_ptr(const MctsDev& m, int gi, uint32_t pool, uint32_t idx)
{
    return m.nodes + (((size_t)gi * 2 + pool) * (size_t)m.cap + idx) * NODE_WORDS;
}

// probe one pool's index for the key held in lanes 0..13 (kw).  Returns node index or -1;
// empty_slot = first empty slot of the probe sequence (where an insert would go).
__device__ __forceinline__ int pool_lookup(const MctsDev& m, int gi, uint32_t pool, uint64_t h, uint32_t kw, int lane, int& empty_slot)
{
    const uint32_t* base = m.index + ((size_t)gi * 2 + pool) * (size_t)m.H;
    const uint32_t tag = (uint32_t)(h >> 48) & 0xffffu;
    const uint32_t start = (uint32_t)h & (uint32_t)(m.H - 1);
    empty_slot = -1;
    for (int win = 0; win < m.H; win += 32) {
        uint32_t slot = (start + (uint32_t)win + (uint32_t)lane) & (uint32_t)(m.H - 1);
        uint32_t e = base[slot];
        uint32_t empties = __ballot_sync(FULL, e == 0);
        uint32_t matches = __ballot_sync(FULL, e != 0 && (e >> 16) == tag);
        int first_empty = empties ? (__ffs((int)empties) - 1) : 32;
        if (first_empty < 32) matches &= (1u << first_empty) - 1u;
        while (matches) {
            int l = __ffs((int)matches) - 1; matches &= matches - 1;
            uint32_t idx = (__shfl_sync(FULL, e, l) & 0xffffu) - 1u;
            const uint32_t* nd = node_ptr(m, gi, pool, idx);
            uint32_t w = lane < 14 ? nd[lane] : 0u;
            if (__ballot_sync(FULL, w == kw) == FULL) return (int)idx;
        }
        if (first_empty < 32) { empty_slot = (int)((start + (uint32_t)win + (uint32_t)first_empty) & (uint32_t)(m.H - 1)); return -1; }
    }
    return -1;
}

__device__ __forceinline__ void index_insert(const MctsDev& m, int gi, uint32_t pool, int slot, uint64_t h, uint32_t idx, int lane)
{
    if (lane == 0) m.index[((size_t)gi * 2 + pool) * (size_t)m.H + slot] = (((uint32_t)(h >> 48) & 0xffffu) << 16) | (idx + 1u);
    __syncwarp();
}

// find the node of the state in w.row: current pool first, then the previous one (migrating it).
// Returns node index in the CURRENT pool or -1 (then ins_slot = where to insert in the current index).
__device__ __forceinline__ int find_node(const MctsDev& m, int gi, uint32_t cur, const WG& w, int lane, uint64_t& h, int& ins_slot)
{
    uint32_t kw = lane < 14 ? w.row[lane] : 0u;
    h = warp_hash(kw, lane);
    int idx = pool_lookup(m, gi, cur, h, kw, lane, ins_slot);
    if (idx >= 0) return idx;
    int dummy;
    int old = pool_lookup(m, gi, cur ^ 1u, h, kw, lane, dummy);
    if (old < 0) return -1;
    uint32_t cnt = m.count[gi * 2 + cur];
    if ((int)cnt >= m.cap || ins_slot < 0) { if (lane == 0) atomicAdd(&m.counters[CNT_POOL_OVERFLOW], 1ull); return -2; }
    const uint32_t* src = node_ptr(m, gi, cur ^ 1u, (uint32_t)old);
    uint32_t* dst = node_ptr(m, gi, cur, cnt);
#pragma unroll
    for (int k = 0; k < NODE_WORDS / 32; ++k) dst[k * 32 + lane] = src[k * 32 + lane];
    __syncwarp();
    if (lane == 0) { m.count[gi * 2 + cur] = cnt + 1; m.migrated[gi] += 1; }
    index_insert(m, gi, cur, ins_slot, h, cnt, lane);
    return (int)cnt;
}

// StateSimulations::getNextBestMoveAndSetVisited, alphazero_mcts.cpp:67-119 (ascending-index iteration = the
// contract).  The N word of a move holds SimulationValue::N in its low 24 bits and active_N in the top byte.  A move
// with N == 0 && active_N == 1 (one descent in flight is about to open it) is passed over (:91-93); if nothing else
// can be chosen the best such move is requested a second time (:108-111).  The chosen move's active_N is incremented.
__device__ __forceinline__ int puct_select(const MctsDev& m, uint32_t* nd, int lane)
{
    const uint64_t valid = (uint64_t)nd[NW_VALID] | ((uint64_t)nd[NW_VALID + 1] << 32);
    const float sq = __fsqrt_rn(__fadd_rn(1.0f, (float)nd[NW_SUMN]));
    float bu = -INFINITY, du = -INFINITY; int bi = 64, di = 64;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        int i = lane + 32 * k;
        if (i < AZ_MOVES && ((valid >> i) & 1ull)) {
            float P = __uint_as_float(nd[NW_P + i]), Q = __uint_as_float(nd[NW_Q + i]);
            const uint32_t nw = nd[NW_N + i], N = nw & 0xffffffu;
            float noiseP = __fadd_rn(__fmul_rn(m.c1, P), m.c2);
            float v = __fmul_rn(__fmul_rn(noiseP, m.cpuct), sq);
            float nn = __fadd_rn(1.0f, (float)N);
            float u = __fadd_rn(Q, __fdiv_rn(v, nn));
            if (N == 0u && (nw >> 24) == 1u) { if (u > du) { du = u; di = i; } }
            else if (u > bu) { bu = u; bi = i; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ou = __shfl_xor_sync(FULL, bu, o); int oi = __shfl_xor_sync(FULL, bi, o);
        if (ou > bu || (ou == bu && oi < bi)) { bu = ou; bi = oi; }
        ou = __shfl_xor_sync(FULL, du, o); oi = __shfl_xor_sync(FULL, di, o);
        if (ou > du || (ou == du && oi < di)) { du = ou; di = oi; }
    }
    if (bi == 64) bi = di;
    if (bi < AZ_MOVES && lane == (bi & 31)) nd[NW_N + bi] += 1u << 24;
    __syncwarp();
    return bi;
}

// StateSimulations::addValue up the recorded path of one descent (AlphaZeroMCTS::search :363-372, SimulationValue::addValue :8-21)
__device__ __forceinline__ void backup_path(const MctsDev& m, int slot, int gi, uint32_t cur, float v, uint32_t len, int lane)
{
    __syncwarp();
    if (len > 0 && lane == 0) {
        for (int d = (int)len - 1; d >= 0; --d) {
            uint32_t e = m.path[(size_t)slot * m.dmax + d];
            uint32_t idx = e & 0xffffu, mv = (e >> 16) & 63u;
            if ((e >> 22) & 1u) v = -v;
            uint32_t* nd = node_ptr(m, gi, cur, idx);
            const uint32_t nw = nd[NW_N + mv], N = nw & 0xffffffu, act = nw >> 24;
            float Q = __uint_as_float(nd[NW_Q + mv]);
            float q = N == 0 ? v : __fdiv_rn(__fadd_rn(__fmul_rn((float)N, Q), v), (float)(N + 1u));
            nd[NW_Q + mv] = __float_as_uint(q);
            nd[NW_N + mv] = (N + 1u) | (((act - 1u) & 0xffu) << 24);
            nd[NW_SUMN] += 1u;
        }
        m.path_len[slot] = 0;
        atomicAdd(&m.counters[CNT_SIMS], 1ull);
        atomicAdd(&m.counters[CNT_PATH], (unsigned long long)len);
    }
    __syncwarp();
}

// part 1 of a simulation step: expand the pending leaf with the evaluator's output
// (StateSimulations ctor :26-42 after NNOutputData::normalize, alphazero_nn_data.cpp:3-27)
// and back the value up the recorded path (AlphaZeroMCTS::search :363-372, addValue :8-21)
// for descent slot = j * n + gi; a state an earlier descent of the round has added is dropped (StateSimulationsStorage::add :203-215)
__device__ __forceinline__ void expand_and_backup(const MctsDev& m, int slot, int gi, uint32_t cur, WG& w, int lane)
{
    const uint32_t len = m.path_len[slot];
    float v = m.term_value[slot];
    if (m.pending[slot]) {
        wg_load(w, m.leaf_state, m.K * m.n, slot, lane);
        const uint64_t valid = m.leaf_valid[slot];
        float p0 = 0.0f, p1 = 0.0f, value;
        if (m.eval_mode == EVAL_NN) {
            p0 = m.nn_policy[(size_t)slot * AZ_MOVES + lane];
            if (lane + 32 < AZ_MOVES) p1 = m.nn_policy[(size_t)slot * AZ_MOVES + 32 + lane];
            value = m.nn_value[slot];
        } else if (m.eval_mode == EVAL_PSEUDO) {
            uint64_t key = az_pn_key((const uint8_t*)w.row, (int)w.g.cur, (int)w.g.round, (int)w.g.phase);
            p0 = az_pn_policy(key, lane);
            if (lane + 32 < AZ_MOVES) p1 = az_pn_policy(key, lane + 32);
            value = az_pn_value(key);
        } else { p0 = 1.0f / 43.0f; p1 = lane + 32 < AZ_MOVES ? 1.0f / 43.0f : 0.0f; value = 0.0f; }
        if (!((valid >> lane) & 1ull)) p0 = 0.0f;
        if (lane + 32 >= AZ_MOVES || !((valid >> (lane + 32)) & 1ull)) p1 = 0.0f;
        // ascending-order fp32 sum over the valid entries, like the reference loop (alphazero_nn_data.cpp:3-27); the invalid ones
        // were set to +0 above and x + 0 == x exactly, so no entry needs a test
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) sum = __fadd_rn(sum, __shfl_sync(FULL, p0, i));
#pragma unroll
        for (int i = 0; i < AZ_MOVES - 32; ++i) sum = __fadd_rn(sum, __shfl_sync(FULL, p1, i));
        if (p0 > 0.0f) p0 = __fdiv_rn(p0, sum);
        if (p1 > 0.0f) p1 = __fdiv_rn(p1, sum);
        uint64_t h; int ins;
        uint32_t kw = lane < 14 ? w.row[lane] : 0u;
        h = warp_hash(kw, lane);
        int found = pool_lookup(m, gi, cur, h, kw, lane, ins);
        uint32_t cnt = m.count[gi * 2 + cur];
        if (found < 0 && (int)cnt < m.cap && ins >= 0) {
            uint32_t* nd = node_ptr(m, gi, cur, cnt);
            if (lane < 14) nd[lane] = kw;
            if (lane == 14) nd[NW_VALID] = (uint32_t)valid;
            if (lane == 15) nd[NW_VALID + 1] = (uint32_t)(valid >> 32);
            if (lane == 16) nd[NW_SUMN] = 0u;
            if (lane == 17) nd[NW_VALUE] = __float_as_uint(value);
            nd[NW_P + lane] = __float_as_uint(p0); nd[NW_Q + lane] = 0u; nd[NW_N + lane] = 0u;
            if (lane < 12) { nd[NW_P + 32 + lane] = __float_as_uint(p1); nd[NW_Q + 32 + lane] = 0u; nd[NW_N + 32 + lane] = 0u; }
            __syncwarp();
            if (lane == 0) m.count[gi * 2 + cur] = cnt + 1;
            index_insert(m, gi, cur, ins, h, cnt, lane);
        } else if (found < 0 && lane == 0) atomicAdd(&m.counters[CNT_POOL_OVERFLOW], 1ull);
        v = value;
        if (lane == 0) { m.pending[slot] = 0; atomicAdd(&m.counters[CNT_EVALS], 1ull); }
    }
    backup_path(m, slot, gi, cur, v, len, lane);
}

// part 2: one descent from the root (AlphaZeroMCTS::search :322-377 unrolled into a loop) for descent slot = j * n + gi.
// sim < 0: only the root lookup of setRootState (:289-307).  A descent that reaches a terminal state backs up here
// (the reference's recursion unwinds at once); one that reaches an unseen state is queued for the evaluator.
__device__ __forceinline__ void descend(const MctsDev& m, int slot, int gi, uint32_t cur, WG& w, const AzTables& T, int sim, int lane)
{
    const size_t ns = (size_t)m.K * m.n;
    wg_load(w, m.root_state, m.n, gi, lane);
    const uint32_t ply = w.row[AZ_W_PLY];
    if (az_game_status(w.g, m.rules) != AZ_STATUS_RUNNING) { if (lane == 0) { m.pending[slot] = 0; m.path_len[slot] = 0; m.term_value[slot] = 0.0f; } return; }
    AzDicePhilox dice; dice.init(m.seed, m.first_game + (uint32_t)gi, ply, (uint32_t)(sim < 0 ? 0 : sim));
    uint32_t depth = 0;
    for (;;) {
        __syncwarp();
        int st = az_game_status(w.g, m.rules);
        if (st != AZ_STATUS_RUNNING) {
            const float tv = st == AZ_STATUS_DRAW ? 0.0f : ((uint32_t)st == w.g.cur ? 1.0f : -1.0f);
            if (lane == 0) { m.term_value[slot] = tv; m.pending[slot] = 0; }
            backup_path(m, slot, gi, cur, tv, depth, lane);
            depth = 0;
            break;
        }
        uint64_t valid = az_valid_moves(w.g, T, m.rules);
        uint64_t h; int ins;
        int idx = find_node(m, gi, cur, w, lane, h, ins);
        if (idx < 0) {                                      // unseen state: queue it for evaluation
            if (lane < 14) m.leaf_state[(size_t)lane * ns + slot] = w.row[lane];
            if (lane == 0) { m.leaf_valid[slot] = valid; m.pending[slot] = idx == -1 ? 1 : 0; m.term_value[slot] = 0.0f; }
            break;
        }
        if (sim < 0) { if (lane == 0) m.pending[slot] = 0; break; }       // setRootState: root already known
        if ((int)depth >= m.dmax) { if (lane == 0) { atomicAdd(&m.counters[CNT_DEPTH_OVERFLOW], 1ull); m.pending[slot] = 0; m.term_value[slot] = 0.0f; } break; }
        uint32_t* nd = node_ptr(m, gi, cur, (uint32_t)idx);
        int mv = puct_select(m, nd, lane);
        uint32_t before = w.g.cur;
        __syncwarp();      // every lane runs the (identical) transition on the shared row: keep them converged
        az_make_move(w.g, w.land, w.scratch, T, m.rules, valid, mv, dice);
        wg_flush(w, lane);
        uint32_t flip = w.g.cur != before ? 1u : 0u;
        if (lane == 0) m.path[(size_t)slot * m.dmax + depth] = (uint32_t)idx | ((uint32_t)mv << 16) | (flip << 22);
        depth++;
    }
    if (lane == 0) m.path_len[slot] = depth;
    __syncwarp();
}


// ---------------------------------------------------------------- self-play samples
// One record = what NNTrainDataStorage::saveTrainingSamples writes per sample (alphazero_nn_data.cpp:115-138):
//   int8 playerIndex | NNInputData (88 B, alphazero_nn_data.h:73-101, g++ x86-64 layout) | float value | float policy[43]
// NNInputData: land[42] @0, playerIndex @42, round u16 @44, then 10 floats @48: reinforcementShare, attackFrequency, canDrawCard,
// phase one-hot x 6, armyShare (alphazero_nn_data.cpp:165-196); the three padding bytes (43, 46, 47) are written as zero.
// Called by the game's warp when the game has ended with `status` (NNTrainDataStorage::updateValues, alphazero_nn_data.cpp:51-65).
__device__ __noinline__ void rec_flush(const MctsDev& m, int gi, int status, uint8_t* s_rec, int lane)
{
    const uint32_t len = m.rec_len[gi];
    if (len == 0) return;
    unsigned long long base = 0;
    bool ok = len <= (uint32_t)m.rec_moves;
    if (lane == 0) ok = az_rec_reserve(m.rec_count, m.rec_cap, len, ok, &base);
    ok = __shfl_sync(FULL, (int)ok, 0) != 0;
    base = ((unsigned long long)__shfl_sync(FULL, (uint32_t)(base >> 32), 0) << 32) | __shfl_sync(FULL, (uint32_t)base, 0);
    if (ok) {
        for (uint32_t i = 0; i < len; ++i) {
            const uint32_t* st = m.rec_state + ((size_t)gi * m.rec_moves + i) * 14;
            const float* pi = m.rec_pi + ((size_t)gi * m.rec_moves + i) * AZ_MOVES;
            __syncwarp();
            az_sample_head(st, status, s_rec, lane);
            for (int k = lane; k < AZ_MOVES; k += 32) {
                const uint32_t u = __float_as_uint(pi[k]);
                for (int b = 0; b < 4; ++b) s_rec[93 + 4 * k + b] = (uint8_t)(u >> (8 * b));
            }
            __syncwarp();
            uint8_t* dst = m.rec_out + (base + i) * REC_BYTES;
            for (int k = lane; k < REC_BYTES; k += 32) dst[k] = s_rec[k];
        }
    }
    __syncwarp();
    if (lane == 0) m.rec_len[gi] = 0;
}

// StateSimulationsStorage::trimNodes (:229-245) as an epoch bump; `extra` additional trims
// (AlphaZeroPlayer::takeTurn's turn-start trim, alphazero_player.cpp:5; clearNodes on newGame :31-34)
__global__ void __launch_bounds__(MCTS_WARPS * 32) k_mcts_begin(MctsDev m, int extra_all)
{
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int gi = blockIdx.x * MCTS_WARPS + warp;
    if (gi >= m.n) return;
    if (m.side_sel && m.side_sel[gi] != (uint8_t)m.side) return;          // the other searcher's game: its table must not age
    uint32_t trims = 1u + (uint32_t)extra_all + (uint32_t)m.extra_trim[gi];
    uint32_t e = m.epoch[gi] + trims;
    for (uint32_t t = 0; t < (trims > 2 ? 2u : trims); ++t) {
        uint32_t pool = (e - t) & 1u;
        uint32_t* ix = m.index + ((size_t)gi * 2 + pool) * (size_t)m.H;
        for (int i = lane; i < m.H; i += 32) ix[i] = 0u;
        if (lane == 0) m.count[gi * 2 + pool] = 0u;
    }
    if (lane == 0) { m.epoch[gi] = e; m.extra_trim[gi] = 0; m.migrated[gi] = 0; }
    if (lane < m.K) { m.pending[lane * m.n + gi] = 0; m.path_len[lane * m.n + gi] = 0; }
}

// one round: complete the K descents of the previous round in order, then start the K descents of this one in order
// (round < 0: setRootState)
__global__ void __launch_bounds__(MCTS_WARPS * 32) k_mcts_sim(MctsDev m, const uint64_t* __restrict__ g_tab, int round, int do_descent)
{
    __shared__ uint64_t s_tab[AZ_TABLE_U64];
    __shared__ WarpSmem s_w;
    for (int i = threadIdx.x; i < AZ_TABLE_U64; i += blockDim.x) s_tab[i] = g_tab[i];
    __syncthreads();
    AzTables T = az_tables_from_smem(s_tab);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int gi = m.g0 + blockIdx.x * MCTS_WARPS + warp;
    if (gi >= m.g0 + m.gcount) return;
    if (m.side_sel && m.side_sel[gi] != (uint8_t)m.side) return;
    WG w; wg_bind(w, s_w, warp);
    const uint32_t cur = m.epoch[gi] & 1u;
    for (int j = 0; j < m.K; ++j) expand_and_backup(m, j * m.n + gi, gi, cur, w, lane);
    if (!do_descent) return;
    if (round < 0) descend(m, gi, gi, cur, w, T, -1, lane);
    else for (int j = 0; j < m.K; ++j) descend(m, j * m.n + gi, gi, cur, w, T, round * m.K + j, lane);
}

// after the last simulation: root statistics (calculateMoveProbability :121-148), move choice
// (pickHigestWeightedMove :397-412 / pickRandomWeightedMove :379-395 under the self-play
// temperature rule, alphazero_trainer.cpp:98-106) and optionally the real move on the env state.
__global__ void __launch_bounds__(MCTS_WARPS * 32) k_mcts_finish(MctsDev m, const uint64_t* __restrict__ g_tab, int pick_mode, int apply_move, int auto_reset)
{
    __shared__ uint64_t s_tab[AZ_TABLE_U64];
    __shared__ WarpSmem s_w;
    __shared__ uint8_t s_rec[MCTS_WARPS][272];
    for (int i = threadIdx.x; i < AZ_TABLE_U64; i += blockDim.x) s_tab[i] = g_tab[i];
    __syncthreads();
    AzTables T = az_tables_from_smem(s_tab);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int gi = blockIdx.x * MCTS_WARPS + warp;
    if (gi >= m.n) return;
    if (m.side_sel && m.side_sel[gi] != (uint8_t)m.side) return;
    WG w; wg_bind(w, s_w, warp);
    const uint32_t cur = m.epoch[gi] & 1u;
    for (int j = 0; j < m.K; ++j) expand_and_backup(m, j * m.n + gi, gi, cur, w, lane);
    wg_load(w, m.root_state, m.n, gi, lane);
    uint32_t ply = w.row[AZ_W_PLY];
    const uint32_t game = m.first_game + (uint32_t)gi;
    int status = az_game_status(w.g, m.rules);
    int move = AZ_NONE;
    if (status == AZ_STATUS_RUNNING) {
        uint64_t h; int ins;
        uint32_t kw = lane < 14 ? w.row[lane] : 0u;
        h = warp_hash(kw, lane);
        int idx = pool_lookup(m, gi, cur, h, kw, lane, ins);
        uint32_t n0 = 0, n1 = 0; float q0 = 0.0f, q1 = 0.0f, pr0 = 0.0f, pr1 = 0.0f, value = 0.0f; uint32_t sumn = 0; uint64_t valid = 0;
        if (idx >= 0) {
            const uint32_t* nd = node_ptr(m, gi, cur, (uint32_t)idx);
            valid = (uint64_t)nd[NW_VALID] | ((uint64_t)nd[NW_VALID + 1] << 32);
            sumn = nd[NW_SUMN]; value = __uint_as_float(nd[NW_VALUE]);
            if ((valid >> lane) & 1ull) { n0 = nd[NW_N + lane] & 0xffffffu; q0 = __uint_as_float(nd[NW_Q + lane]); pr0 = __uint_as_float(nd[NW_P + lane]); }
            if (lane + 32 < AZ_MOVES && ((valid >> (lane + 32)) & 1ull)) { n1 = nd[NW_N + 32 + lane] & 0xffffffu; q1 = __uint_as_float(nd[NW_Q + 32 + lane]); pr1 = __uint_as_float(nd[NW_P + 32 + lane]); }
        }
        float f0 = (float)n0, f1 = (float)n1, sum = 0.0f;
        for (int i = 0; i < AZ_MOVES; ++i) {
            float fi = __shfl_sync(FULL, i < 32 ? f0 : f1, i & 31);
            if ((valid >> i) & 1ull) sum = __fadd_rn(sum, fi);
        }
        float pi0 = __fdiv_rn(f0, sum), pi1 = __fdiv_rn(f1, sum);
        if (!((valid >> lane) & 1ull)) pi0 = __fdiv_rn(0.0f, sum);
        if (lane + 32 >= AZ_MOVES) pi1 = 0.0f;
        bool sample = pick_mode == 1 && (int)w.g.round <= m.temp_threshold;
        az_u32x4 blk = az_rng_block(m.seed, game, ply, AZ_STREAM_REAL, 0);
        if (!sample) {
            float best = 0.0f; int bi = AZ_NONE;
            for (int i = 0; i < AZ_MOVES; ++i) {
                float p = __shfl_sync(FULL, i < 32 ? pi0 : pi1, i & 31);
                if (p > best) { best = p; bi = i; }
            }
            move = bi;
        } else {
            float tot = 0.0f;
            for (int i = 0; i < AZ_MOVES; ++i) tot = __fadd_rn(tot, __shfl_sync(FULL, i < 32 ? pi0 : pi1, i & 31));
            float a = __fmul_rn(tot, az_rng_unit_float(blk.z)), it = 0.0f;
            for (int i = 0; i < AZ_MOVES; ++i) {
                it = __fadd_rn(it, __shfl_sync(FULL, i < 32 ? pi0 : pi1, i & 31));
                if (move == AZ_NONE && it >= a) move = i;
            }
        }
        if (m.out_visits) {
            size_t o = (size_t)gi * AZ_MOVES;
            m.out_visits[o + lane] = n0; m.out_pi[o + lane] = pi0; m.out_q[o + lane] = q0; m.out_p[o + lane] = pr0;
            if (lane + 32 < AZ_MOVES) { m.out_visits[o + 32 + lane] = n1; m.out_pi[o + 32 + lane] = pi1; m.out_q[o + 32 + lane] = q1; m.out_p[o + 32 + lane] = pr1; }
            if (lane == 0) {
                m.out_move[gi] = (uint8_t)move; m.out_value[gi] = value; m.out_sumn[gi] = sumn;
                m.out_table[gi] = (int32_t)(m.count[gi * 2 + cur] + m.count[gi * 2 + (cur ^ 1u)] - m.migrated[gi]);
                atomicMax(&m.counters[CNT_POOL_PEAK], (unsigned long long)m.count[gi * 2 + cur]);     // fullest pool so far (az_mcts_pool_stats)
            }
        }
        if (apply_move && m.rec_out) {
            // threadExecuteTrainingGame pushes (player, NNInputData(rootState), policy) before the move, alphazero_trainer.cpp:108
            const uint32_t k = m.rec_len[gi];
            if (k < (uint32_t)m.rec_moves) {
                if (lane < 14) m.rec_state[((size_t)gi * m.rec_moves + k) * 14 + lane] = w.row[lane];
                float* rp = m.rec_pi + ((size_t)gi * m.rec_moves + k) * AZ_MOVES;
                rp[lane] = pi0; if (lane + 32 < AZ_MOVES) rp[lane + 32] = pi1;
            }
            __syncwarp();
            if (lane == 0) m.rec_len[gi] = k + 1;          // > rec_moves marks the game's samples as overflowed (dropped at the end)
        }
        if (apply_move) {
            uint64_t vm = az_valid_moves(w.g, T, m.rules);
            AzDicePhilox dice; dice.init_with_block0(m.seed, game, ply, AZ_STREAM_REAL, blk);
            __syncwarp();
            int rc = az_make_move(w.g, w.land, w.scratch, T, m.rules, vm, move, dice);
            if (rc == 0) { ply++; if (lane == 0) atomicAdd(&m.counters[CNT_STEPS], 1ull); }
            else if (lane == 0) atomicAdd(&m.counters[CNT_ILLEGAL], 1ull);
            status = az_game_status(w.g, m.rules);
            if (status != AZ_STATUS_RUNNING && lane == 0) {
                atomicAdd(&m.counters[CNT_GAMES], 1ull);
                atomicAdd(&m.counters[status == 0 ? CNT_W0 : (status == 1 ? CNT_W1 : CNT_DRAW)], 1ull);
            }
            if (status != AZ_STATUS_RUNNING && m.rec_out) rec_flush(m, gi, status, s_rec[warp], lane);
            if (status != AZ_STATUS_RUNNING && auto_reset) {
                __syncwarp();
                az_new_game(w.g, w.land, m.seed, game, ply);      // threadExecuteTrainingGame starts the next game with a fresh AlphaZeroMCTS
                if (lane == 0) m.extra_trim[gi] = 2;
                status = AZ_STATUS_RUNNING;
            }
            wg_flush(w, lane);
            if (lane < 14) m.root_state[(size_t)lane * m.n + gi] = w.row[lane];
            if (lane == 14) m.root_state[(size_t)AZ_W_PLY * m.n + gi] = ply;
        }
    } else if (m.out_visits) {
        size_t o = (size_t)gi * AZ_MOVES;
        m.out_visits[o + lane] = 0; m.out_pi[o + lane] = 0.0f; m.out_q[o + lane] = 0.0f; m.out_p[o + lane] = 0.0f;
        if (lane + 32 < AZ_MOVES) { m.out_visits[o + 32 + lane] = 0; m.out_pi[o + 32 + lane] = 0.0f; m.out_q[o + 32 + lane] = 0.0f; m.out_p[o + 32 + lane] = 0.0f; }
        if (lane == 0) { m.out_move[gi] = AZ_NONE; m.out_value[gi] = 0.0f; m.out_sumn[gi] = 0; m.out_table[gi] = 0; }
    }
    if (m.out_status && lane == 0) m.out_status[gi] = (int8_t)status;
}

// arena: a game ended outside this searcher's own move (the opponent's turn, or the other searcher's move): Player::gameFinished ->
// NNTrainDataStorage::updateValues for the samples this searcher staged during the game (alphazero_player.cpp:24-30)
__global__ void __launch_bounds__(MCTS_WARPS * 32) k_mcts_rec_end(MctsDev m, const uint8_t* __restrict__ ended)
{
    __shared__ uint8_t s_rec[MCTS_WARPS][272];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int gi = blockIdx.x * MCTS_WARPS + warp;
    if (gi >= m.n) return;
    const uint32_t e = ended[gi];
    if (e == 0) return;
    rec_flush(m, gi, e == 3 ? AZ_STATUS_DRAW : (int)e - 1, s_rec[warp], lane);
}

// arena leaf batches: compacted position i = j * count + k holds descent slot j * n + games[k]
__global__ void __launch_bounds__(256) k_mcts_gather_leaves(const uint32_t* __restrict__ leaf, int ns, int n, const int32_t* __restrict__ games,
                                                             int count, int K, uint32_t* __restrict__ out)
{
    const int i = blockIdx.x * 256 + threadIdx.x, m = count * K;
    if (i >= m * 16) return;
    const int w = i / m, c = i - w * m, j = c / count, k = c - j * count;
    out[(size_t)w * m + c] = leaf[(size_t)w * ns + (size_t)j * n + games[k]];
}
__global__ void __launch_bounds__(256) k_mcts_scatter_evals(const float* __restrict__ pol_c, const float* __restrict__ val_c, int n,
                                                             const int32_t* __restrict__ games, int count, int K, float* __restrict__ pol,
                                                             float* __restrict__ val)
{
    const int i = blockIdx.x * 256 + threadIdx.x, m = count * K;
    if (i >= m * 44) return;
    const int c = i / 44, e = i - c * 44, j = c / count, k = c - j * count;
    const size_t slot = (size_t)j * n + games[k];
    if (e < 43) pol[slot * 43 + e] = pol_c[(size_t)c * 43 + e];
    else val[slot] = val_c[c];
}
// ascending list of the games with flags[g] == want (one block; a running count over 1024-game strips keeps the order)
__global__ void __launch_bounds__(1024) k_arena_list(const uint8_t* __restrict__ flags, int n, int want, int32_t* __restrict__ out)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g0 = 0; g0 < n; g0 += 1024) {
        const int g = g0 + threadIdx.x;
        const bool on = g < n && flags[g] == (uint8_t)want;
        const unsigned bal = __ballot_sync(FULL, on);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (on) out[before + __popc(bal & ((1u << lane) - 1u))] = g;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- host side
int az_launch_encode(const uint32_t* d_state, int n, float* d_x, cudaStream_t s);   // az_env.cu
uint32_t* az_env_state_ptr(az_env* e); int az_env_n(const az_env* e); int az_env_device(const az_env* e);
uint64_t az_env_seed(const az_env* e); uint32_t az_env_first_game(const az_env* e); const az_rules* az_env_rules(const az_env* e);

struct az_mcts {
    az_env* env = nullptr; az_nn* nn = nullptr;
    int evaluator = EVAL_NN, precision = AZ_NN_FP32, device = 0;
    MctsDev d;
    float* d_x = nullptr;
    // arena: only the slots that take part in a search go through the network (az_arena_play sets these around search_once)
    const int32_t* eval_games = nullptr;   // device list of the participating games, ascending
    int eval_count = 0;                    // its length; 0 = every slot is evaluated
    uint32_t* d_leaf_c = nullptr; float* d_pol_c = nullptr; float* d_val_c = nullptr;   // compacted leaf batch [16][m], [m][43], [m]
    // two game cohorts on two streams (az_mcts_set_cohorts): while one cohort's leaf batch is in the tower, the other cohort's tree
    // kernel, state packing, stem and head tail run, and the pairs of SMs a tower layer's last wave leaves idle go to the other
    // cohort's layer.  Games are independent, so the results do not depend on the cohort count.
    int cohorts = 0;                       // 0 = automatic (two cohorts when every condition of use_cohorts() holds), 1 = off, 2 = on where possible
    cudaStream_t cohort_stream[2] = { nullptr, nullptr };
    cudaEvent_t ev_fork = nullptr, ev_join[2] = { nullptr, nullptr };
    std::vector<void*> allocs;
};

template <class T> static int dalloc(az_mcts* mc, T** p, size_t count, bool zero = true)
{
    AZ_CUDA(cudaMalloc(p, sizeof(T) * count));
    if (zero) AZ_CUDA(cudaMemset(*p, 0, sizeof(T) * count));
    mc->allocs.push_back(*p);
    return AZ_OK;
}

static int next_pow2(int v) { int p = 64; while (p < v) p <<= 1; return p; }

extern "C" int az_mcts_create(az_env* env, az_nn* nn, int evaluator, int precision, az_mcts** out)
{
    AZ_REQUIRE(env && out, "NULL argument");
    AZ_REQUIRE(evaluator >= EVAL_NN && evaluator <= EVAL_UNIFORM, "unknown evaluator");
    AZ_REQUIRE(evaluator != EVAL_NN || nn != nullptr, "evaluator AZ_EVAL_NN needs a network");
    AZ_REQUIRE(precision == AZ_NN_FP32 || precision == AZ_NN_BF16, "unknown precision");
    const az_rules* r = az_env_rules(env);
    int T = r->threads_per_mcts < 1 ? 1 : r->threads_per_mcts;
    int sims = r->mcts_simulations - (r->mcts_simulations % T);          // alphazero_mcts.cpp:265
    AZ_REQUIRE(sims >= 1, "mcts_simulations - mcts_simulations % threads_per_mcts must be >= 1");
    const int K = r->concurrent_descents < 1 ? 1 : r->concurrent_descents;
    AZ_REQUIRE(K <= 16, "concurrent_descents must be <= 16");
    AZ_REQUIRE(sims % K == 0, "the simulation count (mcts_simulations - mcts_simulations % threads_per_mcts) must be a multiple of concurrent_descents");
    AzDeviceGuard guard(az_env_device(env));
    if (evaluator == EVAL_NN && !nn->finalized) { int frc = az_nn_finalize(nn); if (frc) return frc; }   // weights loaded but not yet folded / packed
    az_mcts* mc = new (std::nothrow) az_mcts();
    AZ_REQUIRE(mc != nullptr, "out of host memory");
    mc->env = env; mc->nn = nn; mc->evaluator = evaluator; mc->precision = precision; mc->device = az_env_device(env);
    MctsDev& d = mc->d;
    memset(&d, 0, sizeof d);
    d.n = az_env_n(env); d.K = K;
    d.g0 = 0; d.gcount = d.n;
    d.cap = 3 * (sims + 1) + 64; if (d.cap > 65000) d.cap = 65000;
    d.H = next_pow2(2 * d.cap);
    d.dmax = 192;
    size_t n = (size_t)d.n, ns = n * (size_t)K;          // ns = descent slots = leaf batch
    int rc = 0;
    rc |= dalloc(mc, &d.nodes, n * 2 * d.cap * NODE_WORDS, false);
    rc |= dalloc(mc, &d.index, n * 2 * d.H);
    rc |= dalloc(mc, &d.count, n * 2); rc |= dalloc(mc, &d.epoch, n); rc |= dalloc(mc, &d.migrated, n);
    rc |= dalloc(mc, &d.path, ns * d.dmax); rc |= dalloc(mc, &d.path_len, ns);
    rc |= dalloc(mc, &d.leaf_state, ns * 16); rc |= dalloc(mc, &d.leaf_valid, ns); rc |= dalloc(mc, &d.pending, ns);
    rc |= dalloc(mc, &d.term_value, ns); rc |= dalloc(mc, &d.nn_policy, ns * AZ_MOVES); rc |= dalloc(mc, &d.nn_value, ns);
    rc |= dalloc(mc, &d.extra_trim, n);
    rc |= dalloc(mc, &d.out_visits, n * AZ_MOVES); rc |= dalloc(mc, &d.out_pi, n * AZ_MOVES); rc |= dalloc(mc, &d.out_q, n * AZ_MOVES);
    rc |= dalloc(mc, &d.out_p, n * AZ_MOVES); rc |= dalloc(mc, &d.out_move, n); rc |= dalloc(mc, &d.out_value, n);
    rc |= dalloc(mc, &d.out_sumn, n); rc |= dalloc(mc, &d.out_table, n); rc |= dalloc(mc, &d.out_status, n);
    rc |= dalloc(mc, &d.counters, (size_t)CNT_N);
    if (evaluator == EVAL_NN) {
        rc |= dalloc(mc, &mc->d_x, ns * AZ_INPUT_FLOATS);
        rc |= dalloc(mc, &mc->d_leaf_c, ns * 16); rc |= dalloc(mc, &mc->d_pol_c, ns * AZ_MOVES); rc |= dalloc(mc, &mc->d_val_c, ns);
    }
    if (rc) { for (void* p : mc->allocs) cudaFree(p); delete mc; return AZ_ERR_CUDA; }
    d.root_state = az_env_state_ptr(env);
    d.c1 = 1.0f - r->dir_noise_epsi;                 // (1 - SETTINGS.DIR_NOISE_EPSI), alphazero_mcts.cpp:81
    d.c2 = r->dir_noise_epsi * r->dir_noise_value;   // SETTINGS.DIR_NOISE_EPSI * SETTINGS.DIR_NOISE_VALUE
    d.cpuct = r->cpuct;
    d.rules.allow_yield = r->allow_yield; d.rules.limit_reinforcement = r->limit_reinforcement; d.rules.limit_attack = r->limit_attack;
    d.rules.max_game_rounds = r->max_game_rounds; d.rules.min_unit_move = r->min_unit_move;
    d.eval_mode = evaluator; d.temp_threshold = r->temperature_threshold;
    *out = mc;
    return AZ_OK;
}

extern "C" int az_mcts_destroy(az_mcts* mc)
{
    if (!mc) return AZ_OK;
    AzDeviceGuard guard(mc->device);
    for (void* p : mc->allocs) cudaFree(p);
    for (int c = 0; c < 2; ++c) { if (mc->cohort_stream[c]) cudaStreamDestroy(mc->cohort_stream[c]); if (mc->ev_join[c]) cudaEventDestroy(mc->ev_join[c]); }
    if (mc->ev_fork) cudaEventDestroy(mc->ev_fork);
    delete mc;
    return AZ_OK;
}

extern "C" int az_mcts_set_cohorts(az_mcts* mc, int cohorts)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AZ_REQUIRE(cohorts >= 0 && cohorts <= 2, "cohorts: 0 = automatic, 1 = one stream, 2 = two game cohorts on two streams");
    mc->cohorts = cohorts;
    return AZ_OK;
}

extern "C" int az_mcts_simulations(const az_mcts* mc)
{
    if (!mc) return 0;
    const az_rules* r = az_env_rules(mc->env);
    int T = r->threads_per_mcts < 1 ? 1 : r->threads_per_mcts;
    return r->mcts_simulations - (r->mcts_simulations % T);
}

// StateSimulationsStorage::clearNodes for every game (AlphaZeroPlayer::newGame, alphazero_player.cpp:31-34)
extern "C" int az_mcts_clear(az_mcts* mc, void* stream)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AzDeviceGuard guard(mc->device);
    AZ_CUDA(cudaMemsetAsync(mc->d.extra_trim, 2, (size_t)mc->d.n, (cudaStream_t)stream));
    return AZ_OK;
}

// k_used = descents per tree whose leaves this batch holds: K for a simulation round, 1 for the root round (setRootState uses
// descent slot 0 only, so the other K - 1 slots of every game would be evaluated for nothing)
// the leaf batch of ONE cohort (games [g0, g0 + gcount), one descent per tree): the cohort's slice of the SoA leaf arrays goes
// through the tower with the cohort's own set of work buffers, on the cohort's stream
static int evaluate_cohort(az_mcts* mc, cudaStream_t s, int cohort, int g0, int gcount)
{
    MctsDev& d = mc->d;
    return az_nn_tc_forward(mc->nn, nullptr, d.leaf_state + g0, gcount, d.nn_policy + (size_t)g0 * AZ_MOVES, d.nn_value + g0, s, d.n * d.K, cohort, true);
}

static int evaluate_leaves(az_mcts* mc, cudaStream_t s, int k_used)
{
    if (mc->evaluator != EVAL_NN) return AZ_OK;
    MctsDev& d = mc->d;
    const int ns = d.n * d.K;                            // descent slots (slot = j * n + game): the stride of the SoA leaf arrays
    const int nb = d.n * k_used;                         // the first k_used * n of them go through the network
    if (mc->eval_count > 0 && mc->eval_count < d.n) {
        // arena: gather the participating games' leaves, evaluate that batch, scatter the outputs back to their slots
        const int m = mc->eval_count * k_used;
        k_mcts_gather_leaves<<<(m * 16 + 255) / 256, 256, 0, s>>>(d.leaf_state, ns, d.n, mc->eval_games, mc->eval_count, k_used, mc->d_leaf_c);
        int rc;
        if (mc->precision == AZ_NN_BF16) {
            rc = az_nn_reserve(mc->nn, ns); if (rc) return rc;
            rc = az_nn_tc_forward(mc->nn, nullptr, mc->d_leaf_c, m, mc->d_pol_c, mc->d_val_c, s); if (rc) return rc;
        } else {
            rc = az_launch_encode(mc->d_leaf_c, m, mc->d_x, s); if (rc) return rc;
            rc = az_nn_forward_dev(mc->nn, mc->d_x, m, mc->d_pol_c, mc->d_val_c, AZ_NN_FP32, s); if (rc) return rc;
        }
        k_mcts_scatter_evals<<<(m * 44 + 255) / 256, 256, 0, s>>>(mc->d_pol_c, mc->d_val_c, d.n, mc->eval_games, mc->eval_count, k_used, d.nn_policy, d.nn_value);
        AZ_CUDA(cudaGetLastError());
        return AZ_OK;
    }
    if (mc->precision == AZ_NN_BF16) {
        int rc = az_nn_reserve(mc->nn, ns); if (rc) return rc;
        return az_nn_tc_forward(mc->nn, nullptr, d.leaf_state, nb, d.nn_policy, d.nn_value, s, ns);
    }
    int rc = az_launch_encode(d.leaf_state, ns, mc->d_x, s); if (rc) return rc;      // (encodes every slot: the stride of the SoA input is ns)
    return az_nn_forward_dev(mc->nn, mc->d_x, nb, d.nn_policy, d.nn_value, AZ_NN_FP32, s);
}

// Two cohorts pay when the tower dominates and a half batch still fills the device: tensor-core evaluator, one descent per tree (a
// cohort's leaves are then one contiguous slice), every game searching (no arena compaction / side selection), and at least 2048
// games (1024 per cohort = 192 tile pairs = 2.6 waves per layer; the other cohort's layer fills the tail)
static bool use_cohorts(const az_mcts* mc)
{
    if (mc->cohorts == 1) return false;
    const MctsDev& d = mc->d;
    const bool possible = mc->evaluator == EVAL_NN && mc->precision == AZ_NN_BF16 && d.K == 1 && mc->eval_count == 0 && d.side_sel == nullptr && d.n >= 8;
    if (mc->cohorts == 2) return possible;
    return possible && d.n >= 2048;
}

// one AlphaZeroMCTS::simulate (alphazero_mcts.cpp:255-287) for every game, then the move choice
static int search_once(az_mcts* mc, int extra_all, int pick_mode, int apply_move, int auto_reset, cudaStream_t s)
{
    MctsDev& d = mc->d;
    // the network may have been trained, restored or overwritten since the last search: fold + pack the current weights first
    if (mc->evaluator == EVAL_NN && !mc->nn->finalized) { int frc = az_nn_finalize(mc->nn); if (frc) return frc; }
    d.seed = az_env_seed(mc->env); d.first_game = az_env_first_game(mc->env);
    d.g0 = 0; d.gcount = d.n;
    const uint64_t* tab = az_device_tables();
    int grid = (d.n + MCTS_WARPS - 1) / MCTS_WARPS, sims = az_mcts_simulations(mc);
    k_mcts_begin<<<grid, MCTS_WARPS * 32, 0, s>>>(d, extra_all);
    AZ_CUDA(cudaGetLastError());
    if (use_cohorts(mc)) {
        if (!mc->ev_fork) {
            AZ_CUDA(cudaEventCreateWithFlags(&mc->ev_fork, cudaEventDisableTiming));
            for (int c = 0; c < 2; ++c) {
                AZ_CUDA(cudaStreamCreateWithFlags(&mc->cohort_stream[c], cudaStreamNonBlocking));
                AZ_CUDA(cudaEventCreateWithFlags(&mc->ev_join[c], cudaEventDisableTiming));
            }
        }
        int rc = az_nn_reserve(mc->nn, d.n); if (rc) return rc;
        const int split = (d.n / 2 + MCTS_WARPS - 1) / MCTS_WARPS * MCTS_WARPS;      // cohort 0 = games [0, split), cohort 1 = the rest
        const int g0[2] = { 0, split }, gc[2] = { split, d.n - split };
        AZ_CUDA(cudaEventRecord(mc->ev_fork, s));
        for (int c = 0; c < 2; ++c) AZ_CUDA(cudaStreamWaitEvent(mc->cohort_stream[c], mc->ev_fork, 0));
        for (int i = -1; i < sims; ++i) {                 // round -1 = setRootState; the cohorts' launches alternate so both streams stay fed
            for (int c = 0; c < 2; ++c) {
                MctsDev dc = d; dc.g0 = g0[c]; dc.gcount = gc[c];
                k_mcts_sim<<<(gc[c] + MCTS_WARPS - 1) / MCTS_WARPS, MCTS_WARPS * 32, 0, mc->cohort_stream[c]>>>(dc, tab, i, 1);
                AZ_CUDA(cudaGetLastError());
                rc = evaluate_cohort(mc, mc->cohort_stream[c], c, g0[c], gc[c]); if (rc) return rc;
            }
        }
        for (int c = 0; c < 2; ++c) {
            AZ_CUDA(cudaEventRecord(mc->ev_join[c], mc->cohort_stream[c]));
            AZ_CUDA(cudaStreamWaitEvent(s, mc->ev_join[c], 0));
        }
    } else {
        k_mcts_sim<<<grid, MCTS_WARPS * 32, 0, s>>>(d, tab, -1, 1);             // setRootState
        AZ_CUDA(cudaGetLastError());
        int rc = evaluate_leaves(mc, s, 1); if (rc) return rc;
        for (int i = 0; i < sims / d.K; ++i) {               // rounds of K descents per tree
            k_mcts_sim<<<grid, MCTS_WARPS * 32, 0, s>>>(d, tab, i, 1);
            AZ_CUDA(cudaGetLastError());
            rc = evaluate_leaves(mc, s, d.K); if (rc) return rc;
        }
    }
    k_mcts_finish<<<grid, MCTS_WARPS * 32, 0, s>>>(d, tab, pick_mode, apply_move, auto_reset);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_mcts_search(az_mcts* mc, const uint8_t* h_extra_trim, int pick_mode, int apply_move,
                              uint32_t* h_visits, float* h_pi, uint8_t* h_move, int8_t* h_status, void* stream)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AZ_REQUIRE(pick_mode == 0 || pick_mode == 1, "pick_mode: 0 = argmax (play), 1 = self-play temperature rule");
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    MctsDev& d = mc->d;
    if (h_extra_trim) {
        // added to (not replacing) trims already scheduled on the device, e.g. by az_mcts_clear
        std::vector<uint8_t> cur(d.n);
        AZ_CUDA(cudaMemcpyAsync(cur.data(), d.extra_trim, (size_t)d.n, cudaMemcpyDeviceToHost, s));
        AZ_CUDA(cudaStreamSynchronize(s));
        for (int i = 0; i < d.n; ++i) cur[i] = (uint8_t)(cur[i] + h_extra_trim[i]);
        AZ_CUDA(cudaMemcpyAsync(d.extra_trim, cur.data(), (size_t)d.n, cudaMemcpyHostToDevice, s));
        AZ_CUDA(cudaStreamSynchronize(s));
    }
    int rc = search_once(mc, 0, pick_mode, apply_move, 0, s); if (rc) return rc;
    size_t n = (size_t)d.n;
    if (h_visits) AZ_CUDA(cudaMemcpyAsync(h_visits, d.out_visits, sizeof(uint32_t) * n * AZ_MOVES, cudaMemcpyDeviceToHost, s));
    if (h_pi) AZ_CUDA(cudaMemcpyAsync(h_pi, d.out_pi, sizeof(float) * n * AZ_MOVES, cudaMemcpyDeviceToHost, s));
    if (h_move) AZ_CUDA(cudaMemcpyAsync(h_move, d.out_move, n, cudaMemcpyDeviceToHost, s));
    if (h_status) AZ_CUDA(cudaMemcpyAsync(h_status, d.out_status, n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_mcts_root_stats(az_mcts* mc, float* h_q, float* h_p, uint32_t* h_sumn, float* h_value, int32_t* h_table, void* stream)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    MctsDev& d = mc->d; size_t n = (size_t)d.n;
    if (h_q) AZ_CUDA(cudaMemcpyAsync(h_q, d.out_q, sizeof(float) * n * AZ_MOVES, cudaMemcpyDeviceToHost, s));
    if (h_p) AZ_CUDA(cudaMemcpyAsync(h_p, d.out_p, sizeof(float) * n * AZ_MOVES, cudaMemcpyDeviceToHost, s));
    if (h_sumn) AZ_CUDA(cudaMemcpyAsync(h_sumn, d.out_sumn, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
    if (h_value) AZ_CUDA(cudaMemcpyAsync(h_value, d.out_value, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
    if (h_table) AZ_CUDA(cudaMemcpyAsync(h_table, d.out_table, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

// n_moves lockstep self-play moves (threadExecuteTrainingGame, alphazero_trainer.cpp:80-119) with no
// host synchronisation: search, temperature-rule move choice, real move, finished games re-dealt.
extern "C" int az_selfplay_run(az_mcts* mc, int n_moves, void* stream)
{
    AZ_REQUIRE(mc != nullptr && n_moves >= 0, "bad argument");
    AzDeviceGuard guard(mc->device);
    for (int i = 0; i < n_moves; ++i) {
        int rc = search_once(mc, 0, 1, 1, 1, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return AZ_OK;
}

// ---- self-play samples
extern "C" int az_selfplay_record(az_mcts* mc, size_t capacity_samples, int max_moves_per_game)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AZ_REQUIRE(capacity_samples > 0 && max_moves_per_game > 0, "capacity_samples and max_moves_per_game must be positive");
    AZ_REQUIRE(mc->d.rec_out == nullptr, "recording is already enabled on this handle");
    AzDeviceGuard guard(mc->device);
    MctsDev& d = mc->d;
    size_t n = (size_t)d.n;
    int rc = 0;
    rc |= dalloc(mc, &d.rec_state, n * (size_t)max_moves_per_game * 14, false);
    rc |= dalloc(mc, &d.rec_pi, n * (size_t)max_moves_per_game * AZ_MOVES, false);
    rc |= dalloc(mc, &d.rec_len, n);
    rc |= dalloc(mc, &d.rec_count, (size_t)2);
    uint8_t* out = nullptr;
    rc |= dalloc(mc, &out, capacity_samples * (size_t)REC_BYTES, false);
    if (rc) { d.rec_out = nullptr; return AZ_ERR_CUDA; }
    d.rec_moves = max_moves_per_game; d.rec_cap = capacity_samples;
    d.rec_out = out;                                   // set last: the kernels record iff rec_out != NULL
    return AZ_OK;
}

extern "C" int az_selfplay_samples(az_mcts* mc, uint8_t* h_records, size_t max_records, size_t* n_out, uint64_t* h_dropped, void* stream)
{
    AZ_REQUIRE(mc && n_out, "NULL argument");
    AZ_REQUIRE(mc->d.rec_out != nullptr, "az_selfplay_record has not been called");
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long h[2];
    AZ_CUDA(cudaMemcpyAsync(h, mc->d.rec_count, sizeof h, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    AZ_REQUIRE(h[0] <= mc->d.rec_cap, "sample queue corrupted: committed count exceeds the capacity");
    const size_t have = (size_t)h[0];                  // committed records: az_rec_reserve never reserves past the capacity
    if (h_dropped) *h_dropped = h[1];
    *n_out = have;
    if (!h_records && have) return AZ_OK;              // size query; with an empty queue the call drains (resets the dropped count)
    AZ_REQUIRE(max_records >= have, "h_records is too small: query the size with h_records = NULL first");
    if (have) AZ_CUDA(cudaMemcpyAsync(h_records, mc->d.rec_out, have * (size_t)REC_BYTES, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemsetAsync(mc->d.rec_count, 0, sizeof h, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

// NNTrainDataStorage::saveTrainingSamples (alphazero_nn_data.cpp:115-138): size_t count, then the packed records
extern "C" int az_samples_write_file(const char* path, const uint8_t* h_records, size_t n)
{
    AZ_REQUIRE(path && (h_records || n == 0), "NULL argument");
    FILE* f = fopen(path, "wb");
    if (!f) { az_set_error("cannot open %s for writing", path); return AZ_ERR_INVALID_ARG; }
    uint64_t count = (uint64_t)n;                      // sizeof(size_t) == 8 on the reference's platform
    bool ok = fwrite(&count, sizeof count, 1, f) == 1 && (n == 0 || fwrite(h_records, REC_BYTES, n, f) == n);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { az_set_error("short write to %s", path); return AZ_ERR_INVALID_ARG; }
    return AZ_OK;
}

extern "C" int az_mcts_counters(az_mcts* mc, az_counters* h_out, uint64_t* h_errors, int reset, void* stream)
{
    AZ_REQUIRE(mc && h_out, "NULL argument");
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long h[CNT_N];
    AZ_CUDA(cudaMemcpyAsync(h, mc->d.counters, sizeof h, cudaMemcpyDeviceToHost, s));
    if (reset) AZ_CUDA(cudaMemsetAsync(mc->d.counters, 0, sizeof h, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    h_out->sims = h[CNT_SIMS]; h_out->evals = h[CNT_EVALS]; h_out->steps = h[CNT_STEPS]; h_out->games = h[CNT_GAMES];
    h_out->wins[0] = h[CNT_W0]; h_out->wins[1] = h[CNT_W1]; h_out->draws = h[CNT_DRAW]; h_out->illegal = h[CNT_ILLEGAL]; h_out->path_nodes = h[CNT_PATH];
    if (h_errors) *h_errors = h[CNT_POOL_OVERFLOW] + h[CNT_DEPTH_OVERFLOW];
    return AZ_OK;
}

// How full the node pools get: peak = the largest number of nodes any game's current pool held at the end of a search since the
// counters were last reset, capacity = nodes per pool (3 x (simulations + 1) + 64; a pool that would overflow is counted in
// az_mcts_counters' h_errors).  The pools are sized for the worst case; this is the measured occupancy next to it.
extern "C" int az_mcts_pool_stats(az_mcts* mc, uint64_t* h_peak_nodes, uint64_t* h_capacity_nodes, uint64_t* h_bytes_per_game, void* stream)
{
    AZ_REQUIRE(mc != nullptr, "mcts is NULL");
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long peak = 0;
    AZ_CUDA(cudaMemcpyAsync(&peak, mc->d.counters + CNT_POOL_PEAK, sizeof peak, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    if (h_peak_nodes) *h_peak_nodes = peak;
    if (h_capacity_nodes) *h_capacity_nodes = (uint64_t)mc->d.cap;
    if (h_bytes_per_game) *h_bytes_per_game = 2ull * (uint64_t)mc->d.cap * NODE_WORDS * 4ull + 2ull * (uint64_t)mc->d.H * 4ull;
    return AZ_OK;
}

// ---------------------------------------------------------------- arena: -m play on the device (SURVEY §8f N1 + N2)
// executePlay (src/alphazero_risk.cpp:4-47) -> GameGroup::playGames (game/game.cpp:277-312): every game slot is what one
// threadPlayGame thread is in the reference — player index 0 = AlphaZeroPlayer (searched here by the lockstep MCTS in play mode:
// argmax move, table trimmed when its turn starts, cleared at a new game), player index 1 = the opponent, games claimed in
// mirror pairs, results tallied like GameResults.
int az_env_reset(az_env* e, uint64_t seed, void* stream);

struct az_arena {
    az_mcts* mc = nullptr;
    az_mcts* opp = nullptr;          // AZ_OPPONENT_ALPHAZERO: the searcher of player index 1
    int32_t* d_list[2] = { nullptr, nullptr };   // games waiting for searcher 0 / 1 this tick (leaf-batch compaction)
    ArenaDev a;
    std::vector<void*> allocs;
};

template <class T> static int aalloc(az_arena* ar, T** p, size_t count)
{
    AZ_CUDA(cudaMalloc(p, sizeof(T) * count));
    AZ_CUDA(cudaMemset(*p, 0, sizeof(T) * count));
    ar->allocs.push_back(*p);
    return AZ_OK;
}

extern "C" int az_arena_create(az_mcts* mc, int opponent, int mirror_games, az_arena** out)
{
    AZ_REQUIRE(mc && out, "NULL argument");
    AZ_REQUIRE(opponent == AZ_OPPONENT_SCRIPT || opponent == AZ_OPPONENT_RANDOM, "unknown opponent kind");
    AzDeviceGuard guard(mc->device);
    az_arena* ar = new (std::nothrow) az_arena();
    AZ_REQUIRE(ar != nullptr, "out of host memory");
    ar->mc = mc;
    ArenaDev& a = ar->a;
    memset(&a, 0, sizeof a);
    a.n = mc->d.n; a.opponent = opponent; a.mirror = mirror_games ? 1 : 0;
    size_t n = (size_t)a.n;
    int rc = 0;
    rc |= aalloc(ar, &a.start_state, n * 16); rc |= aalloc(ar, &a.script, n * 2);
    rc |= aalloc(ar, &a.player_start, n); rc |= aalloc(ar, &a.fresh, n); rc |= aalloc(ar, &a.active, n); rc |= aalloc(ar, &a.last_mover, n);
    rc |= aalloc(ar, &a.ended, n);
    rc |= aalloc(ar, &a.res, (size_t)ARENA_N);
    rc |= aalloc(ar, &ar->d_list[0], n);
    if (rc) { for (void* p : ar->allocs) cudaFree(p); delete ar; return AZ_ERR_CUDA; }
    a.state = mc->d.root_state; a.extra_trim = mc->d.extra_trim;
    *out = ar;
    return AZ_OK;
}

// AlphaZero vs AlphaZero (AlphaZeroTrainer::updateIfImprovement's comparison match): two searchers over one set of game states.
// k_arena_advance hands every running slot to the side to move (to_move); each tick runs one search per side restricted to its
// slots (MctsDev::side_sel), so a searcher's table ages exactly when its AlphaZeroPlayer would search.
extern "C" int az_arena_create_versus(az_mcts* mc, az_mcts* opp, int mirror_games, az_arena** out)
{
    AZ_REQUIRE(mc && opp && out, "NULL argument");
    AZ_REQUIRE(mc != opp, "the two searchers must be different handles (each AlphaZeroPlayer owns its table)");
    AZ_REQUIRE(mc->env == opp->env, "both searchers must be built over the same env");
    AzDeviceGuard guard(mc->device);
    az_arena* ar = new (std::nothrow) az_arena();
    AZ_REQUIRE(ar != nullptr, "out of host memory");
    ar->mc = mc; ar->opp = opp;
    ArenaDev& a = ar->a;
    memset(&a, 0, sizeof a);
    a.n = mc->d.n; a.opponent = AZ_OPPONENT_ALPHAZERO; a.mirror = mirror_games ? 1 : 0;
    size_t n = (size_t)a.n;
    int rc = 0;
    rc |= aalloc(ar, &a.start_state, n * 16); rc |= aalloc(ar, &a.script, n * 2);
    rc |= aalloc(ar, &a.player_start, n); rc |= aalloc(ar, &a.fresh, n); rc |= aalloc(ar, &a.active, n); rc |= aalloc(ar, &a.last_mover, n);
    rc |= aalloc(ar, &a.to_move, n); rc |= aalloc(ar, &a.ended, n);
    rc |= aalloc(ar, &a.res, (size_t)ARENA_N);
    rc |= aalloc(ar, &ar->d_list[0], n); rc |= aalloc(ar, &ar->d_list[1], n);
    if (rc) { for (void* p : ar->allocs) cudaFree(p); delete ar; return AZ_ERR_CUDA; }
    a.state = mc->d.root_state; a.extra_trim = mc->d.extra_trim; a.extra_trim_opp = opp->d.extra_trim;
    *out = ar;
    return AZ_OK;
}

extern "C" int az_arena_destroy(az_arena* ar)
{
    if (!ar) return AZ_OK;
    AzDeviceGuard guard(ar->mc->device);
    for (void* p : ar->allocs) cudaFree(p);
    delete ar;
    return AZ_OK;
}

extern "C" int az_arena_play(az_arena* ar, uint64_t n_games, uint64_t seed, az_arena_results* h_out, void* stream)
{
    AZ_REQUIRE(ar && h_out, "NULL argument");
    az_mcts* mc = ar->mc;
    AzDeviceGuard guard(mc->device);
    cudaStream_t s = (cudaStream_t)stream;
    ArenaDev& a = ar->a;
    size_t n = (size_t)a.n;
    int rc = az_env_reset(mc->env, seed, stream); if (rc) return rc;           // fixes the seed of the Philox contract for this match
    a.seed = seed; a.first_game = az_env_first_game(mc->env); a.total_games = n_games;
    // ScriptPlayer objects live as long as the PlayerGroup: their members are NOT reset between matches in the reference either,
    // but a match here starts from fresh players (AZ_SCRIPT_INIT = "never set")
    std::vector<uint32_t> init(n * 2, 0x00ffffffu);
    AZ_CUDA(cudaMemcpyAsync(a.script, init.data(), sizeof(uint32_t) * n * 2, cudaMemcpyHostToDevice, s));
    AZ_CUDA(cudaMemsetAsync(a.player_start, 0, n, s)); AZ_CUDA(cudaMemsetAsync(a.fresh, 1, n, s));
    AZ_CUDA(cudaMemsetAsync(a.active, 1, n, s)); AZ_CUDA(cudaMemsetAsync(a.last_mover, 0xff, n, s));
    AZ_CUDA(cudaMemsetAsync(a.res, 0, sizeof(unsigned long long) * ARENA_N, s));
    AZ_CUDA(cudaMemsetAsync(mc->d.counters, 0, sizeof(unsigned long long) * CNT_N, s));
    az_mcts* opp = ar->opp;
    if (mc->d.rec_out) AZ_CUDA(cudaMemsetAsync(mc->d.rec_len, 0, sizeof(uint32_t) * n, s));          // a match starts with nothing staged
    if (opp && opp->d.rec_out) AZ_CUDA(cudaMemsetAsync(opp->d.rec_len, 0, sizeof(uint32_t) * n, s));
    if (opp) {
        AZ_CUDA(cudaMemsetAsync(opp->d.counters, 0, sizeof(unsigned long long) * CNT_N, s));
        AZ_CUDA(cudaMemsetAsync(a.to_move, 0xff, n, s));
    }
    AZ_CUDA(cudaStreamSynchronize(s));                                        // `init` goes out of scope below
    unsigned long long act[3] = { 0, 0, 0 };                                  // ARENA_ACTIVE, ARENA_TOMOVE0, ARENA_TOMOVE1
    uint64_t ticks = 0;
    for (;;) {
        AZ_CUDA(cudaMemsetAsync(a.res + ARENA_ACTIVE, 0, 3 * sizeof(unsigned long long), s));
        const bool recording = mc->d.rec_out || (opp && opp->d.rec_out);
        if (recording) AZ_CUDA(cudaMemsetAsync(a.ended, 0, n, s));
        rc = az_launch_arena_advance(a, az_env_rules(mc->env), s); if (rc) return rc;
        if (recording) {                                                      // games that ended: value targets for the staged samples
            const int rgrid = (a.n + MCTS_WARPS - 1) / MCTS_WARPS;
            if (mc->d.rec_out) k_mcts_rec_end<<<rgrid, MCTS_WARPS * 32, 0, s>>>(mc->d, a.ended);
            if (opp && opp->d.rec_out) k_mcts_rec_end<<<rgrid, MCTS_WARPS * 32, 0, s>>>(opp->d, a.ended);
            AZ_CUDA(cudaGetLastError());
        }
        AZ_CUDA(cudaMemcpyAsync(act, a.res + ARENA_ACTIVE, sizeof act, cudaMemcpyDeviceToHost, s));
        AZ_CUDA(cudaStreamSynchronize(s));
        if (act[0] == 0) break;
        // leaf batches hold only the slots that search this tick (their number is known on the host: act[])
        const bool compact = mc->evaluator == EVAL_NN;
        if (!opp) {
            if (compact && act[0] < (unsigned long long)a.n) {
                k_arena_list<<<1, 1024, 0, s>>>(a.active, a.n, 1, ar->d_list[0]);
                mc->eval_games = ar->d_list[0]; mc->eval_count = (int)act[0];
            }
            rc = search_once(mc, 0, 0, 1, 0, s);                                // one AlphaZero move on every slot that is waiting for one
            mc->eval_games = nullptr; mc->eval_count = 0;
            if (rc) return rc;
        } else {
            // one move of player 0's searcher on its slots, then one of player 1's on the others; a slot whose move passed the turn
            // over waits for the next tick (k_arena_advance sets the turn-start trim first)
            mc->d.side_sel = a.to_move; mc->d.side = 0; opp->d.side_sel = a.to_move; opp->d.side = 1;
            az_mcts* side_mc[2] = { mc, opp };
            for (int side = 0; side < 2 && !rc; ++side) {
                if (!act[1 + side]) continue;
                az_mcts* m2 = side_mc[side];
                if (m2->evaluator == EVAL_NN && act[1 + side] < (unsigned long long)a.n) {
                    k_arena_list<<<1, 1024, 0, s>>>(a.to_move, a.n, side, ar->d_list[side]);
                    m2->eval_games = ar->d_list[side]; m2->eval_count = (int)act[1 + side];
                }
                rc = search_once(m2, 0, 0, 1, 0, s);
                m2->eval_games = nullptr; m2->eval_count = 0;
            }
            mc->d.side_sel = nullptr; opp->d.side_sel = nullptr;
            if (rc) return rc;
        }
        ++ticks;
    }
    unsigned long long r[ARENA_N], c[CNT_N], co[CNT_N];
    memset(co, 0, sizeof co);
    AZ_CUDA(cudaMemcpyAsync(r, a.res, sizeof r, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(c, mc->d.counters, sizeof c, cudaMemcpyDeviceToHost, s));
    if (opp) AZ_CUDA(cudaMemcpyAsync(co, opp->d.counters, sizeof co, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    if (opp) r[ARENA_OPP_TURNS] = co[CNT_STEPS];
    h_out->count = r[ARENA_COUNT]; h_out->draw = r[ARENA_DRAW];
    h_out->win[0] = r[ARENA_WIN0]; h_out->win[1] = r[ARENA_WIN1];
    h_out->win_and_started[0] = r[ARENA_WAS0]; h_out->win_and_started[1] = r[ARENA_WAS1];
    h_out->opponent_turns = r[ARENA_OPP_TURNS]; h_out->az_moves = c[CNT_STEPS]; h_out->az_sims = c[CNT_SIMS]; h_out->az_evals = c[CNT_EVALS];
    h_out->ticks = ticks; h_out->errors = c[CNT_POOL_OVERFLOW] + c[CNT_DEPTH_OVERFLOW] + c[CNT_ILLEGAL] + co[CNT_POOL_OVERFLOW] + co[CNT_DEPTH_OVERFLOW] + co[CNT_ILLEGAL];
    return AZ_OK;
}
