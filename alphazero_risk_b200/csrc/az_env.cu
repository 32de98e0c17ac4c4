// az_env.cu — lockstep Risk environment kernels for sm_100a and the az_env_* C ABI.
//
// One thread per game.  Game state lives in HBM as a structure of arrays of 32-bit words
// (word w of game g at state[w * n + g]) so every load/store of a warp is one coalesced
// 128-byte line; inside a kernel the 42 land bytes of a thread sit in a shared-memory
// COLUMN (word w of thread t at smem[w * blockDim + t]) which makes the divergent,
// data-dependent byte indexing of the rules bank-conflict free, and the map tables are
// staged in shared memory once per block.
//
// Replaces (reference, /root/reference/src/risk_game): State::newGame state/state.cpp:137-167,
// UtilityNN::getValidMoves / makeMove player/alpha_zero/alphazero_moves.cpp:3-233,
// State::gameStatus state/state.cpp:518-565, NNInputData + setInStateTensor
// neural_network/alphazero_nn_data.cpp:165-196 + alphazero_nn.cpp:31-67.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "az_common.cuh"
#include "az_game.cuh"
#include "az_tables_gen.h"
#include "az_script.cuh"
#include "az_arena.cuh"
#include "az_samples.cuh"

// ---------------------------------------------------------------- error string
static thread_local char g_az_err[512] = "";
void az_set_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_az_err, sizeof g_az_err, fmt, ap);
    va_end(ap);
}
extern "C" const char* az_last_error(void) { return g_az_err; }
extern "C" int az_version(void) { return 100; }
extern "C" int az_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" void az_default_rules(az_rules* r)
{   // /root/reference/src/settings.h:40-62
    r->allow_yield = 1; r->limit_reinforcement = 1; r->limit_attack = 0; r->max_game_rounds = 58; r->min_unit_move = 3;
    r->mcts_simulations = 32; r->threads_per_mcts = 2; r->cpuct = 1.1f; r->dir_noise_value = 0.3f; r->dir_noise_epsi = 0.25f;
    r->temperature_threshold = 43;
    r->concurrent_descents = 1;
}

// ---------------------------------------------------------------- tables in HBM (one copy per device)
static uint64_t* g_tables_dev[64] = { nullptr };
static std::mutex g_tables_mu;

int az_upload_tables()
{
    int dev = 0;
    AZ_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tables_mu);
    if (dev < 0 || dev >= 64) { az_set_error("device index %d out of range", dev); return AZ_ERR_INVALID_ARG; }
    if (g_tables_dev[dev]) return AZ_OK;
    static uint64_t h[AZ_TABLE_U64 + AZ_LUT11_U64];               // guarded by g_tables_mu
    memcpy(h, AZ_NBR_UNION_LUT_H, sizeof AZ_NBR_UNION_LUT_H);
    memcpy(h + 7 * 64, AZ_NBR_MASK_H, sizeof AZ_NBR_MASK_H);
    memcpy(h + 7 * 64 + 42, AZ_NBR_LIST6_H, sizeof AZ_NBR_LIST6_H);
    // the rollout kernel's wide tables (az_nbr_union(AzTablesWide)): entry v of table k = union of the neighbour masks of
    // the lands 11k + b, b = the set bits of v
    for (int k = 0; k < 4; ++k)
        for (int v = 0; v < 2048; ++v) {
            uint64_t u = 0;
            for (int b = 0; b < 11; ++b)
                if (((v >> b) & 1) && 11 * k + b < AZ_LANDS) u |= AZ_NBR_MASK_H[11 * k + b];
            h[AZ_TABLE_U64 + k * 2048 + v] = u;
        }
    uint64_t* d = nullptr;
    AZ_CUDA(cudaMalloc(&d, sizeof h));
    AZ_CUDA(cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice));
    g_tables_dev[dev] = d;
    return AZ_OK;
}
const uint64_t* az_device_tables()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < 64) ? g_tables_dev[dev] : nullptr;
}

// ---------------------------------------------------------------- per-thread game context
#ifndef ENV_BLOCK
// The rollout kernel is bound by the latency of each warp's dependent instructions, so what matters is how many warps share an SM's
// schedulers.  65536 games are 13.8 warps per SM: 448-thread blocks make 147 blocks = one per SM = 14 warps on every SM, where
// 256-thread blocks put two blocks (16 warps) on 108 SMs and one on 40.  Swept on B200 with the lane-per-game rollout (G steps/s):
// 32..128: 13.7-13.8 (before the wide tables), 224: 15.3, 256: 15.1, 320: 13.0, 448: 15.85; the pooled rollout keeps 448 games per
// block and gains nothing from more warps than that (512 / 576 / 640 threads: +0.5 / -0.5 / -1.3 %).  Static shared memory of the
// other kernels stays under 48 KB.
#define ENV_BLOCK 448
#endif
#define ENV_COL_WORDS 22     // 11 land words + 11 fortify-DFS parent words per thread

struct EnvSmem {
    uint64_t tab[AZ_TABLE_U64];
    uint32_t col[ENV_COL_WORDS * ENV_BLOCK];
};

__device__ __forceinline__ AzTables env_stage_tables(EnvSmem& sm, const uint64_t* __restrict__ g_tab)
{
    for (int i = threadIdx.x; i < AZ_TABLE_U64; i += blockDim.x) sm.tab[i] = g_tab[i];
    __syncthreads();
    return az_tables_from_smem(sm.tab);
}

// k_env_rollout's (dynamic) shared memory: the same two members first, so the per-thread helpers below work on it unchanged
struct EnvSmemWide {
    uint64_t tab[AZ_TABLE_U64];
    uint32_t col[ENV_COL_WORDS * ENV_BLOCK];
    uint64_t lut11[AZ_LUT11_U64];
};

__device__ __forceinline__ AzTablesWide env_stage_tables_wide(EnvSmemWide& sm, const uint64_t* __restrict__ g_tab)
{
    for (int i = threadIdx.x; i < AZ_TABLE_U64; i += blockDim.x) sm.tab[i] = g_tab[i];
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(g_tab + AZ_TABLE_U64);     // AZ_TABLE_U64 is even: 16-byte aligned
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(sm.lut11);
    for (int i = threadIdx.x; i < AZ_LUT11_U64 / 2; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    AzTablesWide t;
    t.lut = sm.tab; t.nbr = sm.tab + 7 * 64; t.list6 = sm.tab + 7 * 64 + 42; t.lut11 = sm.lut11;
    return t;
}

struct EnvCtx {
    AzGame g;
    AzLandColumn land, scratch;
    uint32_t ply;
};

__device__ __forceinline__ void env_bind(EnvCtx& c, EnvSmem& sm)
{
    c.land.base = (uint8_t*)(sm.col) + 4 * threadIdx.x;
    c.land.stride_bytes = 4 * ENV_BLOCK;
    c.scratch.base = (uint8_t*)(sm.col + 11 * ENV_BLOCK) + 4 * threadIdx.x;
    c.scratch.stride_bytes = 4 * ENV_BLOCK;
}

__device__ __forceinline__ void env_load(EnvCtx& c, EnvSmem& sm, const uint32_t* __restrict__ st, int n, int gi)
{
    env_bind(c, sm);
    c.g.own0 = c.g.own1 = c.g.gt1 = c.g.full = 0;
    uint32_t w10 = 0;
#pragma unroll
    for (int w = 0; w < 11; ++w) {
        uint32_t v = st[(size_t)w * n + gi];
        sm.col[w * ENV_BLOCK + threadIdx.x] = v;
        az_masks_add_word(c.g, v, w);
        if (w == 10) w10 = v;
    }
    az_unpack_scalars(c.g, w10, st[(size_t)11 * n + gi], st[(size_t)12 * n + gi], st[(size_t)13 * n + gi]);
    c.ply = st[(size_t)AZ_W_PLY * n + gi];
}

__device__ __forceinline__ void env_store(const EnvCtx& c, EnvSmem& sm, uint32_t* __restrict__ st, int n, int gi)
{
#pragma unroll
    for (int w = 0; w < 10; ++w) st[(size_t)w * n + gi] = sm.col[w * ENV_BLOCK + threadIdx.x];
    uint32_t w10 = (sm.col[10 * ENV_BLOCK + threadIdx.x] & 0xffffu) | (c.g.cards0 << 16) | (c.g.cards1 << 24);
    st[(size_t)10 * n + gi] = w10;
    st[(size_t)11 * n + gi] = az_pack_w11(c.g);
    st[(size_t)12 * n + gi] = az_pack_w12(c.g);
    st[(size_t)13 * n + gi] = az_pack_w13(c.g);
    st[(size_t)AZ_W_PLY * n + gi] = c.ply;
}

// ---------------------------------------------------------------- kernels
__global__ void __launch_bounds__(ENV_BLOCK) k_env_reset(uint32_t* __restrict__ st, int n, uint64_t seed, uint32_t first_game)
{
    __shared__ EnvSmem sm;
    int gi = blockIdx.x * ENV_BLOCK + threadIdx.x;
    if (gi >= n) return;
    EnvCtx c; env_bind(c, sm);
#pragma unroll
    for (int w = 0; w < 11; ++w) sm.col[w * ENV_BLOCK + threadIdx.x] = 0;
    az_new_game(c.g, c.land, seed, first_game + (uint32_t)gi, 0u);
    c.ply = 0;
    env_store(c, sm, st, n, gi);
    st[(size_t)15 * n + gi] = 0;
}

// one makeMove per game with externally supplied actions (the vector-env API)
__global__ void __launch_bounds__(ENV_BLOCK) k_env_step(uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab,
                                                         const uint8_t* __restrict__ action, const uint8_t* __restrict__ dice5,
                                                         int8_t* __restrict__ status, uint64_t* __restrict__ valid_after,
                                                         uint64_t seed, uint32_t first_game, AzRulesDev rules)
{
    __shared__ EnvSmem sm;
    AzTables T = env_stage_tables(sm, g_tab);
    int gi = blockIdx.x * ENV_BLOCK + threadIdx.x;
    if (gi >= n) return;
    EnvCtx c; env_load(c, sm, st, n, gi);
    int stt = az_game_status(c.g, rules);
    int out;
    if (stt != AZ_STATUS_RUNNING) out = AZ_STATUS_OVER;
    else {
        uint64_t valid = az_valid_moves(c.g, T, rules);
        int rc;
        if (dice5) {
            AzDiceTape d; d.t = dice5 + (size_t)gi * 5; d.j = 0;
            rc = az_make_move(c.g, c.land, c.scratch, T, rules, valid, (int)action[gi], d);
        } else {
            AzDicePhilox d; d.init(seed, first_game + (uint32_t)gi, c.ply, AZ_STREAM_REAL);
            rc = az_make_move(c.g, c.land, c.scratch, T, rules, valid, (int)action[gi], d);
        }
        if (rc == 0) { c.ply++; env_store(c, sm, st, n, gi); out = az_game_status(c.g, rules); }
        else out = rc;
    }
    status[gi] = (int8_t)out;
    if (valid_after) valid_after[gi] = az_valid_moves(c.g, T, rules);
}

__global__ void __launch_bounds__(ENV_BLOCK) k_env_query(const uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab,
                                                          uint64_t* __restrict__ valid, int8_t* __restrict__ status, AzRulesDev rules)
{
    __shared__ EnvSmem sm;
    AzTables T = env_stage_tables(sm, g_tab);
    int gi = blockIdx.x * ENV_BLOCK + threadIdx.x;
    if (gi >= n) return;
    EnvCtx c; env_load(c, sm, st, n, gi);
    if (valid) valid[gi] = az_valid_moves(c.g, T, rules);
    if (status) status[gi] = (int8_t)az_game_status(c.g, rules);
}

// ---------------------------------------------------------------- the rollout as a pool of games per block
// BASELINE config 2: n_steps uniform-random legal moves per game inside one launch; the state stays in shared memory between
// moves, finished games are re-dealt in place.  The common work of a move (legal mask with ONE neighbour union, Philox block,
// action pick, land writes, attack-army check) is one instruction stream for all phases (az_valid_moves_flat / az_move_flat).
// ncu on the lane-per-game kernel (round 1 / first half of round 2: every lane owns one game for the whole launch, lanes that need
// a re-deal or the fortify component search park until a few of them have accumulated in their warp): 12.8 of 32 lanes active per
// issued instruction — 42 % of the issue slots went to the parked-lane services running with <= 7 lanes and to phase-specific code.
// Here a block's games are not tied to a lane.  Their contexts (land bytes, masks, scalars, ply, moves done: 28 words) live in shared
// memory, there is one queue of game slots per CATEGORY — ATTACK, the other five round phases, "needs a re-deal", "needs the
// component search" — and every warp repeatedly takes up to 32 games of ONE category, loads their contexts into its lanes (land
// bytes into the lane's private column, so the rules code is the same thread code as everywhere else and stays free of bank
// conflicts), makes one move per game, stores the contexts and appends each game to the queue of its new category.  Re-deals and
// component searches of the whole block are thus served in batches of 16+ while nobody waits for them.  No block-wide barrier
// (round 1's sorted kernel lost to its three __syncthreads per move): warps meet only in the queues' shared-memory atomics.  A
// game finishes its n_steps whenever it gets there; games are independent and their random streams are keyed by (game, ply), so
// the order does not change any result.  Measured (profiles/README.md): 26 games per warp-step, 20 active lanes per issued
// instruction, +22 % steps/s over the lane-per-game kernel; grouping all six phases in one queue or one queue per phase both lose.
#define ENV_POOL_SLOTS ENV_BLOCK                 // games per block (the grid is env_grid(n), as for every other env kernel)
#define ENV_POOL_CTX_Q 7                         // uint4 per game context (28 words, 25 used)
#define ENV_POOL_NCAT 4
#define ENV_CAT_OTHER 0
#define ENV_CAT_ATTACK 1
#define ENV_CAT_REDEAL 2
#define ENV_CAT_FORTIFY 3
#define ENV_CAT_NONE 4
#define ENV_POOL_NSUB 8                          // rings per category: slot % 8 (see the comment at the struct)
#define ENV_POOL_QSUB 64                         // ring cells: more than the ENV_POOL_SLOTS / ENV_POOL_NSUB games that can ever be in one
static_assert(ENV_POOL_SLOTS / ENV_POOL_NSUB < ENV_POOL_QSUB && ENV_POOL_SLOTS % 32 == 0 && ENV_POOL_NCAT * ENV_POOL_NSUB == 32, "ring sizing");
#define ENV_POOL_MIN_BATCH 16                    // re-deals / component searches are served once this many wait (swept 8..32 on B200)

// A category's queue is EIGHT rings, one per slot % 8, and lane l of a taking warp is served from ring l % 8.  The eight lanes of a
// quarter-warp then hold slots with eight different residues, so their 16-byte context accesses (112-byte stride: residues 0..7
// map to the eight 16-byte bank groups) are free of bank conflicts — with one ring per category the random slots of a batch cost
// ~2.2 wavefronts per ideal one and the shared-memory pipe sat at 60 % (ncu), which every other latency of the kernel then paid for.
// It also spreads the queue atomics over eight words per category and lets every lane append its own game without any warp-level
// grouping.
struct EnvSmemPool {                             // starts like EnvSmemWide / EnvSmem: the per-thread helpers above work on it unchanged
    uint64_t tab[AZ_TABLE_U64];
    uint32_t col[ENV_COL_WORDS * ENV_BLOCK];     // per LANE: the land words + DFS parents of the game the lane works on right now
    uint64_t lut11[AZ_LUT11_U64];
    uint4 ctx[ENV_POOL_SLOTS * ENV_POOL_CTX_Q];  // per GAME
    // ring cell = (sequence << 16) | slot.  Position p (a 32-bit counter, cell p % QSUB) may be written when the cell's sequence is
    // p (mod 2^16), holds its game when it is p + 1, and is handed to position p + QSUB by its taker: a bounded multi-producer /
    // multi-consumer queue that stays correct however long a warp stalls between reserving a position and using it
    uint32_t q[ENV_POOL_NCAT][ENV_POOL_NSUB][ENV_POOL_QSUB];
    uint32_t head[ENV_POOL_NCAT][ENV_POOL_NSUB], tail[ENV_POOL_NCAT][ENV_POOL_NSUB];
    int cnt[ENV_POOL_NCAT][ENV_POOL_NSUB];       // appended and not yet taken (briefly negative while a taker over-reserves)
    uint32_t remaining;                          // games that still have moves to make
};

__device__ __forceinline__ void pool_load(EnvCtx& c, EnvSmem& sm, const uint4* __restrict__ ctx, int slot, int& done, int& park_li)
{
    const uint4* p = ctx + slot * ENV_POOL_CTX_Q;
    const uint4 a = p[0], b = p[1], d = p[2], e = p[3], f = p[4], h = p[5], k = p[6];
    uint32_t* col = sm.col + threadIdx.x;
    col[0 * ENV_BLOCK] = a.x; col[1 * ENV_BLOCK] = a.y; col[2 * ENV_BLOCK] = a.z; col[3 * ENV_BLOCK] = a.w;
    col[4 * ENV_BLOCK] = b.x; col[5 * ENV_BLOCK] = b.y; col[6 * ENV_BLOCK] = b.z; col[7 * ENV_BLOCK] = b.w;
    col[8 * ENV_BLOCK] = d.x; col[9 * ENV_BLOCK] = d.y; col[10 * ENV_BLOCK] = d.z;
    az_unpack_scalars(c.g, d.z, d.w, e.x, e.y);
    c.ply = e.z; done = (int)e.w;
    c.g.own0 = (uint64_t)f.x | ((uint64_t)f.y << 32); c.g.own1 = (uint64_t)f.z | ((uint64_t)f.w << 32);
    c.g.gt1 = (uint64_t)h.x | ((uint64_t)h.y << 32); c.g.full = (uint64_t)h.z | ((uint64_t)h.w << 32);
    park_li = (int)k.x;
}

__device__ __forceinline__ void pool_store(const EnvCtx& c, const EnvSmem& sm, uint4* __restrict__ ctx, int slot, int done, int park_li)
{
    const uint32_t* col = sm.col + threadIdx.x;
    uint4* p = ctx + slot * ENV_POOL_CTX_Q;
    p[0] = make_uint4(col[0 * ENV_BLOCK], col[1 * ENV_BLOCK], col[2 * ENV_BLOCK], col[3 * ENV_BLOCK]);
    p[1] = make_uint4(col[4 * ENV_BLOCK], col[5 * ENV_BLOCK], col[6 * ENV_BLOCK], col[7 * ENV_BLOCK]);
    const uint32_t w10 = (col[10 * ENV_BLOCK] & 0xffffu) | (c.g.cards0 << 16) | (c.g.cards1 << 24);
    p[2] = make_uint4(col[8 * ENV_BLOCK], col[9 * ENV_BLOCK], w10, az_pack_w11(c.g));
    p[3] = make_uint4(az_pack_w12(c.g), az_pack_w13(c.g), c.ply, (uint32_t)done);
    p[4] = make_uint4((uint32_t)c.g.own0, (uint32_t)(c.g.own0 >> 32), (uint32_t)c.g.own1, (uint32_t)(c.g.own1 >> 32));
    p[5] = make_uint4((uint32_t)c.g.gt1, (uint32_t)(c.g.gt1 >> 32), (uint32_t)c.g.full, (uint32_t)(c.g.full >> 32));
    p[6] = make_uint4((uint32_t)park_li, 0u, 0u, 0u);
}

// where a game goes next: nowhere once its n_steps are done, to the re-deal queue when it has ended (its result is counted
// here, once), else to the queue of its round phase
__device__ __forceinline__ int pool_phase_cat(uint32_t ph) { return ph == AZ_PH_ATTACK ? ENV_CAT_ATTACK : ENV_CAT_OTHER; }

// where a game goes next: nowhere once its n_steps are done, to the re-deal queue when it has ended (its result is counted
// here, once), else to the queue of its round phase
__device__ __forceinline__ int pool_classify(const AzGame& g, int done, int n_steps, const AzRulesDev& rules,
                                             unsigned& games, unsigned& w0, unsigned& w1, unsigned& dr)
{
    if (done >= n_steps) return ENV_CAT_NONE;
    const int stt = az_game_status(g, rules);
    if (stt != AZ_STATUS_RUNNING) { games++; w0 += stt == 0; w1 += stt == 1; dr += stt == AZ_STATUS_DRAW; return ENV_CAT_REDEAL; }
    return pool_phase_cat(g.phase);
}

// appends the lane's game to the queue of its new category (newcat < 0 or ENV_CAT_NONE: nothing to append)
__device__ __forceinline__ void pool_push(EnvSmemPool& sp, int newcat, int slot)
{
    if (newcat >= 0 && newcat < ENV_POOL_NCAT) {
        const int r = slot & (ENV_POOL_NSUB - 1);
        const uint32_t pos = atomicAdd(&sp.tail[newcat][r], 1u);
        volatile uint32_t* e = &sp.q[newcat][r][pos & (ENV_POOL_QSUB - 1)];
        while ((*e >> 16) != (pos & 0xffffu)) { }                        // the cell's previous lap has been taken (it has, unless a warp stalled for a whole lap)
        *e = (((pos + 1u) & 0xffffu) << 16) | (uint32_t)slot;
        atomicAdd(&sp.cnt[newcat][r], 1);                               // a taker that gets ahead of the cell's store spins on the cell
    }
}

// which category the warp serves next.  Lane l looks at ring l % 8 of category l / 8 (a warp can take up to 4 games from each ring);
// a re-deal / component-search batch once ENV_POOL_MIN_BATCH of them wait, else the fuller of the two move queues (ties alternate),
// else whatever is left.  Returns (priority << 2) | rotated category, 0 = every queue is empty.
__device__ __forceinline__ uint32_t pool_choose(const EnvSmemPool& sp, int lane, unsigned rot)
{
    const int grp = lane >> 3, sub = lane & 7;
    int a = *reinterpret_cast<const volatile int*>(&sp.cnt[grp][sub]);
    a = a < 0 ? 0 : (a > 4 ? 4 : a);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    const uint32_t pr = grp >= ENV_CAT_REDEAL ? (a >= ENV_POOL_MIN_BATCH ? 64u + (uint32_t)a : (a ? 1u : 0u)) : (a ? 8u + (uint32_t)a : 0u);
    const uint32_t score = sub == 0 ? (pr << 2) | (((uint32_t)grp + rot) & 3u) : 0u;
    return __reduce_max_sync(0xffffffffu, score);
}

__global__ void __launch_bounds__(ENV_BLOCK, 1) k_env_rollout(uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab,
                                                               int n_steps, uint64_t seed, uint32_t first_game, AzRulesDev rules,
                                                               unsigned long long* __restrict__ counters)
{
    extern __shared__ __align__(16) unsigned char env_dyn_smem[];
    EnvSmemPool& sp = *reinterpret_cast<EnvSmemPool*>(env_dyn_smem);
    EnvSmemWide& smw = *reinterpret_cast<EnvSmemWide*>(env_dyn_smem);
    EnvSmem& sm = *reinterpret_cast<EnvSmem*>(env_dyn_smem);
    for (int i = threadIdx.x; i < ENV_POOL_NCAT * ENV_POOL_NSUB * ENV_POOL_QSUB; i += ENV_BLOCK) (&sp.q[0][0][0])[i] = (uint32_t)(i & (ENV_POOL_QSUB - 1)) << 16;
    if (threadIdx.x < ENV_POOL_NCAT * ENV_POOL_NSUB) { (&sp.head[0][0])[threadIdx.x] = 0; (&sp.tail[0][0])[threadIdx.x] = 0; (&sp.cnt[0][0])[threadIdx.x] = 0; }
    if (threadIdx.x == 0) sp.remaining = 0;
    const AzTablesWide T = env_stage_tables_wide(smw, g_tab);              // __syncthreads inside
    const int lane = threadIdx.x & 31;
    const int gi0 = blockIdx.x * ENV_POOL_SLOTS, gi = gi0 + (int)threadIdx.x;
    const bool live = gi < n;
    unsigned games = 0, w0 = 0, w1 = 0, dr = 0;
    EnvCtx c;
    {
        // every thread brings its own game in and queues it
        int cat = -1;
        if (live) {
            env_load(c, sm, st, n, gi);
            cat = pool_classify(c.g, 0, n_steps, rules, games, w0, w1, dr);
            pool_store(c, sm, sp.ctx, (int)threadIdx.x, 0, 0);
        } else env_bind(c, sm);
        const int nq = __popc(__ballot_sync(0xffffffffu, cat >= 0 && cat < ENV_POOL_NCAT));
        if (lane == 0 && nq) atomicAdd(&sp.remaining, (uint32_t)nq);
        __threadfence_block();
        pool_push(sp, cat, (int)threadIdx.x);
    }
    __syncthreads();
    unsigned rot = threadIdx.x >> 5;
    const int sub = lane & (ENV_POOL_NSUB - 1), grp = lane >> 3;           // ring this lane is served from; its place in the ring's batch of 4
    uint32_t best = pool_choose(sp, lane, rot);
    for (;;) {
        if ((best >> 2) == 0) {
            if (*reinterpret_cast<volatile uint32_t*>(&sp.remaining) == 0) break;
            __nanosleep(64);
            best = pool_choose(sp, lane, rot);
            continue;
        }
        const int cat = (int)(((best & 3u) - rot) & 3u);
        // lanes 0..7 each take up to 4 games from one ring: reserve 4 at once and give back what was not there
        int kq = 0; uint32_t hq = 0;
        if (lane < ENV_POOL_NSUB) {
            const int old = atomicSub(&sp.cnt[cat][lane], 4);
            kq = old >= 4 ? 4 : (old > 0 ? old : 0);
            if (kq < 4) atomicAdd(&sp.cnt[cat][lane], 4 - kq);
            if (kq) hq = atomicAdd(&sp.head[cat][lane], (uint32_t)kq);
        }
        kq = __shfl_sync(0xffffffffu, kq, sub); hq = __shfl_sync(0xffffffffu, hq, sub);
        const bool has = grp < kq;
        if (!__any_sync(0xffffffffu, has)) { best = pool_choose(sp, lane, rot); continue; }
        ++rot;
        int slot = -1;
        if (has) {
            const uint32_t pos = hq + (uint32_t)grp;
            volatile uint32_t* e = &sp.q[cat][sub][pos & (ENV_POOL_QSUB - 1)];
            uint32_t v;
            while (((v = *e) >> 16) != ((pos + 1u) & 0xffffu)) { }
            *e = ((pos + ENV_POOL_QSUB) & 0xffffu) << 16;
            slot = (int)(v & 0xffffu);
        }
        __threadfence_block();
        int newcat = -1, done = 0, park_li = 0;
        if (slot >= 0) {
            pool_load(c, sm, sp.ctx, slot, done, park_li);
            const uint32_t game = first_game + (uint32_t)(gi0 + slot);
            if (cat == ENV_CAT_REDEAL) {
                az_new_game(c.g, c.land, seed, game, c.ply);
                newcat = pool_phase_cat(c.g.phase);
            } else if (cat == ENV_CAT_FORTIFY) {
                az_fortify_finish(c.g, c.land, c.scratch, T, park_li);
                c.ply++; done++;
                newcat = pool_classify(c.g, done, n_steps, rules, games, w0, w1, dr);
            } else if (cat == ENV_CAT_ATTACK) {
                // the batch's phase is known here: the compiler drops the other phases' selects and branches from this copy
                __builtin_assume(c.g.phase == AZ_PH_ATTACK);
                const uint64_t valid = az_valid_moves_flat(c.g, T, rules);
                const az_u32x4 blk = az_rng_block(seed, game, c.ply, AZ_STREAM_REAL, 0);
                const int action = az_nth_set_bit(valid, az_mulhi32(blk.y, (uint32_t)__popcll(valid)));
                az_move_flat(c.g, c.land, T, rules, action, blk.x);
                c.ply++; done++;
                newcat = pool_classify(c.g, done, n_steps, rules, games, w0, w1, dr);
            } else {
                __builtin_assume(c.g.phase != AZ_PH_ATTACK);
                const uint64_t valid = az_valid_moves_flat(c.g, T, rules);
                const az_u32x4 blk = az_rng_block(seed, game, c.ply, AZ_STREAM_REAL, 0);
                const int action = az_nth_set_bit(valid, az_mulhi32(blk.y, (uint32_t)__popcll(valid)));
                if (az_move_flat(c.g, c.land, T, rules, action, blk.x)) { newcat = ENV_CAT_FORTIFY; park_li = action; }
                else { c.ply++; done++; newcat = pool_classify(c.g, done, n_steps, rules, games, w0, w1, dr); }
            }
        }
        // the next choice is made BEFORE this batch goes back (it may miss the warp's own games, which is harmless): its
        // load -> shuffle -> reduce chain then overlaps the context stores below instead of standing alone between two steps
        // (also RESERVING the next batch here was measured and loses 8 %: games held idle cost more than the hidden atomics)
        best = pool_choose(sp, lane, rot);
        if (slot >= 0) pool_store(c, sm, sp.ctx, slot, done, park_li);
        __threadfence_block();
        pool_push(sp, newcat, slot);
        const int nfin = __popc(__ballot_sync(0xffffffffu, newcat == ENV_CAT_NONE));
        if (lane == 0 && nfin) atomicSub(&sp.remaining, (uint32_t)nfin);
    }
    __syncthreads();
    unsigned steps = 0;
    if (live) {
        int done, park_li;
        pool_load(c, sm, sp.ctx, (int)threadIdx.x, done, park_li);
        env_store(c, sm, st, n, gi);
        steps = (unsigned)done;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_xor_sync(0xffffffffu, steps, o); games += __shfl_xor_sync(0xffffffffu, games, o);
        w0 += __shfl_xor_sync(0xffffffffu, w0, o); w1 += __shfl_xor_sync(0xffffffffu, w1, o); dr += __shfl_xor_sync(0xffffffffu, dr, o);
    }
    if (lane == 0) {
        atomicAdd(&counters[0], (unsigned long long)steps);
        if (games) atomicAdd(&counters[1], (unsigned long long)games);
        if (w0) atomicAdd(&counters[2], (unsigned long long)w0);
        if (w1) atomicAdd(&counters[3], (unsigned long long)w1);
        if (dr) atomicAdd(&counters[4], (unsigned long long)dr);
    }
}

// staging of the samples scripted / random turns emit (Player::addTrainingSample), per game; st == NULL: not recording
struct TurnRecDev {
    uint32_t* st;      // [n][max][14]  state before the move
    uint8_t* mv;       // [n][max]      the move
    uint32_t* len;     // [n]           staged samples of the running game (> max = overflowed)
    int max;
};

// ScriptPlayer::takeTurn / RandomPlayer::takeTurn for the side to move of every running game (one whole turn per call, one ply);
// kind0 / kind1 = who plays side 0 / side 1
__global__ void __launch_bounds__(ENV_BLOCK) k_env_script_turn(uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab,
                                                                uint32_t* __restrict__ script, int8_t* __restrict__ status,
                                                                uint64_t seed, uint32_t first_game, AzRulesDev rules, int kind0, int kind1,
                                                                TurnRecDev rec)
{
    __shared__ EnvSmem sm;
    AzTables T = env_stage_tables(sm, g_tab);
    int gi = blockIdx.x * ENV_BLOCK + threadIdx.x;
    if (gi >= n) return;
    EnvCtx c; env_load(c, sm, st, n, gi);
    int out = az_game_status(c.g, rules);
    if (out != AZ_STATUS_RUNNING) out = AZ_STATUS_OVER;
    else {
        const uint32_t side = c.g.cur;
        const int kind = side ? kind1 : kind0;
        AzTurnSink sink; AzTurnSink* sk = nullptr;
        if (rec.st) {
            sink.st = rec.st + (size_t)gi * rec.max * AZ_PRIMARY_WORDS; sink.mv = rec.mv + (size_t)gi * rec.max;
            sink.cap = (uint32_t)rec.max; sink.n = rec.len[gi];
            sk = &sink;
        }
        uint32_t spw = kind == AZ_OPPONENT_SCRIPT ? script[(size_t)gi * 2 + side] : 0u;
        const int rc = kind == AZ_OPPONENT_SCRIPT ? az_script_turn(c.g, c.land, c.scratch, T, rules, spw, seed, first_game + (uint32_t)gi, c.ply, sk)
                                                  : az_random_turn(c.g, c.land, c.scratch, T, rules, seed, first_game + (uint32_t)gi, c.ply, sk);
        if (rc == 0) {
            if (kind == AZ_OPPONENT_SCRIPT) script[(size_t)gi * 2 + side] = spw;
            if (sk) rec.len[gi] = sink.n;
            c.ply++; env_store(c, sm, st, n, gi); out = az_game_status(c.g, rules);
        }
        else out = rc;
    }
    status[gi] = (int8_t)out;
}

// gameFinished -> NNTrainDataStorage::updateValues (script_player.cpp:229-235, random_player.cpp:113-119, alphazero_nn_data.cpp:51-65)
// for the games the turn kernel just ended: one warp per game turns the staged (state, move) pairs into packed 265-byte records
// (one-hot policy, Player::addTrainingSample) in the output queue.
#define REC_END_WARPS 4
__global__ void __launch_bounds__(REC_END_WARPS * 32) k_env_rec_end(TurnRecDev rec, const int8_t* __restrict__ status, int n,
                                                                    uint8_t* __restrict__ out, unsigned long long* __restrict__ count,
                                                                    unsigned long long cap)
{
    __shared__ uint8_t s_rec_all[REC_END_WARPS][272];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * REC_END_WARPS + warp;
    if (gi >= n) return;
    const int stt = status[gi];
    if (stt != 0 && stt != 1 && stt != AZ_STATUS_DRAW) return;
    const uint32_t len = rec.len[gi];
    if (len == 0) return;
    uint8_t* s_rec = s_rec_all[warp];
    unsigned long long base = 0;
    bool ok = len <= (uint32_t)rec.max;
    if (lane == 0) ok = az_rec_reserve(count, cap, len, ok, &base);
    ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
    base = ((unsigned long long)__shfl_sync(0xffffffffu, (uint32_t)(base >> 32), 0) << 32) | __shfl_sync(0xffffffffu, (uint32_t)base, 0);
    if (ok) {
        for (uint32_t i = 0; i < len; ++i) {
            const uint32_t* stp = rec.st + ((size_t)gi * rec.max + i) * AZ_PRIMARY_WORDS;
            const int mv = rec.mv[(size_t)gi * rec.max + i];
            __syncwarp();
            az_sample_head(stp, stt, s_rec, lane);
            for (int k = lane; k < AZ_MOVES; k += 32) {
                const uint32_t u = k == mv ? 0x3f800000u : 0u;               // policy[li2i(move)] = 1.0f
                for (int b = 0; b < 4; ++b) s_rec[93 + 4 * k + b] = (uint8_t)(u >> (8 * b));
            }
            __syncwarp();
            uint8_t* dst = out + (base + i) * AZ_SAMPLE_BYTES;
            for (int k = lane; k < AZ_SAMPLE_BYTES; k += 32) dst[k] = s_rec[k];
        }
    }
    __syncwarp();
    if (lane == 0) rec.len[gi] = 0;
}

// ---------------------------------------------------------------- arena (-m play): everything between two AlphaZero moves
// GameGroup::threadPlayGame / Game::playGames / Game::newGame / GameResults::addGame (game/game.cpp:153-254) for one game slot per
// thread: tally a finished game, start the next one (fresh deal, or the mirror game = previous start state with the sides swapped
// and player 1 to move, game.cpp:170-179), let the scripted opponent (player index 1) play its turns, and stop as soon as
// player 0 (the AlphaZero side, searched by the MCTS kernels) is to move.  Pairs of games are claimed from a shared counter like
// Counter::hasNext(2) (game.cpp:12-24).
__global__ void __launch_bounds__(ENV_BLOCK) k_arena_advance(ArenaDev a, const uint64_t* __restrict__ g_tab, AzRulesDev rules)
{
    __shared__ EnvSmem sm;
    AzTables T = env_stage_tables(sm, g_tab);
    const int gi = blockIdx.x * ENV_BLOCK + threadIdx.x;
    if (gi >= a.n) return;
    if (!a.active[gi]) return;
    EnvCtx c; env_load(c, sm, a.state, a.n, gi);
    const uint32_t game = a.first_game + (uint32_t)gi;
    uint32_t player_start = a.player_start[gi];
    bool fresh = a.fresh[gi] != 0, active = true;
    uint32_t last = a.last_mover[gi];
    unsigned trim = 0, trim_opp = 0;
    for (int it = 0; it < 64; ++it) {
        const int st = az_game_status(c.g, rules);
        if (fresh || st != AZ_STATUS_RUNNING) {
            if (!fresh) {                                             // GameResults::addGame, game.cpp:193-213
                atomicAdd(&a.res[ARENA_COUNT], 1ull);
                a.ended[gi] = (uint8_t)(st == AZ_STATUS_DRAW ? 3 : 1 + st);       // Player::gameFinished -> trainStorage->updateValues
                if (st == AZ_STATUS_DRAW) atomicAdd(&a.res[ARENA_DRAW], 1ull);
                else {
                    atomicAdd(&a.res[ARENA_WIN0 + st], 1ull);
                    if ((uint32_t)st == player_start) atomicAdd(&a.res[ARENA_WAS0 + st], 1ull);
                }
                player_start ^= 1u;                                   // Game::incPlayerStart
            }
            if (!fresh && player_start == 1u) {
                // second game of the claimed pair
                if (a.mirror) {                                       // state = previousStartState; invertPlayers(); setCurrentPlayerTurn(1)
#pragma unroll
                    for (int w = 0; w < 11; ++w) {
                        uint32_t v = a.start_state[(size_t)w * a.n + gi];
                        uint32_t o = 0;                               // swap owners 0 <-> 1 (bit 6 of a byte whose bit 7 is clear)
#pragma unroll
                        for (int b = 0; b < 4; ++b) { uint32_t by = (v >> (8 * b)) & 0xffu; if (4 * w + b < AZ_LANDS && (by & 0x80u) == 0) by ^= 0x40u; o |= by << (8 * b); }
                        sm.col[w * ENV_BLOCK + threadIdx.x] = o;
                    }
                    const uint32_t w10 = sm.col[10 * ENV_BLOCK + threadIdx.x];
                    // cards travel with PlayerStatus (both 0 at a game start); scalars as dealt
                    c.g.own0 = c.g.own1 = c.g.gt1 = c.g.full = 0;
#pragma unroll
                    for (int w = 0; w < 11; ++w) az_masks_add_word(c.g, sm.col[w * ENV_BLOCK + threadIdx.x], w);
                    az_unpack_scalars(c.g, (w10 & 0xffffu) | ((w10 >> 24) << 16) | (((w10 >> 16) & 0xffu) << 24),
                                      a.start_state[(size_t)11 * a.n + gi], a.start_state[(size_t)12 * a.n + gi], a.start_state[(size_t)13 * a.n + gi]);
                } else az_new_game(c.g, c.land, a.seed, game, c.ply);
                c.g.cur = 1u;
            } else {
                // claim the next pair (Counter::hasNext(2)); a slot that gets none is done
                const unsigned long long before = atomicAdd(&a.res[ARENA_CLAIMED], 2ull);
                if (before + 2ull > a.total_games) { atomicAdd(&a.res[ARENA_CLAIMED], (unsigned long long)-2ll); active = false; break; }
                player_start = 0u;
                az_new_game(c.g, c.land, a.seed, game, c.ply);        // State::newGame; setCurrentPlayerTurn(playerStart = 0)
                env_store(c, sm, a.start_state, a.n, gi);             // previousStartState
            }
            fresh = false;
            trim = 2; trim_opp = 2; last = 0xffu;                     // Player::newGame: AlphaZeroPlayer clears its table
            continue;
        }
        if (a.opponent == AZ_OPPONENT_ALPHAZERO) {                    // both sides search: hand the slot to the side to move
            if (c.g.cur == 0u) { if (last != 0u && trim < 2) trim = 1; }          // AlphaZeroPlayer::takeTurn trims when its turn starts
            else if (last != 1u && trim_opp < 2) trim_opp = 1;
            last = c.g.cur;
            break;
        }
        if (c.g.cur == 1u) {                                          // the opponent's whole turn
            uint32_t spw = a.script[(size_t)gi * 2 + 1];
            if (a.opponent == AZ_OPPONENT_SCRIPT) az_script_turn(c.g, c.land, c.scratch, T, rules, spw, a.seed, game, c.ply);
            else az_random_turn(c.g, c.land, c.scratch, T, rules, a.seed, game, c.ply);
            a.script[(size_t)gi * 2 + 1] = spw;
            c.ply++;
            atomicAdd(&a.res[ARENA_OPP_TURNS], 1ull);
            last = 1u;
            continue;
        }
        // player 0 = AlphaZero is to move: AlphaZeroPlayer::takeTurn trims once when its turn starts (alphazero_player.cpp:5)
        if (last != 0u && trim < 2) trim = 1;
        last = 0u;
        break;
    }
    env_store(c, sm, a.state, a.n, gi);
    a.player_start[gi] = (uint8_t)player_start; a.fresh[gi] = 0; a.last_mover[gi] = (uint8_t)last;
    a.active[gi] = active ? 1 : 0;
    if (active) { a.extra_trim[gi] = (uint8_t)(a.extra_trim[gi] + trim); atomicAdd(&a.res[ARENA_ACTIVE], 1ull); }
    if (a.opponent == AZ_OPPONENT_ALPHAZERO) {
        a.to_move[gi] = active ? (uint8_t)c.g.cur : (uint8_t)0xff;
        if (active) { a.extra_trim_opp[gi] = (uint8_t)(a.extra_trim_opp[gi] + trim_opp); atomicAdd(&a.res[ARENA_TOMOVE0 + c.g.cur], 1ull); }
    }
}

// AoS Data image (state/state.h:86-105, g++ x86-64 layout) <-> device SoA
__device__ __forceinline__ void put48(uint8_t* p, uint64_t v) { for (int i = 0; i < 6; ++i) p[i] = (uint8_t)(v >> (8 * i)); }
__device__ __forceinline__ uint64_t get48(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 6; ++i) v |= (uint64_t)p[i] << (8 * i); return v; }

// The 160-byte image of every game of the block is assembled in shared memory (41-word rows: conflict-free) and leaves as ONE
// contiguous run of 4-byte words per block — a thread writing its own image byte by byte issued 278 uncoalesced stores per game
// (150 us per 16384 games, round-2 launch list).
#define ENV_IMG_WORDS (AZ_DATA_BYTES / 4)
#define ENV_IMG_STRIDE (ENV_IMG_WORDS + 1)
__global__ void __launch_bounds__(ENV_BLOCK) k_env_export(const uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab,
                                                           uint8_t* __restrict__ aos)
{
    __shared__ EnvSmem sm;
    extern __shared__ __align__(16) uint32_t env_img[];              // [ENV_BLOCK][ENV_IMG_STRIDE]
    AzTables T = env_stage_tables(sm, g_tab);
    const int g0 = blockIdx.x * ENV_BLOCK, gi = g0 + threadIdx.x;
    if (gi < n) {
        EnvCtx c; env_load(c, sm, st, n, gi);
        uint32_t* dw = env_img + threadIdx.x * ENV_IMG_STRIDE;
        for (int i = 0; i < ENV_IMG_WORDS; ++i) dw[i] = 0;
        uint8_t* d = reinterpret_cast<uint8_t*>(dw);
        int total[2] = { 0, 0 };
        for (int i = 0; i < AZ_LANDS; ++i) {
            uint32_t v = c.land.get(i); d[i] = (uint8_t)v;
            if ((v >> 6) < 2) total[v >> 6] += (int)(v & 63u);
        }
        for (uint32_t p = 0; p < 2; ++p) {
            uint8_t* ps = d + 48 + 48 * p;
            uint64_t o = c.g.own(p);
            put48(ps + 0, o); put48(ps + 8, o & c.g.gt1); put48(ps + 16, o & c.g.full);
            put48(ps + 24, az_nbr_union(T, o) & ~o); put48(ps + 32, az_attack_army(c.g, T, p));
            ps[38] = (uint8_t)(total[p] & 0xff); ps[39] = (uint8_t)((total[p] >> 8) & 0xff);
            ps[40] = (uint8_t)(p ? c.g.cards1 : c.g.cards0);
        }
        d[144] = (uint8_t)(c.g.round & 0xff); d[145] = (uint8_t)(c.g.round >> 8);
        d[146] = (uint8_t)c.g.cur; d[147] = (uint8_t)c.g.card_sets; d[148] = (uint8_t)c.g.reinf; d[149] = (uint8_t)c.g.phase;
        d[150] = (uint8_t)c.g.mob_from; d[151] = (uint8_t)c.g.mob_to; d[152] = (uint8_t)c.g.allow_draw; d[153] = (uint8_t)c.g.attacks;
    }
    __syncthreads();
    const int nb = n - g0 < ENV_BLOCK ? n - g0 : ENV_BLOCK;
    uint32_t* out = reinterpret_cast<uint32_t*>(aos) + (size_t)g0 * ENV_IMG_WORDS;       // cudaMalloc'ed staging buffer: 4-byte aligned
    for (int j = threadIdx.x; j < nb * ENV_IMG_WORDS; j += ENV_BLOCK) {
        const int gl = j / ENV_IMG_WORDS;
        out[j] = env_img[gl * ENV_IMG_STRIDE + (j - gl * ENV_IMG_WORDS)];
    }
}

__global__ void __launch_bounds__(ENV_BLOCK) k_env_import(uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab,
                                                           const uint8_t* __restrict__ aos, int* __restrict__ bad)
{
    __shared__ EnvSmem sm;
    AzTables T = env_stage_tables(sm, g_tab);
    int gi = blockIdx.x * ENV_BLOCK + threadIdx.x;
    if (gi >= n) return;
    const uint8_t* d = aos + (size_t)gi * AZ_DATA_BYTES;
    EnvCtx c; env_bind(c, sm);
    c.g.own0 = c.g.own1 = c.g.gt1 = c.g.full = 0;
    int total[2] = { 0, 0 };
    for (int w = 0; w < 11; ++w) {
        uint32_t v = 0;
        for (int b = 0; b < 4; ++b) { int i = 4 * w + b; if (i < AZ_LANDS) v |= (uint32_t)d[i] << (8 * b); }
        sm.col[w * ENV_BLOCK + threadIdx.x] = v;
        az_masks_add_word(c.g, v, w);
    }
    for (int i = 0; i < AZ_LANDS; ++i) { uint32_t v = d[i]; if ((v >> 6) < 2) total[v >> 6] += (int)(v & 63u); if ((v >> 6) == 3) atomicAdd(bad, 1); }
    c.g.cards0 = d[48 + 40]; c.g.cards1 = d[96 + 40];
    c.g.round = (uint32_t)d[144] | ((uint32_t)d[145] << 8); c.g.cur = d[146]; c.g.card_sets = d[147]; c.g.reinf = d[148]; c.g.phase = d[149];
    c.g.mob_from = d[150]; c.g.mob_to = d[151]; c.g.allow_draw = d[152]; c.g.attacks = d[153];
    // the derived fields of the image must agree with landArmy[] (State::consistencyCheck, state.cpp:1209-1429)
    int wrong = 0;
    for (uint32_t p = 0; p < 2; ++p) {
        const uint8_t* ps = d + 48 + 48 * p;
        uint64_t o = c.g.own(p);
        wrong += get48(ps) != o; wrong += get48(ps + 8) != (o & c.g.gt1); wrong += get48(ps + 16) != (o & c.g.full);
        wrong += get48(ps + 24) != (az_nbr_union(T, o) & ~o); wrong += get48(ps + 32) != az_attack_army(c.g, T, p);
        wrong += (int)(int16_t)((uint16_t)ps[38] | ((uint16_t)ps[39] << 8)) != total[p];
    }
    wrong += c.g.cur > 1; wrong += c.g.phase > AZ_PH_FORTIFY;
    if (wrong) atomicAdd(bad, 1);
    c.ply = 0;
    env_store(c, sm, st, n, gi);
    st[(size_t)15 * n + gi] = 0;
}

// NNInputData(State) + setInStateTensor -> fp32 [n][7][6][13]; one warp per game, coalesced stores
__global__ void __launch_bounds__(128) k_env_encode(const uint32_t* __restrict__ st, int n, float* __restrict__ x)
{
    __shared__ uint32_t s_words[4][16];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int gi = blockIdx.x * 4 + warp;
    if (gi >= n) return;
    if (lane < 14) s_words[warp][lane] = st[(size_t)lane * n + gi];
    __syncwarp();
    const uint8_t* land = (const uint8_t*)s_words[warp];
    AzGame g;
    az_unpack_scalars(g, s_words[warp][10], s_words[warp][11], s_words[warp][12], s_words[warp][13]);
    // lanes 0..31 hold lands 0..31, lanes 0..9 also lands 32..41
    uint32_t b0 = land[lane], b1 = lane < 10 ? land[32 + lane] : (3u << 6);
    uint32_t m0a = __ballot_sync(0xffffffffu, (b0 >> 6) == 0), m0b = __ballot_sync(0xffffffffu, (b1 >> 6) == 0);
    uint32_t m1a = __ballot_sync(0xffffffffu, (b0 >> 6) == 1), m1b = __ballot_sync(0xffffffffu, (b1 >> 6) == 1);
    g.own0 = (uint64_t)m0a | ((uint64_t)m0b << 32); g.own1 = (uint64_t)m1a | ((uint64_t)m1b << 32);
    int t0 = ((b0 >> 6) == 0 ? (int)(b0 & 63u) : 0) + ((b1 >> 6) == 0 ? (int)(b1 & 63u) : 0);
    int t1 = ((b0 >> 6) == 1 ? (int)(b0 & 63u) : 0) + ((b1 >> 6) == 1 ? (int)(b1 & 63u) : 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); }
    const uint32_t cur = g.cur, enemy = cur ^ 1u;
    float ref = (float)az_reinforcement_value(g.own(cur)), eref = (float)az_reinforcement_value(g.own(enemy));
    float reinf_share = __fdiv_rn(ref, __fadd_rn(ref, eref));
    float att = __fdiv_rn((float)g.attacks, 8.0f); att = att < 1.0f ? att : 1.0f;
    float ta = (float)(cur ? t1 : t0), eta = (float)(cur ? t0 : t1);
    float army_share = __fdiv_rn(ta, __fadd_rn(ta, eta));
    float* out = x + (size_t)gi * AZ_INPUT_FLOATS;
    for (int idx = lane; idx < AZ_INPUT_FLOATS; idx += 32) {
        int l = idx / 13, ch = idx - l * 13;
        uint32_t v = land[l], o = v >> 6;
        float fa = __fdiv_rn((float)(v & 63u), 32.0f);
        float val;
        if (ch == 0) val = o == cur ? fa : 0.0f;
        else if (ch == 1) val = o == enemy ? fa : 0.0f;
        else if (ch == 2) val = o == AZ_NEUTRAL ? fa : 0.0f;
        else if (ch == 3) val = army_share;
        else if (ch == 4) val = reinf_share;
        else if (ch == 5) val = att;
        else if (ch == 6) val = g.allow_draw ? 1.0f : 0.0f;
        else val = (g.phase == (uint32_t)(ch - 7)) ? 1.0f : 0.0f;
        out[idx] = val;
    }
}

// ---------------------------------------------------------------- host side
struct az_env {
    int n = 0, device = 0;
    uint32_t first_game = 0;
    uint64_t seed = 0;
    az_rules rules;
    uint32_t* d_state = nullptr;
    unsigned long long* d_counters = nullptr;   // 8 x u64
    // staging for the host-buffer entry points
    uint8_t* d_action = nullptr; uint8_t* d_dice = nullptr; int8_t* d_status = nullptr; uint64_t* d_valid = nullptr;
    uint8_t* d_aos = nullptr; float* d_x = nullptr; int* d_bad = nullptr; uint32_t* d_script = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // samples of scripted / random turns (az_env_record_turns)
    TurnRecDev rec = { nullptr, nullptr, nullptr, 0 };
    uint8_t* d_rec_out = nullptr; unsigned long long* d_rec_count = nullptr; size_t rec_cap = 0;
};

static AzRulesDev dev_rules(const az_rules& r)
{
    AzRulesDev d; d.allow_yield = r.allow_yield; d.limit_reinforcement = r.limit_reinforcement; d.limit_attack = r.limit_attack;
    d.max_game_rounds = r.max_game_rounds; d.min_unit_move = r.min_unit_move; return d;
}
static inline int env_grid(int n) { return (n + ENV_BLOCK - 1) / ENV_BLOCK; }

int az_launch_arena_advance(const ArenaDev& a, const az_rules* rules, cudaStream_t s)
{
    k_arena_advance<<<env_grid(a.n), ENV_BLOCK, 0, s>>>(a, az_device_tables(), dev_rules(*rules));
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}


extern "C" int az_env_destroy(az_env* e);

extern "C" int az_env_create(int n_games, const az_rules* rules, int device, uint32_t first_game_id, az_env** out)
{
    AZ_REQUIRE(out != nullptr, "out is NULL");
    AZ_REQUIRE(n_games > 0, "n_games must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); az_set_error("no CUDA device: libaz_b200 has no CPU fallback"); return AZ_ERR_NO_DEVICE; }
    AZ_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    AzDeviceGuard guard(device);
    int rc = az_upload_tables();
    if (rc != AZ_OK) return rc;
    az_env* e = new (std::nothrow) az_env();
    AZ_REQUIRE(e != nullptr, "out of host memory");
    e->n = n_games; e->device = device; e->first_game = first_game_id;
    if (rules) e->rules = *rules; else az_default_rules(&e->rules);
    // a failed allocation must not leak the handle or what was allocated before it
    const size_t state_bytes = sizeof(uint32_t) * AZ_STATE_WORDS * (size_t)n_games;
    cudaError_t ce = cudaMalloc(&e->d_state, state_bytes);
    if (ce == cudaSuccess) ce = cudaMemset(e->d_state, 0, state_bytes);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_counters, 8 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMemset(e->d_counters, 0, 8 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_bad, sizeof(int));
    if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev1);
    if (ce != cudaSuccess) {
        az_set_error("az_env_create(%d games): %s", n_games, cudaGetErrorString(ce));
        cudaGetLastError();
        az_env_destroy(e);
        return AZ_ERR_CUDA;
    }
    *out = e;
    return AZ_OK;
}

extern "C" int az_env_destroy(az_env* e)
{
    if (!e) return AZ_OK;
    AzDeviceGuard guard(e->device);
    cudaFree(e->d_state); cudaFree(e->d_counters); cudaFree(e->d_action); cudaFree(e->d_dice); cudaFree(e->d_status);
    cudaFree(e->d_valid); cudaFree(e->d_aos); cudaFree(e->d_x); cudaFree(e->d_bad); cudaFree(e->d_script);
    cudaFree(e->rec.st); cudaFree(e->rec.mv); cudaFree(e->rec.len); cudaFree(e->d_rec_out); cudaFree(e->d_rec_count);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    delete e;
    return AZ_OK;
}

extern "C" int az_env_size(const az_env* e) { return e ? e->n : 0; }

template <class T> static int ensure(T** p, size_t count)
{
    if (*p) return AZ_OK;
    AZ_CUDA(cudaMalloc(p, sizeof(T) * count));
    return AZ_OK;
}

extern "C" int az_env_reset(az_env* e, uint64_t seed, void* stream)
{
    AZ_REQUIRE(e != nullptr, "env is NULL");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    e->seed = seed;
    if (e->rec.len) AZ_CUDA(cudaMemsetAsync(e->rec.len, 0, sizeof(uint32_t) * (size_t)e->n, s));     // nothing staged for a fresh deal
    k_env_reset<<<env_grid(e->n), ENV_BLOCK, 0, s>>>(e->d_state, e->n, seed, e->first_game);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_env_import_aos(az_env* e, const uint8_t* h_data, void* stream)
{
    AZ_REQUIRE(e && h_data, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure(&e->d_aos, (size_t)e->n * AZ_DATA_BYTES); if (rc) return rc;
    AZ_CUDA(cudaMemcpyAsync(e->d_aos, h_data, (size_t)e->n * AZ_DATA_BYTES, cudaMemcpyHostToDevice, s));
    AZ_CUDA(cudaMemsetAsync(e->d_bad, 0, sizeof(int), s));
    k_env_import<<<env_grid(e->n), ENV_BLOCK, 0, s>>>(e->d_state, e->n, az_device_tables(), e->d_aos, e->d_bad);
    AZ_CUDA(cudaGetLastError());
    int bad = 0;
    AZ_CUDA(cudaMemcpyAsync(&bad, e->d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    if (bad) { az_set_error("az_env_import_aos: %d game image(s) whose masks/totals disagree with landArmy[]", bad); return AZ_ERR_BAD_STATE; }
    return AZ_OK;
}

extern "C" int az_env_export_aos(az_env* e, uint8_t* h_data, void* stream)
{
    AZ_REQUIRE(e && h_data, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure(&e->d_aos, (size_t)e->n * AZ_DATA_BYTES); if (rc) return rc;
    const size_t img_bytes = (size_t)ENV_BLOCK * ENV_IMG_STRIDE * sizeof(uint32_t);       // with EnvSmem: above the 48 KB default
    static bool smem_set[64] = { false };
    if (!smem_set[e->device & 63]) {
        AZ_CUDA(cudaFuncSetAttribute(k_env_export, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)img_bytes));
        smem_set[e->device & 63] = true;
    }
    k_env_export<<<env_grid(e->n), ENV_BLOCK, img_bytes, s>>>(e->d_state, e->n, az_device_tables(), e->d_aos);
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA(cudaMemcpyAsync(h_data, e->d_aos, (size_t)e->n * AZ_DATA_BYTES, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_env_valid_moves(az_env* e, uint64_t* h_mask, void* stream)
{
    AZ_REQUIRE(e && h_mask, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure(&e->d_valid, (size_t)e->n); if (rc) return rc;
    k_env_query<<<env_grid(e->n), ENV_BLOCK, 0, s>>>(e->d_state, e->n, az_device_tables(), e->d_valid, nullptr, dev_rules(e->rules));
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA(cudaMemcpyAsync(h_mask, e->d_valid, sizeof(uint64_t) * (size_t)e->n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_env_status(az_env* e, int8_t* h_status, void* stream)
{
    AZ_REQUIRE(e && h_status, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure(&e->d_status, (size_t)e->n); if (rc) return rc;
    k_env_query<<<env_grid(e->n), ENV_BLOCK, 0, s>>>(e->d_state, e->n, az_device_tables(), nullptr, e->d_status, dev_rules(e->rules));
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA(cudaMemcpyAsync(h_status, e->d_status, (size_t)e->n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_env_step_dev(az_env* e, const uint8_t* d_action, const uint8_t* d_dice, int8_t* d_status,
                               uint64_t* d_valid_after, void* stream)
{
    AZ_REQUIRE(e && d_action && d_status, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    AZ_CUDA(cudaEventRecord(e->ev0, s));
    k_env_step<<<env_grid(e->n), ENV_BLOCK, 0, s>>>(e->d_state, e->n, az_device_tables(), d_action, d_dice, d_status, d_valid_after,
                                                     e->seed, e->first_game, dev_rules(e->rules));
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA(cudaEventRecord(e->ev1, s));
    e->timed = true;
    return AZ_OK;
}

extern "C" int az_env_step(az_env* e, const uint8_t* h_action, const uint8_t* h_dice, int8_t* h_status, void* stream)
{
    AZ_REQUIRE(e && h_action && h_status, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure(&e->d_action, (size_t)e->n); if (rc) return rc;
    rc = ensure(&e->d_status, (size_t)e->n); if (rc) return rc;
    AZ_CUDA(cudaMemcpyAsync(e->d_action, h_action, (size_t)e->n, cudaMemcpyHostToDevice, s));
    if (h_dice) {
        rc = ensure(&e->d_dice, (size_t)e->n * 5); if (rc) return rc;
        AZ_CUDA(cudaMemcpyAsync(e->d_dice, h_dice, (size_t)e->n * 5, cudaMemcpyHostToDevice, s));
    }
    rc = az_env_step_dev(e, e->d_action, h_dice ? e->d_dice : nullptr, e->d_status, nullptr, stream);
    if (rc) return rc;
    AZ_CUDA(cudaMemcpyAsync(h_status, e->d_status, (size_t)e->n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

// one turn of every running game; kind0 / kind1 = AZ_OPPONENT_SCRIPT or AZ_OPPONENT_RANDOM for side 0 / side 1
static int env_play_turn(az_env* e, int kind0, int kind1, uint32_t* h_script, int8_t* h_status, cudaStream_t s)
{
    const bool scripted = kind0 == AZ_OPPONENT_SCRIPT || kind1 == AZ_OPPONENT_SCRIPT;
    int rc = ensure(&e->d_status, (size_t)e->n); if (rc) return rc;
    if (scripted) {
        rc = ensure(&e->d_script, (size_t)e->n * 2); if (rc) return rc;
        AZ_CUDA(cudaMemcpyAsync(e->d_script, h_script, sizeof(uint32_t) * 2 * (size_t)e->n, cudaMemcpyHostToDevice, s));
    }
    k_env_script_turn<<<env_grid(e->n), ENV_BLOCK, 0, s>>>(e->d_state, e->n, az_device_tables(), scripted ? e->d_script : nullptr, e->d_status,
                                                            e->seed, e->first_game, dev_rules(e->rules), kind0, kind1, e->rec);
    AZ_CUDA(cudaGetLastError());
    if (e->rec.st) {
        k_env_rec_end<<<(e->n + REC_END_WARPS - 1) / REC_END_WARPS, REC_END_WARPS * 32, 0, s>>>(e->rec, e->d_status, e->n, e->d_rec_out,
                                                                                                  e->d_rec_count, (unsigned long long)e->rec_cap);
        AZ_CUDA(cudaGetLastError());
    }
    if (scripted) AZ_CUDA(cudaMemcpyAsync(h_script, e->d_script, sizeof(uint32_t) * 2 * (size_t)e->n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(h_status, e->d_status, (size_t)e->n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_env_script_turn(az_env* e, uint32_t* h_script, int8_t* h_status, void* stream)
{
    AZ_REQUIRE(e && h_script && h_status, "NULL argument");
    AzDeviceGuard guard(e->device);
    return env_play_turn(e, AZ_OPPONENT_SCRIPT, AZ_OPPONENT_SCRIPT, h_script, h_status, (cudaStream_t)stream);
}

extern "C" int az_env_random_turn(az_env* e, int8_t* h_status, void* stream)
{
    AZ_REQUIRE(e && h_status, "NULL argument");
    AzDeviceGuard guard(e->device);
    return env_play_turn(e, AZ_OPPONENT_RANDOM, AZ_OPPONENT_RANDOM, nullptr, h_status, (cudaStream_t)stream);
}

extern "C" int az_env_play_turn(az_env* e, int kind_side0, int kind_side1, uint32_t* h_script, int8_t* h_status, void* stream)
{
    AZ_REQUIRE(e && h_status, "NULL argument");
    AZ_REQUIRE((kind_side0 == AZ_OPPONENT_SCRIPT || kind_side0 == AZ_OPPONENT_RANDOM) &&
               (kind_side1 == AZ_OPPONENT_SCRIPT || kind_side1 == AZ_OPPONENT_RANDOM), "kind: AZ_OPPONENT_SCRIPT or AZ_OPPONENT_RANDOM");
    AZ_REQUIRE(h_script || (kind_side0 == AZ_OPPONENT_RANDOM && kind_side1 == AZ_OPPONENT_RANDOM), "h_script is NULL but a side is scripted");
    AzDeviceGuard guard(e->device);
    return env_play_turn(e, kind_side0, kind_side1, h_script, h_status, (cudaStream_t)stream);
}

// ---- samples of scripted / random turns (Player::addTrainingSample; the data of AlphaZeroTrainer::trainOnGeneratedData)
extern "C" int az_env_record_turns(az_env* e, size_t capacity_samples, int max_samples_per_game)
{
    AZ_REQUIRE(e != nullptr, "env is NULL");
    AZ_REQUIRE(capacity_samples > 0 && max_samples_per_game > 0, "capacity_samples and max_samples_per_game must be positive");
    AZ_REQUIRE(e->rec.st == nullptr, "recording is already enabled on this handle");
    AzDeviceGuard guard(e->device);
    const size_t n = (size_t)e->n, m = (size_t)max_samples_per_game;
    uint32_t* st = nullptr;
    cudaError_t ce = cudaMalloc(&st, sizeof(uint32_t) * n * m * AZ_PRIMARY_WORDS);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->rec.mv, n * m);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->rec.len, sizeof(uint32_t) * n);
    if (ce == cudaSuccess) ce = cudaMemset(e->rec.len, 0, sizeof(uint32_t) * n);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_rec_out, capacity_samples * (size_t)AZ_SAMPLE_BYTES);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_rec_count, 2 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMemset(e->d_rec_count, 0, 2 * sizeof(unsigned long long));
    if (ce != cudaSuccess) {
        az_set_error("CUDA error %s while allocating the sample staging", cudaGetErrorString(ce));
        cudaGetLastError();
        cudaFree(st); cudaFree(e->rec.mv); cudaFree(e->rec.len); cudaFree(e->d_rec_out); cudaFree(e->d_rec_count);
        e->rec.mv = nullptr; e->rec.len = nullptr; e->d_rec_out = nullptr; e->d_rec_count = nullptr;
        return AZ_ERR_CUDA;
    }
    e->rec.max = max_samples_per_game; e->rec_cap = capacity_samples;
    e->rec.st = st;                                      // set last: the kernels record iff rec.st != NULL
    return AZ_OK;
}

extern "C" int az_env_turn_samples(az_env* e, uint8_t* h_records, size_t max_records, size_t* n_out, uint64_t* h_dropped, void* stream)
{
    AZ_REQUIRE(e && n_out, "NULL argument");
    AZ_REQUIRE(e->rec.st != nullptr, "az_env_record_turns has not been called");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long h[2];
    AZ_CUDA(cudaMemcpyAsync(h, e->d_rec_count, sizeof h, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    AZ_REQUIRE(h[0] <= e->rec_cap, "sample queue corrupted: committed count exceeds the capacity");
    const size_t have = (size_t)h[0];                    // committed records: az_rec_reserve never reserves past the capacity
    if (h_dropped) *h_dropped = h[1];
    *n_out = have;
    if (!h_records && have) return AZ_OK;                // size query; with an empty queue the call drains (resets the dropped count)
    AZ_REQUIRE(max_records >= have, "h_records is too small: query the size with h_records = NULL first");
    if (have) AZ_CUDA(cudaMemcpyAsync(h_records, e->d_rec_out, have * (size_t)AZ_SAMPLE_BYTES, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemsetAsync(e->d_rec_count, 0, sizeof h, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_env_encode_dev(az_env* e, float* d_x, void* stream)
{
    AZ_REQUIRE(e && d_x, "NULL argument");
    AzDeviceGuard guard(e->device);
    k_env_encode<<<(e->n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(e->d_state, e->n, d_x);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_env_encode(az_env* e, float* h_x, void* stream)
{
    AZ_REQUIRE(e && h_x, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure(&e->d_x, (size_t)e->n * AZ_INPUT_FLOATS); if (rc) return rc;
    rc = az_env_encode_dev(e, e->d_x, stream); if (rc) return rc;
    AZ_CUDA(cudaMemcpyAsync(h_x, e->d_x, sizeof(float) * (size_t)e->n * AZ_INPUT_FLOATS, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

// accessors for the other translation units of the library (az_mcts.cu)
int az_launch_encode(const uint32_t* d_state, int n, float* d_x, cudaStream_t s)
{
    k_env_encode<<<(n + 3) / 4, 128, 0, s>>>(d_state, n, d_x);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}
uint32_t* az_env_state_ptr(az_env* e) { return e->d_state; }
int az_env_n(const az_env* e) { return e->n; }
int az_env_device(const az_env* e) { return e->device; }
uint64_t az_env_seed(const az_env* e) { return e->seed; }
uint32_t az_env_first_game(const az_env* e) { return e->first_game; }
const az_rules* az_env_rules(const az_env* e) { return &e->rules; }

extern "C" int az_env_rollout(az_env* e, int n_steps, void* stream)
{
    AZ_REQUIRE(e != nullptr, "env is NULL");
    AZ_REQUIRE(n_steps >= 0, "n_steps must be >= 0");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    AZ_CUDA(cudaEventRecord(e->ev0, s));
    static bool smem_set[64] = { false };                      // wide tables + lane columns + game contexts: above the 48 KB default
    if (!smem_set[e->device & 63]) {
        AZ_CUDA(cudaFuncSetAttribute(k_env_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EnvSmemPool)));
        smem_set[e->device & 63] = true;
    }
    k_env_rollout<<<env_grid(e->n), ENV_BLOCK, sizeof(EnvSmemPool), s>>>(e->d_state, e->n, az_device_tables(), n_steps, e->seed, e->first_game,
                                                                        dev_rules(e->rules), e->d_counters);
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA(cudaEventRecord(e->ev1, s));
    e->timed = true;
    return AZ_OK;
}

extern "C" int az_env_counters(az_env* e, az_counters* h_out, int reset, void* stream)
{
    AZ_REQUIRE(e && h_out, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long h[8];
    AZ_CUDA(cudaMemcpyAsync(h, e->d_counters, sizeof h, cudaMemcpyDeviceToHost, s));
    if (reset) AZ_CUDA(cudaMemsetAsync(e->d_counters, 0, sizeof h, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    h_out->steps = h[0]; h_out->games = h[1]; h_out->wins[0] = h[2]; h_out->wins[1] = h[3]; h_out->draws = h[4];
    h_out->illegal = h[5]; h_out->sims = h[6]; h_out->evals = h[7]; h_out->path_nodes = 0;
    return AZ_OK;
}

extern "C" int az_env_last_kernel_ms(az_env* e, float* ms)
{
    AZ_REQUIRE(e && ms, "NULL argument");
    if (!e->timed) { az_set_error("no timed launch yet"); return AZ_ERR_NOT_READY; }
    AzDeviceGuard guard(e->device);
    AZ_CUDA(cudaEventSynchronize(e->ev1));
    AZ_CUDA(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return AZ_OK;
}
