// az_env6.cu — the SIX-PLAYER extension of the lockstep Risk environment (BASELINE.json configs[3]) and its az_env6_* C ABI.
//
// The reference has no six-player game (PLAYER_COUNT = 2, /root/reference/src/risk_game/state/state.h:13), so this is a
// throughput-only extension with NO reference parity: its rules are SIXPLAYER.md, its only checker is oracle/risk6_oracle.c
// (tests/test_env6_gpu.py compares bit for bit).  Everything that has a two-player counterpart follows that counterpart's
// arithmetic (az_game.cuh) and shares its map tables, neighbour-union lookup, fortify-source search and Philox contract.  The
// two-player path (az_env_*) is untouched.
//
// Layout: one thread per game; 28 words per game in HBM as a structure of arrays (word w of game g at state[w * n + g]):
//   words 0..10   army bytes of the 42 lands            words 11..21  owner bytes (seat 0..5)
//   word 22       cards of seats 0..3 (a byte each)     word 23       cards of seats 4, 5 | allow_draw << 16 | attacks << 24
//   word 24       round | cur << 16 | card_sets << 24   word 25       reinf | phase << 8 | mob_from << 16 | mob_to << 24
//   word 26       set-up pools of the six seats (4 bits each)          word 27       ply
// Inside a kernel the land bytes sit in shared-memory columns (bank = lane), six ownership masks + army > 1 + army == 32 in registers.
#include <cstring>
#include <new>

#include "az_common.cuh"

#define E6_BLOCK 128                    // configs[3] has 16384 games per GPU: 128 blocks = four warps on 128 of the 148 SMs (256: 64 SMs)
#include "az_game6.cuh"

struct E6Smem {
    uint64_t tab[AZ_TABLE_U64];
    uint32_t col[33 * E6_BLOCK];        // 11 army words + 11 owner words + 11 fortify-DFS parent words per thread
};
__device__ __forceinline__ void e6_bind(E6Ctx& c, E6Smem& sm)
{
    c.army.base = (uint8_t*)(sm.col) + 4 * threadIdx.x; c.army.stride_bytes = 4 * E6_BLOCK;
    c.owner.base = (uint8_t*)(sm.col + 11 * E6_BLOCK) + 4 * threadIdx.x; c.owner.stride_bytes = 4 * E6_BLOCK;
    c.scratch.base = (uint8_t*)(sm.col + 22 * E6_BLOCK) + 4 * threadIdx.x; c.scratch.stride_bytes = 4 * E6_BLOCK;
}
__device__ __forceinline__ void e6_load(E6Ctx& c, E6Smem& sm, const uint32_t* __restrict__ st, int n, int gi)
{
    e6_bind(c, sm);
    for (int w = 0; w < 22; ++w) sm.col[w * E6_BLOCK + threadIdx.x] = st[(size_t)w * n + gi];
    e6_masks(c);
    e6_unpack(c.g, st[(size_t)22 * n + gi], st[(size_t)23 * n + gi], st[(size_t)24 * n + gi], st[(size_t)25 * n + gi], st[(size_t)26 * n + gi]);
    c.ply = st[(size_t)27 * n + gi];
}
__device__ __forceinline__ void e6_store(const E6Ctx& c, E6Smem& sm, uint32_t* __restrict__ st, int n, int gi)
{
    for (int w = 0; w < 22; ++w) st[(size_t)w * n + gi] = sm.col[w * E6_BLOCK + threadIdx.x];
    const AzGame6& g = c.g;
    st[(size_t)22 * n + gi] = g.cards_lo;
    st[(size_t)23 * n + gi] = (g.cards_hi & 0xffffu) | ((g.allow_draw & 0xffu) << 16) | ((g.attacks & 0xffu) << 24);
    st[(size_t)24 * n + gi] = (g.round & 0xffffu) | (g.cur << 16) | (g.card_sets << 24);
    st[(size_t)25 * n + gi] = (g.reinf & 0xffu) | (g.phase << 8) | (g.mob_from << 16) | (g.mob_to << 24);
    st[(size_t)26 * n + gi] = g.pools & 0xffffffu;
    st[(size_t)27 * n + gi] = c.ply;
}

// ---------------------------------------------------------------- kernels
__device__ __forceinline__ AzTables e6_stage(E6Smem& sm, const uint64_t* __restrict__ g_tab)
{
    for (int i = threadIdx.x; i < AZ_TABLE_U64; i += blockDim.x) sm.tab[i] = g_tab[i];
    __syncthreads();
    return az_tables_from_smem(sm.tab);
}

__global__ void __launch_bounds__(E6_BLOCK) k_env6_reset(uint32_t* __restrict__ st, int n, uint64_t seed, uint32_t first_game)
{
    __shared__ E6Smem sm;
    const int gi = blockIdx.x * E6_BLOCK + threadIdx.x;
    if (gi >= n) return;
    E6Ctx c; e6_bind(c, sm);
    for (int w = 0; w < 22; ++w) sm.col[w * E6_BLOCK + threadIdx.x] = 0;
    e6_new_game(c, seed, first_game + (uint32_t)gi, 0);
    c.ply = 0;
    e6_store(c, sm, st, n, gi);
}

// n_steps uniform-random legal moves per game (action = k-th set bit of the mask, k = mulhi(word 1 of the real-move block,
// popcount); dice = word 0), finished games re-dealt in place: the six-player form of az_env_rollout.  counters: steps, games,
// draws, wins of seats 0..5
__global__ void __launch_bounds__(E6_BLOCK) k_env6_rollout(uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab, int n_steps,
                                                          uint64_t seed, uint32_t first_game, AzRulesDev rules,
                                                          unsigned long long* __restrict__ counters)
{
    __shared__ E6Smem sm;
    const AzTables T = e6_stage(sm, g_tab);
    const int gi = blockIdx.x * E6_BLOCK + threadIdx.x;
    if (gi >= n) return;
    E6Ctx c; e6_load(c, sm, st, n, gi);
    const uint32_t game = first_game + (uint32_t)gi;
    unsigned games = 0, draws = 0;
    for (int s = 0; s < n_steps; ++s) {
        const int stt = e6_status(c.g, rules);
        if (stt != AZ_STATUS_RUNNING) {
            games++;
            if (stt == AZ_STATUS_DRAW) draws++; else atomicAdd(&counters[3 + stt], 1ull);
            e6_new_game(c, seed, game, c.ply);
        }
        const uint64_t valid = e6_valid(c.g, T, rules);
        const az_u32x4 blk = az_rng_block(seed, game, c.ply, AZ_STREAM_REAL, 0);
        const int action = az_nth_set_bit(valid, az_mulhi32(blk.y, (uint32_t)__popcll(valid)));
        { E6DiceWord dice; dice.w = blk.x; e6_move(c, T, rules, action, dice); }
        c.ply++;
    }
    e6_store(c, sm, st, n, gi);
    atomicAdd(&counters[0], (unsigned long long)n_steps);
    if (games) atomicAdd(&counters[1], (unsigned long long)games);
    if (draws) atomicAdd(&counters[2], (unsigned long long)draws);
}

// one host-chosen action per game (the lockstep tests): status byte = AZ_STATUS_ILLEGAL / AZ_STATUS_OVER (state untouched) or the
// game status after the move
__global__ void __launch_bounds__(E6_BLOCK) k_env6_step(uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab,
                                                       const uint8_t* __restrict__ action, uint64_t seed, uint32_t first_game,
                                                       AzRulesDev rules, int8_t* __restrict__ status)
{
    __shared__ E6Smem sm;
    const AzTables T = e6_stage(sm, g_tab);
    const int gi = blockIdx.x * E6_BLOCK + threadIdx.x;
    if (gi >= n) return;
    E6Ctx c; e6_load(c, sm, st, n, gi);
    if (e6_status(c.g, rules) != AZ_STATUS_RUNNING) { status[gi] = AZ_STATUS_OVER; return; }
    const int a = action[gi];
    const uint64_t valid = e6_valid(c.g, T, rules);
    if (a > AZ_SKIP || !((valid >> a) & 1ull)) { status[gi] = AZ_STATUS_ILLEGAL; return; }
    const az_u32x4 blk = az_rng_block(seed, first_game + (uint32_t)gi, c.ply, AZ_STREAM_REAL, 0);
    { E6DiceWord dice; dice.w = blk.x; e6_move(c, T, rules, a, dice); }
    c.ply++;
    e6_store(c, sm, st, n, gi);
    status[gi] = (int8_t)e6_status(c.g, rules);
}

__global__ void __launch_bounds__(E6_BLOCK) k_env6_query(const uint32_t* __restrict__ st, int n, const uint64_t* __restrict__ g_tab,
                                                        AzRulesDev rules, uint64_t* __restrict__ valid, int8_t* __restrict__ status)
{
    __shared__ E6Smem sm;
    const AzTables T = e6_stage(sm, g_tab);
    const int gi = blockIdx.x * E6_BLOCK + threadIdx.x;
    if (gi >= n) return;
    E6Ctx c; e6_load(c, sm, st, n, gi);
    if (valid) valid[gi] = e6_valid(c.g, T, rules);
    if (status) status[gi] = (int8_t)e6_status(c.g, rules);
}

// 108-byte host image (== r6_state of the oracle): army[42], owner[42], cards[6], pool[6], round u16, cur, card_sets, reinf,
// phase, mob_from, mob_to, allow_draw, attacks, 2 pad bytes
__global__ void __launch_bounds__(E6_BLOCK) k_env6_export(const uint32_t* __restrict__ st, int n, uint8_t* __restrict__ img)
{
    __shared__ E6Smem sm;
    const int gi = blockIdx.x * E6_BLOCK + threadIdx.x;
    if (gi >= n) return;
    E6Ctx c; e6_load(c, sm, st, n, gi);
    uint8_t* d = img + (size_t)gi * E6_IMG;
    for (int i = 0; i < AZ_LANDS; ++i) { d[i] = (uint8_t)c.army.get(i); d[42 + i] = (uint8_t)c.owner.get(i); }
    for (uint32_t p = 0; p < E6_PLAYERS; ++p) { d[84 + p] = (uint8_t)e6_cards(c.g, p); d[90 + p] = (uint8_t)e6_pool(c.g, p); }
    d[96] = (uint8_t)(c.g.round & 0xffu); d[97] = (uint8_t)(c.g.round >> 8);
    d[98] = (uint8_t)c.g.cur; d[99] = (uint8_t)c.g.card_sets; d[100] = (uint8_t)c.g.reinf; d[101] = (uint8_t)c.g.phase;
    d[102] = (uint8_t)c.g.mob_from; d[103] = (uint8_t)c.g.mob_to; d[104] = (uint8_t)c.g.allow_draw; d[105] = (uint8_t)c.g.attacks;
    d[106] = 0; d[107] = 0;
}
__global__ void __launch_bounds__(E6_BLOCK) k_env6_import(uint32_t* __restrict__ st, int n, const uint8_t* __restrict__ img, int* __restrict__ bad)
{
    __shared__ E6Smem sm;
    const int gi = blockIdx.x * E6_BLOCK + threadIdx.x;
    if (gi >= n) return;
    const uint8_t* d = img + (size_t)gi * E6_IMG;
    E6Ctx c; e6_bind(c, sm);
    for (int w = 0; w < 22; ++w) sm.col[w * E6_BLOCK + threadIdx.x] = 0;
    bool ok = true;
    for (int i = 0; i < AZ_LANDS; ++i) {
        ok = ok && d[i] >= 1 && d[i] <= AZ_ARMY_MAX && d[42 + i] < E6_PLAYERS;
        c.army.set(i, d[i]); c.owner.set(i, d[42 + i]);
    }
    e6_masks(c);
    AzGame6& g = c.g;
    g.cards_lo = g.cards_hi = 0; g.pools = 0;
    for (uint32_t p = 0; p < E6_PLAYERS; ++p) { e6_set_cards(g, p, d[84 + p]); g.pools |= (uint32_t)(d[90 + p] & 0xfu) << (4 * p); ok = ok && d[90 + p] <= 13; }
    g.round = (uint32_t)d[96] | ((uint32_t)d[97] << 8); g.cur = d[98]; g.card_sets = d[99]; g.reinf = d[100]; g.phase = d[101];
    g.mob_from = d[102]; g.mob_to = d[103]; g.allow_draw = d[104]; g.attacks = d[105];
    ok = ok && g.cur < E6_PLAYERS && g.phase <= AZ_PH_FORTIFY && g.phase != AZ_PH_SETUP_NEUTRAL;
    if (!ok) { atomicAdd(bad, 1); return; }
    c.ply = st[(size_t)27 * n + gi];                              // the slot's move counter is not part of the image
    e6_store(c, sm, st, n, gi);
}

// ---------------------------------------------------------------- C ABI
struct az_env6 {
    int n = 0, device = 0;
    uint32_t first_game = 0;
    uint64_t seed = 0;
    az_rules rules;
    uint32_t* d_state = nullptr;
    unsigned long long* d_counters = nullptr;   // steps, games, draws, wins[6]
    uint8_t* d_img = nullptr; uint8_t* d_action = nullptr; int8_t* d_status = nullptr; uint64_t* d_valid = nullptr; int* d_bad = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

// accessors for az_mcts6.cu
uint32_t* az_env6_state_ptr(az_env6* e) { return e->d_state; }
int az_env6_n(const az_env6* e) { return e->n; }
int az_env6_device(const az_env6* e) { return e->device; }
uint64_t az_env6_seed(const az_env6* e) { return e->seed; }
uint32_t az_env6_first_game(const az_env6* e) { return e->first_game; }
const az_rules* az_env6_rules(const az_env6* e) { return &e->rules; }

static AzRulesDev e6_rules(const az_rules& r)
{
    AzRulesDev d; d.allow_yield = r.allow_yield; d.limit_reinforcement = r.limit_reinforcement; d.limit_attack = r.limit_attack;
    d.max_game_rounds = r.max_game_rounds; d.min_unit_move = r.min_unit_move; return d;
}
static inline int e6_grid(int n) { return (n + E6_BLOCK - 1) / E6_BLOCK; }

extern "C" int az_env6_create(int n_games, const az_rules* rules, int device, uint32_t first_game_id, az_env6** out)
{
    AZ_REQUIRE(out != nullptr, "out is NULL");
    AZ_REQUIRE(n_games > 0, "n_games must be positive");
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { az_set_error("no CUDA device: libaz_b200 has no CPU fallback"); return AZ_ERR_NO_DEVICE; }
    AZ_REQUIRE(device >= 0 && device < have, "device index out of range");
    AzDeviceGuard guard(device);
    int rc = az_upload_tables(); if (rc) return rc;
    az_env6* e = new (std::nothrow) az_env6();
    AZ_REQUIRE(e != nullptr, "out of host memory");
    e->n = n_games; e->device = device; e->first_game = first_game_id;
    if (rules) e->rules = *rules; else az_default_rules(&e->rules);
    const size_t n = (size_t)n_games;
    cudaError_t ce = cudaMalloc(&e->d_state, sizeof(uint32_t) * E6_WORDS * n);
    if (ce == cudaSuccess) ce = cudaMemset(e->d_state, 0, sizeof(uint32_t) * E6_WORDS * n);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_counters, sizeof(unsigned long long) * 16);
    if (ce == cudaSuccess) ce = cudaMemset(e->d_counters, 0, sizeof(unsigned long long) * 16);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_img, n * E6_IMG);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_action, n);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_status, n);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_valid, sizeof(uint64_t) * n);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_bad, sizeof(int));
    if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e->ev1);
    if (ce != cudaSuccess) { az_set_error("az_env6_create: %s", cudaGetErrorString(ce)); az_env6_destroy(e); return AZ_ERR_CUDA; }
    *out = e;
    return AZ_OK;
}

extern "C" int az_env6_destroy(az_env6* e)
{
    if (!e) return AZ_OK;
    AzDeviceGuard guard(e->device);
    cudaFree(e->d_state); cudaFree(e->d_counters); cudaFree(e->d_img); cudaFree(e->d_action); cudaFree(e->d_status); cudaFree(e->d_valid); cudaFree(e->d_bad);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    delete e;
    return AZ_OK;
}

extern "C" int az_env6_reset(az_env6* e, uint64_t seed, void* stream)
{
    AZ_REQUIRE(e != nullptr, "env is NULL");
    AzDeviceGuard guard(e->device);
    e->seed = seed;
    k_env6_reset<<<e6_grid(e->n), E6_BLOCK, 0, (cudaStream_t)stream>>>(e->d_state, e->n, seed, e->first_game);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_env6_rollout(az_env6* e, int n_steps, void* stream)
{
    AZ_REQUIRE(e != nullptr, "env is NULL");
    AZ_REQUIRE(n_steps >= 0, "n_steps must be >= 0");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    AZ_CUDA(cudaEventRecord(e->ev0, s));
    k_env6_rollout<<<e6_grid(e->n), E6_BLOCK, 0, s>>>(e->d_state, e->n, az_device_tables(), n_steps, e->seed, e->first_game, e6_rules(e->rules),
                                                     e->d_counters);
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA(cudaEventRecord(e->ev1, s));
    e->timed = true;
    return AZ_OK;
}

extern "C" int az_env6_last_kernel_ms(az_env6* e, float* ms)
{
    AZ_REQUIRE(e && ms, "NULL argument");
    AZ_REQUIRE(e->timed, "no timed launch yet");
    AzDeviceGuard guard(e->device);
    AZ_CUDA(cudaEventSynchronize(e->ev1));
    AZ_CUDA(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return AZ_OK;
}

extern "C" int az_env6_step(az_env6* e, const uint8_t* h_action, int8_t* h_status, void* stream)
{
    AZ_REQUIRE(e && h_action && h_status, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    AZ_CUDA(cudaMemcpyAsync(e->d_action, h_action, (size_t)e->n, cudaMemcpyHostToDevice, s));
    k_env6_step<<<e6_grid(e->n), E6_BLOCK, 0, s>>>(e->d_state, e->n, az_device_tables(), e->d_action, e->seed, e->first_game, e6_rules(e->rules), e->d_status);
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA(cudaMemcpyAsync(h_status, e->d_status, (size_t)e->n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_env6_query(az_env6* e, uint64_t* h_valid, int8_t* h_status, void* stream)
{
    AZ_REQUIRE(e && (h_valid || h_status), "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    k_env6_query<<<e6_grid(e->n), E6_BLOCK, 0, s>>>(e->d_state, e->n, az_device_tables(), e6_rules(e->rules), h_valid ? e->d_valid : nullptr,
                                                   h_status ? e->d_status : nullptr);
    AZ_CUDA(cudaGetLastError());
    if (h_valid) AZ_CUDA(cudaMemcpyAsync(h_valid, e->d_valid, sizeof(uint64_t) * (size_t)e->n, cudaMemcpyDeviceToHost, s));
    if (h_status) AZ_CUDA(cudaMemcpyAsync(h_status, e->d_status, (size_t)e->n, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_env6_export(az_env6* e, uint8_t* h_images, void* stream)
{
    AZ_REQUIRE(e && h_images, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    k_env6_export<<<e6_grid(e->n), E6_BLOCK, 0, s>>>(e->d_state, e->n, e->d_img);
    AZ_CUDA(cudaGetLastError());
    AZ_CUDA(cudaMemcpyAsync(h_images, e->d_img, (size_t)e->n * E6_IMG, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return AZ_OK;
}

extern "C" int az_env6_import(az_env6* e, const uint8_t* h_images, void* stream)
{
    AZ_REQUIRE(e && h_images, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    AZ_CUDA(cudaMemcpyAsync(e->d_img, h_images, (size_t)e->n * E6_IMG, cudaMemcpyHostToDevice, s));
    AZ_CUDA(cudaMemsetAsync(e->d_bad, 0, sizeof(int), s));
    k_env6_import<<<e6_grid(e->n), E6_BLOCK, 0, s>>>(e->d_state, e->n, e->d_img, e->d_bad);
    AZ_CUDA(cudaGetLastError());
    int bad = 0;
    AZ_CUDA(cudaMemcpyAsync(&bad, e->d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    if (bad) { az_set_error("az_env6_import: %d malformed game image(s)", bad); return AZ_ERR_BAD_STATE; }
    return AZ_OK;
}

extern "C" int az_env6_counters(az_env6* e, az_counters6* h_out, int reset, void* stream)
{
    AZ_REQUIRE(e && h_out, "NULL argument");
    AzDeviceGuard guard(e->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long h[16];
    AZ_CUDA(cudaMemcpyAsync(h, e->d_counters, sizeof h, cudaMemcpyDeviceToHost, s));
    if (reset) AZ_CUDA(cudaMemsetAsync(e->d_counters, 0, sizeof h, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    h_out->steps = h[0]; h_out->games = h[1]; h_out->draws = h[2];
    for (int p = 0; p < E6_PLAYERS; ++p) h_out->wins[p] = h[3 + p];
    return AZ_OK;
}
