// az_game.cuh — device-side Risk game logic (sm_100a), shared by the env kernels (one thread
// per game) and the MCTS descent (one warp per game).
//
// Behaviour follows /root/reference/src/risk_game: state/state.cpp (transitions),
// player/alpha_zero/alphazero_moves.cpp (legal-move mask + action decoding),
// player/game_helper.cpp:51-109 (fortify component DFS).  The design does not: the game is
// held as the ~56-byte PRIMARY state (42 land bytes + 12 scalar bytes); the five 48-bit masks
// per player that the reference maintains incrementally (State::setLandArmy,
// state.cpp:279-385) are never stored.  Four 42-bit register masks (owner 0, owner 1,
// army > 1, army == 32) are kept alongside the land bytes and every derived set the rules
// need is one neighbour-union N(S) away, evaluated with a 7-slice x 64-entry shared-memory
// lookup table (no per-land loop, no divergence).
#pragma once

#include <stdint.h>
#include <cuda_runtime.h>

#include "az_b200.h"
#include "az_philox.h"

#define AZ_ALL_LANDS 0x3ffffffffffull   // LandSet::ALL_LANDS_MASK, land/land_set.cpp:27-33
#define AZ_SKIP_MASK (1ull << AZ_SKIP)  // Land::SKIP_MOVE_MASK, land/land.cpp:313
#define AZ_ARMY_MAX 32                  // LAND_ARMY_MAX, state/state.h:22
#define AZ_NEUTRAL 2                    // NEUTRAL_PLAYER, state/state.h:38

enum { AZ_PH_SETUP = 0, AZ_PH_SETUP_NEUTRAL = 1, AZ_PH_REINFORCEMENT = 2, AZ_PH_ATTACK = 3, AZ_PH_MOBILIZATION = 4, AZ_PH_FORTIFY = 5 };

// packed primary state: 14 x u32 words (byte layout == oracle ro_state)
//   bytes 0..41  land[i] = army | owner << 6        (LandArmy, state/state.h:24-34)
//   bytes 42,43  cards[0], cards[1]                  (PlayerStatus::playerCards, simple-card mode)
//   word 11      round | cur << 16 | card_sets << 24
//   word 12      reinf | phase << 8 | mob_from << 16 | mob_to << 24
//   word 13      allow_draw | attacks << 8
// device SoA adds word 14 = ply (real moves made by this slot, never reset), word 15 = spare.
#define AZ_PRIMARY_WORDS 14
#define AZ_W_PLY 14

// n-th (0-based) set bit of a 42/43-bit mask, n < popcount(m): a 5-step binary search on popcounts
// (__fns is a software loop on sm_100 — it was 13 % of the rollout kernel's instructions)
__device__ __forceinline__ int az_nth_set_bit(uint64_t m, uint32_t n)
{
    uint32_t w = (uint32_t)m;
    const uint32_t pl = (uint32_t)__popc(w);
    int base = 0;
    if (n >= pl) { n -= pl; w = (uint32_t)(m >> 32); base = 32; }
    uint32_t c = (uint32_t)__popc(w & 0xffffu);
    if (n >= c) { n -= c; w >>= 16; base += 16; }
    c = (uint32_t)__popc(w & 0xffu);
    if (n >= c) { n -= c; w >>= 8; base += 8; }
    c = (uint32_t)__popc(w & 0xfu);
    if (n >= c) { n -= c; w >>= 4; base += 4; }
    c = (uint32_t)__popc(w & 0x3u);
    if (n >= c) { n -= c; w >>= 2; base += 2; }
    if (n >= (w & 1u)) base += 1;
    return base;
}

// ---------------------------------------------------------------- constant tables
struct AzTablesConst {
    uint64_t lut[7 * 64];
    uint64_t nbr[42];
    uint64_t list6[42];
};
#define AZ_TABLE_U64 (7 * 64 + 42 + 42)

struct AzTables {           // pointers into shared memory
    const uint64_t* lut;
    const uint64_t* nbr;
    const uint64_t* list6;
};

__device__ __forceinline__ AzTables az_tables_from_smem(const uint64_t* s)
{
    AzTables t; t.lut = s; t.nbr = s + 7 * 64; t.list6 = s + 7 * 64 + 42; return t;
}

// N(S): union of the neighbour masks of the lands in S
__device__ __forceinline__ uint64_t az_nbr_union(const AzTables& T, uint64_t S)
{
    uint64_t u = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) u |= T.lut[k * 64 + (int)((S >> (6 * k)) & 63u)];
    return u;
}

// The 512-move rollout kernel trades shared memory for instructions: four 2048-entry tables indexed by 11-bit slices of S
// (64 KB, built at upload time) instead of seven 64-entry ones.  Everything that takes the tables as a template parameter
// picks the matching az_nbr_union overload.
#define AZ_LUT11_U64 (4 * 2048)
struct AzTablesWide : AzTables {
    const uint64_t* lut11;
};
__device__ __forceinline__ uint64_t az_nbr_union(const AzTablesWide& T, uint64_t S)
{
    const uint32_t lo = (uint32_t)S, hi = (uint32_t)(S >> 32);
    const uint32_t i0 = lo & 2047u, i1 = (lo >> 11) & 2047u, i2 = __funnelshift_r(lo, hi, 22) & 2047u, i3 = (hi >> 1) & 2047u;
    return (T.lut11[i0] | T.lut11[2048 + i1]) | (T.lut11[4096 + i2] | T.lut11[6144 + i3]);
}

struct AzRulesDev {
    int allow_yield, limit_reinforcement, limit_attack, max_game_rounds, min_unit_move;
};

// ---------------------------------------------------------------- game registers
struct AzGame {
    uint64_t own0, own1;     // lands owned by player 0 / 1          (PlayerStatus::ownedLands)
    uint64_t gt1;            // lands with army > 1 (any owner)      (-> ownedLandsWithArmy)
    uint64_t full;           // lands with army == 32 (any owner)    (-> ownedFullLands)
    uint32_t round, cur, card_sets, reinf, phase, mob_from, mob_to, allow_draw, attacks;
    uint32_t cards0, cards1;

    __device__ __forceinline__ uint64_t own(uint32_t p) const { return p ? own1 : own0; }
};

__device__ __forceinline__ void az_unpack_scalars(AzGame& g, uint32_t w10, uint32_t w11, uint32_t w12, uint32_t w13)
{
    g.cards0 = (w10 >> 16) & 0xff; g.cards1 = (w10 >> 24) & 0xff;
    g.round = w11 & 0xffff; g.cur = (w11 >> 16) & 0xff; g.card_sets = (w11 >> 24) & 0xff;
    g.reinf = w12 & 0xff; g.phase = (w12 >> 8) & 0xff; g.mob_from = (w12 >> 16) & 0xff; g.mob_to = (w12 >> 24) & 0xff;
    g.allow_draw = w13 & 0xff; g.attacks = (w13 >> 8) & 0xff;
}
__device__ __forceinline__ uint32_t az_pack_w11(const AzGame& g) { return (g.round & 0xffff) | (g.cur << 16) | (g.card_sets << 24); }
__device__ __forceinline__ uint32_t az_pack_w12(const AzGame& g) { return (g.reinf & 0xff) | (g.phase << 8) | (g.mob_from << 16) | (g.mob_to << 24); }
__device__ __forceinline__ uint32_t az_pack_w13(const AzGame& g) { return (g.allow_draw & 0xff) | ((g.attacks & 0xff) << 8); }

// add the 4 lands of packed word `w` (lands 4*wi .. 4*wi+3) to the register masks
__device__ __forceinline__ void az_masks_add_word(AzGame& g, uint32_t w, int wi)
{
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        int i = wi * 4 + b;
        if (i < AZ_LANDS) {
            uint32_t v = (w >> (8 * b)) & 0xff, a = v & 63, o = v >> 6;
            uint64_t m = 1ull << i;
            if (o == 0) g.own0 |= m;
            if (o == 1) g.own1 |= m;
            if (a > 1) g.gt1 |= m;
            if (a == AZ_ARMY_MAX) g.full |= m;
        }
    }
}

// Land byte storage.  LandT must provide: uint32_t get(int i) const; void set(int i, uint32_t v);
// Column layout for one-thread-per-game kernels: word w of thread t lives at base[w * stride + t],
// so every access of a warp hits 32 different banks whatever the (divergent) land index is.
struct AzLandColumn {
    uint8_t* base;      // (uint8_t*)(smem words) + 4 * threadIdx
    int stride_bytes;   // 4 * blockDim
    __device__ __forceinline__ uint32_t get(int i) const { return base[(i >> 2) * stride_bytes + (i & 3)]; }
    __device__ __forceinline__ void set(int i, uint32_t v) { base[(i >> 2) * stride_bytes + (i & 3)] = (uint8_t)v; }
};
// Row layout (contiguous bytes) for the warp-per-game MCTS kernels
struct AzLandRow {
    uint8_t* base;
    __device__ __forceinline__ uint32_t get(int i) const { return base[i]; }
    __device__ __forceinline__ void set(int i, uint32_t v) { base[i] = (uint8_t)v; }
};

template <class LandT>
__device__ __forceinline__ void az_set_land(AzGame& g, LandT& land, int i, uint32_t army, uint32_t owner)
{
    land.set(i, (army & 63u) | (owner << 6));
    uint64_t m = 1ull << i;
    g.own0 = owner == 0 ? (g.own0 | m) : (g.own0 & ~m);
    g.own1 = owner == 1 ? (g.own1 | m) : (g.own1 & ~m);
    g.gt1 = army > 1 ? (g.gt1 | m) : (g.gt1 & ~m);
    g.full = army == AZ_ARMY_MAX ? (g.full | m) : (g.full & ~m);
}

// the same write when the land keeps its owner (`owner` is only repeated into the byte): the ownership masks do not change
template <class LandT>
__device__ __forceinline__ void az_set_army(AzGame& g, LandT& land, int i, uint32_t army, uint32_t owner)
{
    land.set(i, (army & 63u) | (owner << 6));
    uint64_t m = 1ull << i;
    g.gt1 = army > 1 ? (g.gt1 | m) : (g.gt1 & ~m);
    g.full = army == AZ_ARMY_MAX ? (g.full | m) : (g.full & ~m);
}

// ---------------------------------------------------------------- rules
// State::calculateReinforcementValue, state/state.cpp:457-491
__device__ __forceinline__ int az_reinforcement_value(uint64_t owned)
{
    int v = __popcll(owned) / 3;
    v += ((owned & 0x1ffull) == 0x1ffull) ? 5 : 0;                    // North America
    v += ((owned & 0x1e00ull) == 0x1e00ull) ? 2 : 0;                  // South America
    v += ((owned & 0xfe000ull) == 0xfe000ull) ? 5 : 0;                // Europe
    v += ((owned & 0x3f00000ull) == 0x3f00000ull) ? 3 : 0;            // Africa
    v += ((owned & 0x3ffc000000ull) == 0x3ffc000000ull) ? 7 : 0;      // Asia
    v += ((owned & 0x3c000000000ull) == 0x3c000000000ull) ? 2 : 0;    // Australia
    return v < 3 ? 3 : v;
}

// State::gameStatus, state/state.cpp:518-565
__device__ __forceinline__ int az_game_status(const AzGame& g, const AzRulesDev& r)
{
    int p0 = __popcll(g.own0), p1 = __popcll(g.own1);
    if (p0 == 0) return 1;
    if (p1 == 0) return 0;
    if (r.allow_yield) { if (p0 >= 30) return 0; if (p1 >= 30) return 1; }
    if ((int)g.round > r.max_game_rounds) return p0 > p1 ? 0 : (p0 < p1 ? 1 : AZ_STATUS_DRAW);
    return AZ_STATUS_RUNNING;
}

// PlayerStatus::attackLandsWithArmy of player p: N(owned ∧ army>1) \ owned
template <class TabT>
__device__ __forceinline__ uint64_t az_attack_army(const AzGame& g, const TabT& T, uint32_t p)
{
    uint64_t o = g.own(p);
    return az_nbr_union(T, o & g.gt1) & ~o;
}

// UtilityNN::getValidMoves, player/alpha_zero/alphazero_moves.cpp:3-70
__device__ __forceinline__ uint64_t az_valid_moves(const AzGame& g, const AzTables& T, const AzRulesDev& r)
{
    uint64_t oc = g.own(g.cur), oe = g.own(g.cur ^ 1u);
    switch (g.phase) {
    case AZ_PH_SETUP:
    case AZ_PH_REINFORCEMENT: {
        uint64_t o = oc & ~g.full;
        if (o == 0) return AZ_SKIP_MASK;
        if (r.limit_reinforcement) {
            // o ∧ (enemy.attackLands ∪ neutralAttackLands) == o ∧ N(all lands that are not mine)
            uint64_t b = o & az_nbr_union(T, AZ_ALL_LANDS & ~oc);
            return b ? b : o;
        }
        return o;
    }
    case AZ_PH_SETUP_NEUTRAL: return AZ_ALL_LANDS & ~oc & ~oe;
    case AZ_PH_ATTACK: {
        uint64_t aa = az_attack_army(g, T, g.cur);
        if (r.limit_attack) return aa ? aa : AZ_SKIP_MASK;
        return aa | AZ_SKIP_MASK;
    }
    case AZ_PH_MOBILIZATION: return (1ull << g.mob_from) | (1ull << g.mob_to);
    default: // FORTIFY: owned ∧ enemy.attackLands == owned ∧ N(enemy lands)
        if (r.limit_reinforcement) return (oc & az_nbr_union(T, oe)) | AZ_SKIP_MASK;
        return oc | AZ_SKIP_MASK;
    }
}

// State::gotoAttack, state/state.cpp:20-40
__device__ __forceinline__ void az_goto_attack(AzGame& g, const AzTables& T)
{
    g.phase = AZ_PH_ATTACK; g.mob_from = AZ_NONE; g.mob_to = AZ_NONE; g.reinf = 0;
    if (az_attack_army(g, T, g.cur) == 0) g.phase = AZ_PH_FORTIFY;
}

// State::nextPlayerGameTurn, state/state.cpp:748-766 (+ drawCard, simple-card mode, :618-626)
__device__ __forceinline__ void az_end_turn(AzGame& g)
{
    if (g.allow_draw) {
        if (g.cur) g.cards1 = (g.cards1 + 1) & 0xff; else g.cards0 = (g.cards0 + 1) & 0xff;
        g.allow_draw = 0;
    }
    g.round = (g.round + 1) & 0xffff;
    g.cur ^= 1u;
    g.attacks = 0;
    g.phase = AZ_PH_REINFORCEMENT;
    g.reinf = (uint32_t)az_reinforcement_value(g.own(g.cur));
}

// Dice sources.  next() returns 1..6 in the reference's order of consumption
// (attacker dice then defender dice, state.cpp:832-833).
struct AzDicePhilox {                 // include/az_philox.h contract
    uint64_t seed; uint32_t game, ply, sim, j, w; az_u32x4 blk; bool have_blk;
    __device__ __forceinline__ void init(uint64_t s, uint32_t g, uint32_t p, uint32_t sm) { seed = s; game = g; ply = p; sim = sm; j = 0; w = 0; have_blk = false; }
    __device__ __forceinline__ void init_with_block0(uint64_t s, uint32_t g, uint32_t p, uint32_t sm, az_u32x4 b) { init(s, g, p, sm); blk = b; have_blk = true; }
    __device__ __forceinline__ int next()
    {
        if (j % 5u == 0) {
            if (j % 20u == 0 && !(j == 0 && have_blk)) blk = az_rng_block(seed, game, ply, sim, j / 20u);
            w = az_u32x4_word(blk, (int)((j / 5u) & 3u));
        }
        uint64_t p = (uint64_t)w * 6u;
        w = (uint32_t)p; ++j;
        return (int)(p >> 32) + 1;
    }
};
struct AzDiceTape {                   // explicit dice (golden vectors)
    const uint8_t* t; int j;
    __device__ __forceinline__ int next() { int v = t[j < 5 ? j : 4]; ++j; return v; }
};

// fortify source selection: UtilityNN::makeMove FORTIFY branch (alphazero_moves.cpp:172-231) over the
// owned connected component of `li` listed in the DFS PRE-ORDER of GameHelper::LandSetMovement::add
// (game_helper.cpp:51-82; seed = lowest owned index of the component, children in neighbour-list
// order): the source is the land (!= li) with the most movable armies, interior lands (all neighbours
// owned) before border lands, FIRST in pre-order among equals.
//
// The pre-order only matters for ties, so the maximum is found first with mask arithmetic: candidates =
// component lands with army >= 2, restricted to interior ones if any exist; if a single land holds the
// maximum it is the answer and no walk is needed.  Otherwise the DFS runs until it first visits one of the
// tied lands.  (The walk replaces the reference's recursion by parent pointers: after returning to a
// node, rescanning its list from the start finds the same next child because everything before it is
// already seen.)
template <class LandT, class ScratchT, class TabT>
__device__ __forceinline__ void az_fortify_source(const AzGame& g, const LandT& land, ScratchT& parent, const TabT& T,
                                                  int li, int& from_out, int& amount_out)
{
    const uint64_t owned = g.own(g.cur);
    uint64_t comp = 1ull << li;
    // Only lands with army >= 2 can be the source, so the closure may stop as soon as it has absorbed every such land the
    // mover owns (`rest` empty); the full component is only needed by the tie walk below, which then finishes it.
    const uint64_t movable = owned & g.gt1 & ~(1ull << li);
    from_out = -1; amount_out = 0;
    if (movable == 0) return;
    bool closed = false;
    for (;;) {     // (a per-land BFS has fewer instructions but a longer dependent chain: measured 4 % slower)
        uint64_t n = (comp | az_nbr_union(T, comp)) & owned;
        if (n == comp) { closed = true; break; }
        comp = n;
        if ((movable & ~comp) == 0) break;
    }
    uint64_t cand = comp & movable;                                     // army - 1 > 0
    if (cand == 0) return;
    const uint64_t inter = cand & ~az_nbr_union(T, AZ_ALL_LANDS & ~owned);   // not adjacent to any land the mover does not own
    if (inter) cand = inter;
    int best = 0;
    uint64_t tie = 0;
    for (uint64_t k = cand; k; k &= k - 1) {
        const int v = __ffsll((long long)k) - 1;
        const int a = (int)(land.get(v) & 63u);
        if (a > best) { best = a; tie = 0; }
        if (a == best) tie |= 1ull << v;
    }
    amount_out = best - 1;
    if ((tie & (tie - 1)) == 0) { from_out = __ffsll((long long)tie) - 1; return; }
    // several lands hold the maximum: the first one in DFS pre-order wins
    if (!closed) for (;;) { uint64_t n = (comp | az_nbr_union(T, comp)) & owned; if (n == comp) break; comp = n; }
    uint64_t seen = 0;
    int v = __ffsll((long long)comp) - 1;
    for (;;) {
        if ((tie >> v) & 1ull) { from_out = v; return; }
        seen |= 1ull << v;
        for (;;) {                                                       // next unvisited owned land in DFS order
            const uint64_t nxt = T.nbr[v] & owned & ~seen;
            if (nxt) {
                uint64_t lst = T.list6[v];
                int u = (int)(lst & 63u);
                while (!((nxt >> u) & 1ull)) { lst >>= 6; u = (int)(lst & 63u); }
                parent.set(u, (uint32_t)v);
                v = u;
                break;
            }
            v = (int)parent.get(v);   // backtrack (never past the seed: an unvisited tied land guarantees a candidate upstream)
        }
    }
}

// UtilityNN::makeMove, player/alpha_zero/alphazero_moves.cpp:72-233.  `valid` = az_valid_moves(g).
// Returns AZ_STATUS_ILLEGAL (state untouched) if the action is not a legal move, else 0.
template <class LandT, class ScratchT, class DiceT>
__device__ __forceinline__ int az_make_move(AzGame& g, LandT& land, ScratchT& scratch, const AzTables& T, const AzRulesDev& r,
                                            uint64_t valid, int action, DiceT& dice)
{
    if (action < 0 || action > AZ_SKIP || !((valid >> action) & 1ull)) return AZ_STATUS_ILLEGAL;
    const uint32_t cur = g.cur;
    if (action == AZ_SKIP) {                                  // alphazero_moves.cpp:79-92
        if (g.phase == AZ_PH_REINFORCEMENT) az_goto_attack(g, T);
        else if (g.phase == AZ_PH_ATTACK) g.phase = AZ_PH_FORTIFY;
        else if (g.phase == AZ_PH_FORTIFY) az_end_turn(g);
        else return AZ_STATUS_ILLEGAL;                        // reference: logic_error
        return 0;
    }
    const int li = action;
    switch (g.phase) {
    case AZ_PH_SETUP: {                                       // State::setupReinforcementMove, state.cpp:1009-1030
        g.reinf = (g.reinf - 2) & 0xff;
        az_set_land(g, land, li, (land.get(li) & 63u) + 2, cur);
        g.phase = AZ_PH_SETUP_NEUTRAL;
        break;
    }
    case AZ_PH_SETUP_NEUTRAL: {                               // setupReinforcementNeutralMove :1032-1053 + nextPlayerSetupTurn :725-746
        az_set_land(g, land, li, (land.get(li) & 63u) + 1, AZ_NEUTRAL);
        g.phase = AZ_PH_SETUP; g.round = (g.round + 1) & 0xffff; g.cur ^= 1u;
        if (g.reinf == 0) { g.phase = AZ_PH_REINFORCEMENT; g.reinf = (uint32_t)az_reinforcement_value(g.own(g.cur)); }
        break;
    }
    case AZ_PH_REINFORCEMENT: {                               // alphazero_moves.cpp:104-121, game_helper.cpp:3-17, state.cpp:1091-1117
        uint32_t cards = cur ? g.cards1 : g.cards0;
        if (cards >= 3) {
            cards -= 3;
            if (cur) g.cards1 = cards; else g.cards0 = cards;
            g.card_sets = (g.card_sets + 1) & 0xff;
            int cs = (int)g.card_sets;
            int gained = cs <= 5 ? 2 + 2 * cs : 15 + (cs - 6) * 5;
            g.reinf = (g.reinf + (uint32_t)gained) & 0xff;
        }
        int rf = (int)g.reinf / 2;                            // FAST_ATTACK_MOBILIZATION branch
        if (rf < r.min_unit_move) rf = r.min_unit_move < (int)g.reinf ? r.min_unit_move : (int)g.reinf;
        int army = (int)(land.get(li) & 63u);
        int space = AZ_ARMY_MAX - army;
        if (space < rf) rf = space;
        g.reinf = (g.reinf - (uint32_t)rf) & 0xff;            // State::reinforcementMove, state.cpp:976-998
        az_set_land(g, land, li, (uint32_t)(army + rf), cur);
        if (g.reinf == 0) az_goto_attack(g, T);
        break;
    }
    case AZ_PH_ATTACK: {                                      // alphazero_moves.cpp:122-145, State::attackMove state.cpp:769-918
        int best = 0, from = -1;
        {
            uint64_t lst = T.list6[li];
            uint64_t cand = g.own(cur) & g.gt1;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                int n = (int)(lst & 63u); lst >>= 6;
                if (n != 63 && ((cand >> n) & 1ull)) {
                    int v = (int)(land.get(n) & 63u) - 1;
                    if (v > best) { best = v; from = n; }
                }
            }
        }
        if (from < 0) return AZ_STATUS_ILLEGAL;
        g.attacks = (g.attacks + 1) & 0xff;
        uint32_t tob = land.get(li);
        int a = (int)(land.get(from) & 63u), d = (int)(tob & 63u), units = 1;
        const uint32_t defender = tob >> 6;
        if (d > 0) {
            const int na = a >= 4 ? 3 : (a == 3 ? 2 : 1);
            const int nd = d >= 2 ? 2 : 1;
            units = na;
            int a0 = dice.next(), a1 = 0, a2 = 0;
            if (na > 1) a1 = dice.next();
            if (na > 2) a2 = dice.next();
            int d0 = dice.next(), d1 = 0;
            if (nd > 1) d1 = dice.next();
            // two highest attacker dice, sorted defender dice (State::getDiceRolls, state.cpp:645-684)
            int hi = max(a0, max(a1, a2));
            int lo = min(a0, max(a1, a2)); lo = max(lo, min(a1, a2));       // median of three = second highest
            int dh = max(d0, d1), dl = min(d0, d1);
            if (hi > dh) d--; else { a--; units--; }
            if (na >= 2 && nd == 2) { if (lo > dl) d--; else { a--; units--; } }
        }
        if (d == 0) {
            a -= units;
            if (a > 1) { g.phase = AZ_PH_MOBILIZATION; g.mob_from = (uint32_t)from; g.mob_to = (uint32_t)li; }
            g.allow_draw = 1;
            az_set_land(g, land, from, (uint32_t)a, cur);
            az_set_land(g, land, li, (uint32_t)units, cur);
        } else {
            az_set_land(g, land, from, (uint32_t)a, cur);
            az_set_land(g, land, li, (uint32_t)d, defender);
        }
        if (g.phase == AZ_PH_ATTACK && az_attack_army(g, T, cur) == 0) g.phase = AZ_PH_FORTIFY;
        break;
    }
    case AZ_PH_MOBILIZATION: {                                // alphazero_moves.cpp:146-171, State::attackReinforcementMove state.cpp:920-947
        if ((uint32_t)li == g.mob_from) az_goto_attack(g, T);
        else {
            int from = (int)g.mob_from, to = (int)g.mob_to;
            int af = (int)(land.get(from) & 63u), at = (int)(land.get(to) & 63u);
            int v = af - 1;
            int rf = v / 2;
            if (rf < r.min_unit_move) rf = r.min_unit_move < v ? r.min_unit_move : v;
            az_set_land(g, land, from, (uint32_t)(af - rf), cur);
            az_set_land(g, land, to, (uint32_t)(at + rf), cur);
            if (af - rf == 1) az_goto_attack(g, T);
        }
        break;
    }
    default: {                                                // FORTIFY, alphazero_moves.cpp:172-231
        int at = (int)(land.get(li) & 63u);
        if (at != AZ_ARMY_MAX) {
            int from, amount;
            az_fortify_source(g, land, scratch, T, li, from, amount);
            if (from >= 0) {
                int space = AZ_ARMY_MAX - at;
                int mv = space < amount ? space : amount;     // State::fortifyMove, state.cpp:949-974
                int af = (int)(land.get(from) & 63u);
                az_set_land(g, land, from, (uint32_t)(af - mv), cur);
                az_set_land(g, land, li, (uint32_t)(at + mv), cur);
            }
        }
        az_end_turn(g);
        break;
    }
    }
    return 0;
}

// ---------------------------------------------------------------- one-instruction-stream variants for the thread-per-game rollout
// A warp of the rollout kernel holds 32 games in (up to) six different phases.  az_valid_moves / az_make_move branch per
// phase, so the warp executes the SUM of the phase paths with ~4 of 32 lanes active (ncu, round 1).  The two functions
// below compute the same results with the expensive parts (neighbour-union table walks, land writes, the attack-army
// check) hoisted into code every lane runs once; only short phase-specific arithmetic stays under branches.

// == az_valid_moves (UtilityNN::getValidMoves, alphazero_moves.cpp:3-70) with a single neighbour union
template <class TabT>
__device__ __forceinline__ uint64_t az_valid_moves_flat(const AzGame& g, const TabT& T, const AzRulesDev& r)
{
    const uint64_t oc = g.own(g.cur), oe = g.own(g.cur ^ 1u);
    const uint32_t ph = g.phase;
    const uint64_t s_in = ph == AZ_PH_ATTACK ? (oc & g.gt1) : (ph == AZ_PH_FORTIFY ? oe : (AZ_ALL_LANDS & ~oc));
    const uint64_t u = az_nbr_union(T, s_in);
    const uint64_t o = oc & ~g.full;
    uint64_t place = o;                                           // SETUP / REINFORCEMENT
    if (r.limit_reinforcement && (o & u)) place = o & u;
    if (o == 0) place = AZ_SKIP_MASK;
    const uint64_t aa = u & ~oc;                                  // ATTACK
    const uint64_t att = r.limit_attack ? (aa ? aa : AZ_SKIP_MASK) : (aa | AZ_SKIP_MASK);
    const uint64_t fort = (r.limit_reinforcement ? (oc & u) : oc) | AZ_SKIP_MASK;
    const uint64_t mob = (1ull << (g.mob_from & 63u)) | (1ull << (g.mob_to & 63u));
    uint64_t v = place;
    v = ph == AZ_PH_SETUP_NEUTRAL ? (AZ_ALL_LANDS & ~oc & ~oe) : v;
    v = ph == AZ_PH_ATTACK ? att : v;
    v = ph == AZ_PH_MOBILIZATION ? mob : v;
    v = ph == AZ_PH_FORTIFY ? fort : v;
    return v;
}

// == az_make_move for a LEGAL action (UtilityNN::makeMove, alphazero_moves.cpp:72-233), dice = the base-6 digits of
// `dice_word` (word 0 of the real-move Philox block: at most 5 dice per move, include/az_philox.h).
// Returns 1 WITHOUT touching the state when the move is a FORTIFY onto a non-full land: that one needs the owned
// component search (az_fortify_source), which the rollout kernel batches over several lanes; returns 0 otherwise.
template <class LandT, class TabT>
__device__ __forceinline__ int az_move_flat(AzGame& g, LandT& land, const TabT& T, const AzRulesDev& r, int action, uint32_t dice_word)
{
    const uint32_t cur = g.cur, ph = g.phase;
    const bool skip = action == AZ_SKIP;
    const int li = skip ? 0 : action;
    const uint32_t tob = land.get(li);
    const int at = (int)(tob & 63u);
    if (ph == AZ_PH_FORTIFY && !skip && at != AZ_ARMY_MAX) return 1;

    bool ga = false, et = false;                 // State::gotoAttack / nextPlayerGameTurn after the land writes
    int ia = li, ib = li;
    uint32_t va = 0, vb = 0, ob = cur;
    bool wa = false, wb = false, captured = false;
    if (skip) {                                                    // alphazero_moves.cpp:79-92
        ga = ph == AZ_PH_REINFORCEMENT;
        et = ph == AZ_PH_FORTIFY;
        if (ph == AZ_PH_ATTACK) g.phase = AZ_PH_FORTIFY;
    } else if (ph == AZ_PH_ATTACK) {                               // alphazero_moves.cpp:122-145, State::attackMove state.cpp:769-918
        int best = 0, from = li;
        uint64_t lst = T.list6[li];
        const uint64_t cand = g.own(cur) & g.gt1;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int n = (int)(lst & 63u); lst >>= 6;
            const bool ok = n != 63 && ((cand >> n) & 1ull);
            const int v = ok ? (int)(land.get(ok ? n : li) & 63u) - 1 : 0;
            if (v > best) { best = v; from = n; }
        }
        g.attacks = (g.attacks + 1) & 0xff;
        int a = best + 1, d = at, units = 1;
        const uint32_t defender = tob >> 6;
        if (d > 0) {
            const int na = a >= 4 ? 3 : (a == 3 ? 2 : 1);
            const int nd = d >= 2 ? 2 : 1;
            units = na;
            uint32_t w = dice_word;
            uint64_t p;
            p = (uint64_t)w * 6u; w = (uint32_t)p; const int a0 = (int)(p >> 32) + 1;
            int a1 = 0, a2 = 0, d1 = 0;
            if (na > 1) { p = (uint64_t)w * 6u; w = (uint32_t)p; a1 = (int)(p >> 32) + 1; }
            if (na > 2) { p = (uint64_t)w * 6u; w = (uint32_t)p; a2 = (int)(p >> 32) + 1; }
            p = (uint64_t)w * 6u; w = (uint32_t)p; const int d0 = (int)(p >> 32) + 1;
            if (nd > 1) { p = (uint64_t)w * 6u; w = (uint32_t)p; d1 = (int)(p >> 32) + 1; }
            const int hi = max(a0, max(a1, a2));
            int lo = min(a0, max(a1, a2)); lo = max(lo, min(a1, a2));
            const int dh = max(d0, d1), dl = min(d0, d1);
            if (hi > dh) d--; else { a--; units--; }
            if (na >= 2 && nd == 2) { if (lo > dl) d--; else { a--; units--; } }
        }
        ia = from; ib = li; wa = wb = true;
        if (d == 0) {
            a -= units;
            if (a > 1) { g.phase = AZ_PH_MOBILIZATION; g.mob_from = (uint32_t)from; g.mob_to = (uint32_t)li; }
            g.allow_draw = 1;
            va = (uint32_t)a; vb = (uint32_t)units; captured = true;
        } else { va = (uint32_t)a; vb = (uint32_t)d; ob = defender; }
    } else if (ph == AZ_PH_MOBILIZATION) {                         // alphazero_moves.cpp:146-171, State::attackReinforcementMove
        if ((uint32_t)li == g.mob_from) ga = true;
        else {
            const int from = (int)g.mob_from;
            const int af = (int)(land.get(from) & 63u);
            const int v = af - 1;
            int rf = v / 2;
            if (rf < r.min_unit_move) rf = r.min_unit_move < v ? r.min_unit_move : v;
            ia = from; va = (uint32_t)(af - rf); ib = li; vb = (uint32_t)(at + rf); wa = wb = true;
            ga = af - rf == 1;
        }
    } else if (ph == AZ_PH_FORTIFY) {                              // target already full: nothing moves, the turn ends
        et = true;
    } else {                                                       // SETUP / SETUP_NEUTRAL / REINFORCEMENT: one land gains armies
        int gain;
        if (ph == AZ_PH_SETUP) {                                   // State::setupReinforcementMove, state.cpp:1009-1030
            g.reinf = (g.reinf - 2) & 0xff; gain = 2; g.phase = AZ_PH_SETUP_NEUTRAL;
        } else if (ph == AZ_PH_SETUP_NEUTRAL) {                    // setupReinforcementNeutralMove + nextPlayerSetupTurn
            gain = 1; ob = AZ_NEUTRAL;
            g.phase = AZ_PH_SETUP; g.round = (g.round + 1) & 0xffff; g.cur ^= 1u;
            if (g.reinf == 0) { g.phase = AZ_PH_REINFORCEMENT; g.reinf = (uint32_t)az_reinforcement_value(g.own(g.cur)); }
        } else {                                                   // alphazero_moves.cpp:104-121, playCards, State::reinforcementMove
            uint32_t cards = cur ? g.cards1 : g.cards0;
            if (cards >= 3) {
                cards -= 3;
                if (cur) g.cards1 = cards; else g.cards0 = cards;
                g.card_sets = (g.card_sets + 1) & 0xff;
                const int cs = (int)g.card_sets;
                g.reinf = (g.reinf + (uint32_t)(cs <= 5 ? 2 + 2 * cs : 15 + (cs - 6) * 5)) & 0xff;
            }
            int rf = (int)g.reinf / 2;
            if (rf < r.min_unit_move) rf = r.min_unit_move < (int)g.reinf ? r.min_unit_move : (int)g.reinf;
            const int space = AZ_ARMY_MAX - at;
            if (space < rf) rf = space;
            g.reinf = (g.reinf - (uint32_t)rf) & 0xff;
            gain = rf;
            ga = g.reinf == 0;
        }
        vb = (uint32_t)(at + gain); wb = true;
    }
    // land A (attack / mobilisation source) and land B keep their owners in every case but a capture, so only the army
    // masks are updated here; a capture moves land B's bit between the ownership masks
    if (wa) az_set_army(g, land, ia, va, cur);
    if (wb) az_set_army(g, land, ib, vb, ob);
    if (captured) {
        const uint64_t m = 1ull << ib;
        if (cur == 0) { g.own0 |= m; g.own1 &= ~m; } else { g.own1 |= m; g.own0 &= ~m; }
    }
    if (ga) { g.phase = AZ_PH_ATTACK; g.mob_from = AZ_NONE; g.mob_to = AZ_NONE; g.reinf = 0; }
    // gotoAttack (state.cpp:20-40) and the end of attackMove (:909-912) share one test, run once for every lane:
    // the ATTACK phase is only entered / kept while the mover still has a land that can attack
    const uint64_t aa = az_attack_army(g, T, cur);
    if (g.phase == AZ_PH_ATTACK && aa == 0) g.phase = AZ_PH_FORTIFY;
    if (et) az_end_turn(g);
    return 0;
}

// the deferred half of a FORTIFY move (alphazero_moves.cpp:172-231 + State::fortifyMove state.cpp:949-974 + nextPlayerGameTurn)
template <class LandT, class ScratchT, class TabT>
__device__ __forceinline__ void az_fortify_finish(AzGame& g, LandT& land, ScratchT& scratch, const TabT& T, int li)
{
    const uint32_t cur = g.cur;
    const int at = (int)(land.get(li) & 63u);
    int from, amount;
    az_fortify_source(g, land, scratch, T, li, from, amount);
    if (from >= 0) {
        const int space = AZ_ARMY_MAX - at;
        const int mv = space < amount ? space : amount;
        const int af = (int)(land.get(from) & 63u);
        az_set_army(g, land, from, (uint32_t)(af - mv), cur);
        az_set_army(g, land, li, (uint32_t)(at + mv), cur);
    }
    az_end_turn(g);
}

// State::newGame, state/state.cpp:137-167 (Utility::randomMask, land/land.cpp:100-112):
// 42 draws, draw i selects the k-th remaining land, k = mulhi(word, 42 - i); lands go to
// player 0, player 1, neutral, player 0, ... with one army each.
template <class LandT>
__device__ __forceinline__ void az_new_game(AzGame& g, LandT& land, uint64_t seed, uint32_t game, uint32_t ply)
{
    g.own0 = g.own1 = g.gt1 = g.full = 0;
    g.round = 1; g.cur = 0; g.card_sets = 0; g.reinf = 52; g.phase = AZ_PH_SETUP;
    g.mob_from = AZ_NONE; g.mob_to = AZ_NONE; g.allow_draw = 0; g.attacks = 0; g.cards0 = g.cards1 = 0;
    uint64_t avail = AZ_ALL_LANDS;
    az_u32x4 blk;
    for (uint32_t i = 0; i < 42; ++i) {
        if ((i & 3u) == 0) blk = az_rng_block(seed, game, ply, AZ_STREAM_DEAL, i >> 2);
        uint32_t k = az_mulhi32(az_u32x4_word(blk, (int)(i & 3u)), 42u - i);
        int l = az_nth_set_bit(avail, k);
        avail &= ~(1ull << l);
        uint32_t owner = i % 3u;                              // P0, P1, neutral, P0, ...
        land.set(l, 1u | (owner << 6));
        if (owner == 0) g.own0 |= 1ull << l;
        if (owner == 1) g.own1 |= 1ull << l;
    }
}
