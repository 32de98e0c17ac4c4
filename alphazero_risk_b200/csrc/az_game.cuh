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

// n-th (0-based) set bit of a 42/43-bit mask
__device__ __forceinline__ int az_nth_set_bit(uint64_t m, uint32_t n)
{
    uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    uint32_t pl = (uint32_t)__popc(lo);
    return n < pl ? (int)__fns(lo, 0, (int)n + 1) : 32 + (int)__fns(hi, 0, (int)(n - pl) + 1);
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
__device__ __forceinline__ uint64_t az_attack_army(const AzGame& g, const AzTables& T, uint32_t p)
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
// order).  The recursion is replaced by a parent-pointer walk: after returning to a node, rescanning
// its list from the start finds the same next child because everything before it is already seen.
template <class LandT, class ScratchT>
__device__ __forceinline__ void az_fortify_source(const AzGame& g, const LandT& land, ScratchT& parent, const AzTables& T,
                                                  int li, int& from_out, int& amount_out)
{
    const uint64_t owned = g.own(g.cur);
    uint64_t comp = 1ull << li;
    for (;;) { uint64_t n = (comp | az_nbr_union(T, comp)) & owned; if (n == comp) break; comp = n; }
    const int seed = __ffsll((long long)comp) - 1;
    int best_i = 0, from_i = -1, best_b = 0, from_b = -1;
    uint64_t seen = 0;
    int v = seed;
    for (;;) {
        // visit v (pre-order position)
        seen |= 1ull << v;
        if (v != li) {
            int val = (int)(land.get(v) & 63u) - 1;
            uint64_t nb = T.nbr[v];
            if ((nb & owned) == nb) { if (val > best_i) { best_i = val; from_i = v; } }
            else                    { if (val > best_b) { best_b = val; from_b = v; } }
        }
        if (seen == comp) break;
        // next unvisited owned land in DFS order
        for (;;) {
            uint64_t cand = T.nbr[v] & owned & ~seen;
            if (cand) {
                uint64_t lst = T.list6[v];
                int u = (int)(lst & 63u);
                while (!((cand >> u) & 1ull)) { lst >>= 6; u = (int)(lst & 63u); }
                parent.set(u, (uint32_t)v);
                v = u;
                break;
            }
            v = (int)parent.get(v);   // backtrack (never past the seed: seen != comp guarantees a candidate upstream)
        }
    }
    if (from_i >= 0) { from_b = from_i; best_b = best_i; }
    from_out = from_b; amount_out = best_b;
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
