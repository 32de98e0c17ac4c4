// az_script.cuh — the reference's scripted opponent on the device (one thread per game).
//
// Behaviour: ScriptPlayer::takeTurn, /root/reference/src/risk_game/player/script/script_player.cpp:162-227 (target choice
// :19-70, attackLand :71-136, fortify :138-160 over GameHelper::PlayerMovement, player/game_helper.cpp:51-109) on the State
// primitives of state/state.cpp (reinforcementMove :976-998, attackMove :769-918, attackReinforcementMove :920-947,
// fortifyMove :949-974, setupReinforcementMove / setupReinforcementNeutralMove :1009-1053, playCards :1091-1117,
// nextPlayerGameTurn :748-766).  One call plays the player's WHOLE turn.  Dice are consumed sequentially from the
// AZ_STREAM_OPP stream of (game, ply), Utility::randomMask's rInt() from AZ_STREAM_OPP_INT (include/az_philox.h).
#pragma once

#include "az_game.cuh"

// land/land_set.cpp:12-24 — continent masks (NA, SA, EU, AF, AS, AU) and member lands in declaration order, 6 bits each,
// 63-terminated (the first attackable land of the chosen continent is the target)
__device__ __constant__ uint64_t AZ_CONT_MASK[6] = { 0x1ffull, 0x1e00ull, 0xfe000ull, 0x3f00000ull, 0x3ffc000000ull, 0x3c000000000ull };
__device__ __constant__ uint8_t AZ_CONT_LANDS[6][13] = {
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 63, 63, 63, 63 },
    { 9, 10, 11, 12, 63, 63, 63, 63, 63, 63, 63, 63, 63 },
    { 13, 14, 15, 16, 17, 19, 18, 63, 63, 63, 63, 63, 63 },
    { 20, 21, 22, 24, 25, 23, 63, 63, 63, 63, 63, 63, 63 },
    { 26, 33, 35, 36, 27, 28, 29, 30, 31, 32, 34, 37, 63 },
    { 38, 39, 40, 41, 63, 63, 63, 63, 63, 63, 63, 63, 63 },
};

// ScriptPlayer members that survive between calls (attackingLandSet, landAttackTo, landAttackFrom, attackFromArmy), one byte each;
// 0xff = never set.  The reference keeps stale values when a search finds no candidate; so does this.
struct AzScript {
    uint32_t set, to, from, from_army;
    __device__ __forceinline__ void unpack(uint32_t w) { set = w & 0xff; to = (w >> 8) & 0xff; from = (w >> 16) & 0xff; from_army = w >> 24; }
    __device__ __forceinline__ uint32_t pack() const { return (set & 0xff) | ((to & 0xff) << 8) | ((from & 0xff) << 16) | (from_army << 24); }
};
#define AZ_SCRIPT_INIT 0x00ffffffu

// Player::addTrainingSample (player/base/player.cpp:9-17) inside a scripted / random turn: every call site of script_player.cpp and
// random_player.cpp hands over (state before the move, move).  The sink is this game's staging slice in HBM: packed primary states
// (14 words each, the layout of AZ_PRIMARY_WORDS) and one move byte per sample; `n` counts every call, also those past `cap`.
struct AzTurnSink {
    uint32_t* st; uint8_t* mv; uint32_t cap, n;
};
template <class LandT>
__device__ __noinline__ void az_turn_emit(AzTurnSink* k, const AzGame& g, const LandT& land, int move)
{
    if (k == nullptr) return;
    if (k->n < k->cap) {
        uint32_t* dst = k->st + (size_t)k->n * AZ_PRIMARY_WORDS;
        for (int w = 0; w < 10; ++w)
            dst[w] = land.get(4 * w) | (land.get(4 * w + 1) << 8) | (land.get(4 * w + 2) << 16) | (land.get(4 * w + 3) << 24);
        dst[10] = land.get(40) | (land.get(41) << 8) | (g.cards0 << 16) | (g.cards1 << 24);
        dst[11] = az_pack_w11(g); dst[12] = az_pack_w12(g); dst[13] = az_pack_w13(g);
        k->mv[k->n] = (uint8_t)move;
    }
    k->n++;
}

template <class LandT>
struct AzScriptCtx {
    AzGame& g; LandT& land; const AzTables& T; const AzRulesDev& r;
    AzScript sp;
    AzDicePhilox dice;
    uint64_t owned_attack_mask, attack_mask;      // ScriptPlayer::ownedAttackLandBitMask / attackLandBitMask
    AzTurnSink* sink = nullptr;
    __device__ AzScriptCtx(AzGame& g_, LandT& l_, const AzTables& T_, const AzRulesDev& r_) : g(g_), land(l_), T(T_), r(r_) {}
};

// updateAttackLandSetPriority / updateAttackLandSet / updateAttackLandTo / updateAttackLandFrom, script_player.cpp:19-70.
// GameHelper::sortLandSet (game_helper.cpp:19-39) is a strict total order, so "first set of the sorted list with an attackable
// land" is the minimum of that order over the sets that have one.
template <class LandT>
__device__ __forceinline__ void az_script_pick_target(AzScriptCtx<LandT>& c)
{
    const uint64_t owned = c.g.own(c.g.cur);
    int best = -1, best_no = 0, best_na = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const uint64_t m = AZ_CONT_MASK[k] & ~owned;
        const int no = __popcll(m), na = __popcll(m & c.attack_mask);
        if (na > 0) {
            bool before = best < 0 || no < best_no || (no == best_no && (na > best_na || (na == best_na && AZ_CONT_MASK[k] > AZ_CONT_MASK[best])));
            if (before) { best = k; best_no = no; best_na = na; }
        }
    }
    if (best >= 0) c.sp.set = (uint32_t)best;
    if (c.sp.set < 6) {
        for (int i = 0; i < 13; ++i) {
            const int l = AZ_CONT_LANDS[c.sp.set][i];
            if (l == 63) break;
            if ((c.attack_mask >> l) & 1ull) { c.sp.to = (uint32_t)l; break; }
        }
    }
    c.sp.from_army = 0;
    if (c.sp.to < AZ_LANDS) {
        uint64_t lst = c.T.list6[c.sp.to];
        for (int k = 0; k < 6; ++k) {
            const int n = (int)(lst & 63u); lst >>= 6;
            if (n == 63) break;
            if ((c.owned_attack_mask >> n) & 1ull) {
                const uint32_t a = c.land.get(n) & 63u;
                if (a > c.sp.from_army) { c.sp.from_army = a; c.sp.from = (uint32_t)n; }
            }
        }
    }
}

// State::attackMove, state/state.cpp:769-918; true when the land was captured
template <class LandT>
__device__ __forceinline__ bool az_script_attack(AzScriptCtx<LandT>& c, int from, int to)
{
    AzGame& g = c.g;
    const uint32_t cur = g.cur;
    g.attacks = (g.attacks + 1) & 0xff;
    const uint32_t tob = c.land.get(to);
    int a = (int)(c.land.get(from) & 63u), d = (int)(tob & 63u), units = 1;
    const uint32_t defender = tob >> 6;
    if (d > 0) {
        const int na = a >= 4 ? 3 : (a == 3 ? 2 : 1);
        const int nd = d >= 2 ? 2 : 1;
        units = na;
        int a0 = c.dice.next(), a1 = 0, a2 = 0;
        if (na > 1) a1 = c.dice.next();
        if (na > 2) a2 = c.dice.next();
        int d0 = c.dice.next(), d1 = 0;
        if (nd > 1) d1 = c.dice.next();
        const int hi = max(a0, max(a1, a2));
        int lo = min(a0, max(a1, a2)); lo = max(lo, min(a1, a2));
        const int dh = max(d0, d1), dl = min(d0, d1);
        if (hi > dh) d--; else { a--; units--; }
        if (na >= 2 && nd == 2) { if (lo > dl) d--; else { a--; units--; } }
    }
    bool captured = false;
    if (d == 0) {
        a -= units;
        if (a > 1) { g.phase = AZ_PH_MOBILIZATION; g.mob_from = (uint32_t)from; g.mob_to = (uint32_t)to; }
        g.allow_draw = 1;
        az_set_land(g, c.land, from, (uint32_t)a, cur);
        az_set_land(g, c.land, to, (uint32_t)units, cur);
        captured = true;
    } else {
        az_set_land(g, c.land, from, (uint32_t)a, cur);
        az_set_land(g, c.land, to, (uint32_t)d, defender);
    }
    if (g.phase == AZ_PH_ATTACK && az_attack_army(g, c.T, cur) == 0) g.phase = AZ_PH_FORTIFY;
    return captured;
}

// ScriptPlayer::attackLand, script_player.cpp:71-136
template <class LandT>
__device__ __forceinline__ void az_script_attack_land(AzScriptCtx<LandT>& c)
{
    AzGame& g = c.g;
    const uint32_t me = g.cur;
    while (g.reinf > 0) {
        const uint64_t owned = g.own(me), not_full = owned & ~g.full;
        int to = (int)c.sp.from;
        if (((not_full >> c.sp.from) & 1ull) == 0) {
            uint64_t nb = c.T.nbr[c.sp.to] & not_full;
            if (nb == 0) {
                // enemy.attackLands | neutralAttackLands restricted to my lands == N(every land that is not mine)
                nb = not_full & az_nbr_union(c.T, AZ_ALL_LANDS & ~owned);
                if (nb == 0) nb = not_full;
            }
            to = __ffsll((long long)nb) - 1;
        }
        int army = (int)(c.land.get(to) & 63u);
        int amount = AZ_ARMY_MAX - army;
        if ((int)g.reinf < amount) amount = (int)g.reinf;
        // State::reinforcementMove in MIN_UNIT_MOVE steps (state.cpp:976-998): only the sum and the last step's gotoAttack are
        // observable — unless the samples are recorded: then every step's state is one (script_player.cpp:99-109)
        if (c.sink) {
            for (int rem = amount; rem > 0; ) {
                const int step = rem < c.r.min_unit_move ? rem : c.r.min_unit_move;
                az_turn_emit(c.sink, g, c.land, to);
                g.reinf = (g.reinf - (uint32_t)step) & 0xff;
                army += step; rem -= step;
                az_set_land(g, c.land, to, (uint32_t)army, me);
                if (g.reinf == 0) az_goto_attack(g, c.T);
            }
        } else {
            g.reinf = (g.reinf - (uint32_t)amount) & 0xff;
            az_set_land(g, c.land, to, (uint32_t)(army + amount), me);
            if (amount > 0 && g.reinf == 0) az_goto_attack(g, c.T);
        }
        if (amount == 0) break;                    // every owned land is full: the reference would spin here
    }
    c.sp.from_army = c.land.get(c.sp.from) & 63u;
    while (c.sp.from_army > 1) {
        az_turn_emit(c.sink, g, c.land, (int)c.sp.to);                            // script_player.cpp:115
        const bool captured = az_script_attack(c, (int)c.sp.from, (int)c.sp.to);
        c.sp.from_army = c.land.get(c.sp.from) & 63u;
        if (captured && c.sp.from_army > 1) {
            // State::attackReinforcementMove in MIN_UNIT_MOVE steps (state.cpp:920-947) until one army is left behind
            const int left = (int)c.sp.from_army - 1;
            const int at = (int)(c.land.get(g.mob_to) & 63u);
            const uint32_t mf = g.mob_from, mt = g.mob_to;
            if (c.sink) {                                                         // one sample per step, script_player.cpp:122-131
                int af = (int)c.sp.from_army, t = at;
                for (int rem = left; rem > 0; ) {
                    const int step = rem < c.r.min_unit_move ? rem : c.r.min_unit_move;
                    az_turn_emit(c.sink, g, c.land, (int)c.sp.to);
                    af -= step; t += step; rem -= step;
                    az_set_land(g, c.land, (int)mf, (uint32_t)af, me);
                    az_set_land(g, c.land, (int)mt, (uint32_t)t, me);
                }
            } else {
                az_set_land(g, c.land, (int)mf, 1u, me);
                az_set_land(g, c.land, (int)mt, (uint32_t)(at + left), me);
            }
            az_goto_attack(g, c.T);
            break;
        }
    }
}

// ScriptPlayer::fortify, script_player.cpp:138-160: GameHelper::PlayerMovement lists the owned components in ascending seed order,
// each in DFS pre-order (children in neighbour-list order); per component the interior land with the largest army (first strict
// maximum) is the source, the border land with the most foreign neighbours (first strict maximum) the target; the component with
// the largest source army is used (std::sort on <= 16 elements = libstdc++'s stable insertion sort: the first maximum).
template <class LandT, class ScratchT>
__device__ __forceinline__ void az_script_fortify(AzGame& g, LandT& land, ScratchT& parent, const AzTables& T, AzTurnSink* sink = nullptr)
{
    const uint32_t me = g.cur;
    const uint64_t owned = g.own(me);
    if ((owned & g.gt1) == 0) return;
    uint64_t seen = 0;
    int best_from = -1, best_to = -1, best_amount = -1;
    while (owned & ~seen) {
        const int seed = __ffsll((long long)(owned & ~seen)) - 1;
        int from = -1, from_amount = 0, to = -1, to_nbrs = 0;
        int v = seed;
        for (;;) {
            seen |= 1ull << v;
            const uint64_t foreign = ~owned & T.nbr[v];
            if (foreign == 0) { const int a = (int)(land.get(v) & 63u); if (a > from_amount) { from = v; from_amount = a; } }
            else { const int cnt = __popcll(foreign); if (cnt > to_nbrs) { to_nbrs = cnt; to = v; } }
            // next unvisited owned land in DFS order, backtracking through parent pointers; done when back at the seed with nothing left
            bool done = false;
            for (;;) {
                const uint64_t nxt = T.nbr[v] & owned & ~seen;
                if (nxt) {
                    uint64_t lst = T.list6[v];
                    int u = (int)(lst & 63u);
                    while (!((nxt >> u) & 1ull)) { lst >>= 6; u = (int)(lst & 63u); }
                    parent.set(u, (uint32_t)v);
                    v = u;
                    break;
                }
                if (v == seed) { done = true; break; }
                v = (int)parent.get(v);
            }
            if (done) break;
        }
        if (from_amount > best_amount) { best_amount = from_amount; best_from = from; best_to = to; }
    }
    if (best_amount > 0 && best_to >= 0) {          // State::fortifyMove, state.cpp:949-974
        const int af = (int)(land.get(best_from) & 63u), at = (int)(land.get(best_to) & 63u);
        int amount = af - 1;
        const int space = AZ_ARMY_MAX - at;
        if (space < amount) amount = space;
        az_turn_emit(sink, g, land, best_to);                                     // script_player.cpp:151
        az_set_land(g, land, best_from, (uint32_t)(af - amount), me);
        az_set_land(g, land, best_to, (uint32_t)(at + amount), me);
    } else az_turn_emit(sink, g, land, AZ_SKIP);                                  // :157
}

// ScriptPlayer::takeTurn.  `sp_word` = the packed AzScript of this (game slot, side); returns 0, or AZ_STATUS_ILLEGAL when the
// game is not at the start of a turn (the script never resumes one).
template <class LandT, class ScratchT>
__device__ __forceinline__ int az_script_turn(AzGame& g, LandT& land, ScratchT& scratch, const AzTables& T, const AzRulesDev& r,
                                              uint32_t& sp_word, uint64_t seed, uint32_t game, uint32_t ply, AzTurnSink* sink = nullptr)
{
    AzScriptCtx<LandT> c(g, land, T, r);
    c.sink = sink;
    c.sp.unpack(sp_word);
    c.dice.init(seed, game, ply, AZ_STREAM_OPP);
    const uint32_t me = g.cur;
    const uint64_t owned = g.own(me), enemy = g.own(me ^ 1u);
    c.owned_attack_mask = owned;
    c.attack_mask = az_nbr_union(T, owned) & ~owned;
    if (g.phase == AZ_PH_SETUP) {
        az_script_pick_target(c);
        az_turn_emit(sink, g, land, (int)c.sp.from);                  // script_player.cpp:176
        g.reinf = (g.reinf - 2) & 0xff;                               // setupReinforcementMove, state.cpp:1009-1030
        az_set_land(g, land, (int)c.sp.from, (land.get(c.sp.from) & 63u) + 2, me);
        const uint64_t neutral = AZ_ALL_LANDS & ~owned & ~enemy;
        const uint64_t enemy_attack = az_nbr_union(T, enemy) & ~enemy;
        uint64_t near_enemy = neutral & enemy_attack & ~c.attack_mask;
        if (near_enemy == 0) near_enemy = neutral & enemy_attack;
        const uint64_t pool = near_enemy ? near_enemy : neutral;
        const uint32_t k = az_rng_opp_int(seed, game, ply, 0) % (uint32_t)__popcll(pool);      // Utility::randomMask, land.cpp:100-112
        const int l = az_nth_set_bit(pool, k);
        g.phase = AZ_PH_SETUP_NEUTRAL;                                // (the state the sample sees)
        az_turn_emit(sink, g, land, l);                               // :198
        az_set_land(g, land, l, (land.get(l) & 63u) + 1, AZ_NEUTRAL);   // setupReinforcementNeutralMove + nextPlayerSetupTurn
        g.phase = AZ_PH_SETUP; g.round = (g.round + 1) & 0xffff; g.cur ^= 1u;
        if (g.reinf == 0) { g.phase = AZ_PH_REINFORCEMENT; g.reinf = (uint32_t)az_reinforcement_value(g.own(g.cur)); }
        sp_word = c.sp.pack();
        return 0;
    }
    if (g.phase != AZ_PH_REINFORCEMENT) return AZ_STATUS_ILLEGAL;
    uint32_t cards = me ? g.cards1 : g.cards0;                        // GameHelper::playCards + State::playCards, state.cpp:1091-1117
    if (cards >= 3) {
        cards -= 3;
        if (me) g.cards1 = cards; else g.cards0 = cards;
        g.card_sets = (g.card_sets + 1) & 0xff;
        const int cs = (int)g.card_sets;
        g.reinf = (g.reinf + (uint32_t)(cs <= 5 ? 2 + 2 * cs : 15 + (cs - 6) * 5)) & 0xff;
    }
    int guard = 0;
    while ((c.attack_mask != 0 || g.reinf > 0) && ++guard < 4096) {
        az_script_pick_target(c);
        az_script_attack_land(c);
        const uint64_t o = g.own(me);
        c.owned_attack_mask = o & g.gt1;
        c.attack_mask = az_attack_army(g, T, me);
    }
    az_script_fortify(g, land, scratch, T, sink);
    az_end_turn(g);
    sp_word = c.sp.pack();
    return 0;
}

// ---------------------------------------------------------------- RandomPlayer
// RandomPlayer::takeTurn, /root/reference/src/risk_game/player/random/random_player.cpp:22-111: uniformly random moves on the State
// primitives (one army per reinforcement move, random attack-from land, coin-flip mobilisation, random fortify amount) until the
// turn passes.  pickRandomMove = one rInt() (k-th lowest set bit, k = rInt() % count), the coin = one rFloat(); both advance the
// opponent word sequence (az_rng_opp_word); dice from AZ_STREAM_OPP.
template <class LandT>
struct AzRandomCtx {
    uint64_t seed; uint32_t game, ply, int_j;
    __device__ __forceinline__ int pick(uint64_t mask)
    {
        const uint32_t k = az_rng_opp_int(seed, game, ply, int_j++) % (uint32_t)__popcll(mask);
        return az_nth_set_bit(mask, k);
    }
};

template <class LandT, class ScratchT>
__device__ __forceinline__ int az_random_turn(AzGame& g, LandT& land, ScratchT& scratch, const AzTables& T, const AzRulesDev& r,
                                              uint64_t seed, uint32_t game, uint32_t ply, AzTurnSink* sink = nullptr)
{
    AzScriptCtx<LandT> c(g, land, T, r);                 // only for az_script_attack (State::attackMove with sequential dice)
    c.dice.init(seed, game, ply, AZ_STREAM_OPP);
    AzRandomCtx<LandT> rc; rc.seed = seed; rc.game = game; rc.ply = ply; rc.int_j = 0;
    const uint32_t me = g.cur;
    int guard = 0;
    while (az_game_status(g, r) == AZ_STATUS_RUNNING && g.cur == me && ++guard < 8192) {
        const uint64_t owned = g.own(me);
        switch (g.phase) {
        case AZ_PH_SETUP: {
            const int li = rc.pick(owned);
            az_turn_emit(sink, g, land, li);                                      // random_player.cpp:29
            g.reinf = (g.reinf - 2) & 0xff;
            az_set_land(g, land, li, (land.get(li) & 63u) + 2, me);
            g.phase = AZ_PH_SETUP_NEUTRAL;
            break;
        }
        case AZ_PH_SETUP_NEUTRAL: {
            const int li = rc.pick(AZ_ALL_LANDS & ~g.own0 & ~g.own1);
            az_turn_emit(sink, g, land, li);                                      // :35
            az_set_land(g, land, li, (land.get(li) & 63u) + 1, AZ_NEUTRAL);
            g.phase = AZ_PH_SETUP; g.round = (g.round + 1) & 0xffff; g.cur ^= 1u;
            if (g.reinf == 0) { g.phase = AZ_PH_REINFORCEMENT; g.reinf = (uint32_t)az_reinforcement_value(g.own(g.cur)); }
            break;
        }
        case AZ_PH_REINFORCEMENT: {
            uint32_t cards = me ? g.cards1 : g.cards0;
            if (cards >= 3) {
                cards -= 3;
                if (me) g.cards1 = cards; else g.cards0 = cards;
                g.card_sets = (g.card_sets + 1) & 0xff;
                const int cs = (int)g.card_sets;
                g.reinf = (g.reinf + (uint32_t)(cs <= 5 ? 2 + 2 * cs : 15 + (cs - 6) * 5)) & 0xff;
            }
            const int li = rc.pick(owned & ~g.full);
            az_turn_emit(sink, g, land, li);                                      // :43 (after playCards)
            g.reinf = (g.reinf - 1) & 0xff;
            az_set_land(g, land, li, (land.get(li) & 63u) + 1, me);
            if (g.reinf == 0) az_goto_attack(g, T);
            break;
        }
        case AZ_PH_ATTACK: {
            const int li = rc.pick(az_attack_army(g, T, me) | AZ_SKIP_MASK);
            az_turn_emit(sink, g, land, li);                                      // :49
            if (li == AZ_SKIP) g.phase = AZ_PH_FORTIFY;
            else {
                const int from = rc.pick(T.nbr[li] & owned & g.gt1);
                az_script_attack(c, from, li);
            }
            break;
        }
        case AZ_PH_MOBILIZATION: {
            const float f = az_rng_unit_float(az_rng_opp_word(seed, game, ply, rc.int_j++));
            if (f > 0.5f) {
                const int af = (int)(land.get(g.mob_from) & 63u), at = (int)(land.get(g.mob_to) & 63u);
                int amount = af - 1;
                if (r.min_unit_move < amount) amount = r.min_unit_move;
                const uint32_t mf = g.mob_from, mt = g.mob_to;
                az_turn_emit(sink, g, land, (int)mt);                             // :68
                az_set_land(g, land, (int)mf, (uint32_t)(af - amount), me);
                az_set_land(g, land, (int)mt, (uint32_t)(at + amount), me);
                if (af - amount == 1) az_goto_attack(g, T);
            } else { az_turn_emit(sink, g, land, (int)g.mob_from); az_goto_attack(g, T); }     // :73
            break;
        }
        default: {
            const int to = rc.pick((owned & ~g.full) | AZ_SKIP_MASK);
            az_turn_emit(sink, g, land, to);                                      // :82
            if (to != AZ_SKIP) {
                uint64_t comp = 1ull << to;
                for (;;) { const uint64_t n = (comp | az_nbr_union(T, comp)) & owned; if (n == comp) break; comp = n; }
                const uint64_t pool = comp & ~(1ull << to) & g.gt1;
                if (pool) {
                    const int from = rc.pick(pool);
                    const int af = (int)(land.get(from) & 63u), at = (int)(land.get(to) & 63u);
                    int amount = af - 1;
                    const int space = AZ_ARMY_MAX - at;
                    if (space < amount) amount = space;
                    const int moved = (int)(az_rng_opp_int(seed, game, ply, rc.int_j++) % (uint32_t)amount);
                    az_set_land(g, land, from, (uint32_t)(af - moved), me);
                    az_set_land(g, land, to, (uint32_t)(at + moved), me);
                }
            }
            az_end_turn(g);
            break;
        }
        }
    }
    (void)scratch;
    return 0;
}
