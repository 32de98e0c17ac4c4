/*
 * TEST INFRASTRUCTURE — risk6_oracle.c: plain-C statement of the SIX-PLAYER extension of the Risk rules (BASELINE.json configs[3]),
 * the only checker of the CUDA six-player environment (az_env6_*, alphazero_risk_b200/csrc/az_env6.cu).
 *
 * PARITY UNPINNED: the reference has no six-player game (PLAYER_COUNT = 2 is a compile-time constant,
 * /root/reference/src/risk_game/state/state.h:13; two owner bits with 2 = neutral, state.h:24-41), so there is nothing to pin this
 * file against.  It states SIXPLAYER.md; everything that has a two-player counterpart follows that counterpart's arithmetic and
 * cites it (paths relative to /root/reference/src/risk_game).  Never linked into or called by the product.
 */
#include <string.h>

#include "az_philox.h"
#include "risk_oracle.h"
#include "risk6_oracle.h"

#define ALL_LANDS 0x3ffffffffffull
#define SKIP_MASK (1ull << R6_SKIP)

static int popc(uint64_t v) { return __builtin_popcountll(v); }
static int ctz(uint64_t v) { return __builtin_ctzll(v); }

static uint64_t nbr_union(uint64_t set)
{
    uint64_t u = 0;
    while (set) { int i = ctz(set); set &= set - 1; u |= RO_NBR_MASK[i]; }
    return u;
}

typedef struct derived6 { uint64_t owned[R6_PLAYERS], gt1, full; } derived6;

static void derive(const r6_state* s, derived6* d)
{
    memset(d, 0, sizeof *d);
    for (int i = 0; i < RO_LANDS; ++i) {
        d->owned[s->owner[i]] |= 1ull << i;
        if (s->army[i] > 1) d->gt1 |= 1ull << i;
        if (s->army[i] == RO_ARMY_MAX) d->full |= 1ull << i;
    }
}

static uint64_t attack_army(const derived6* d, int p) { return nbr_union(d->owned[p] & d->gt1) & ~d->owned[p]; }

/* State::gameStatus (state/state.cpp:518-565) for six players: last player standing, the 30-land yield rule, the round limit */
int r6_game_status(const r6_state* s, const ro_rules* r)
{
    derived6 d; derive(s, &d);
    int alive = 0, last = -1, best = -1, best_n = -1, tie = 0;
    for (int p = 0; p < R6_PLAYERS; ++p) {
        int n = popc(d.owned[p]);
        if (n > 0) { alive++; last = p; }
        if (n > best_n) { best_n = n; best = p; tie = 0; } else if (n == best_n) tie = 1;
    }
    if (alive == 1) return last;
    if (r->allow_yield && best_n >= 30) return best;            /* at most one player can hold 30 of 42 lands */
    if ((int)s->round > r->max_game_rounds) return tie ? R6_DRAW : best;
    return R6_NOT_ENDED;
}

/* UtilityNN::getValidMoves (player/alpha_zero/alphazero_moves.cpp:3-70): "the enemy" becomes "every other player" */
uint64_t r6_valid_moves(const r6_state* s, const ro_rules* r)
{
    derived6 d; derive(s, &d);
    const uint64_t oc = d.owned[s->cur];
    const uint64_t border = nbr_union(ALL_LANDS & ~oc);          /* lands adjacent to a land the mover does not own */
    switch (s->phase) {
    case R6_SETUP:
    case R6_REINFORCEMENT: {
        uint64_t o = oc & ~d.full;
        if (o == 0) return SKIP_MASK;
        if (r->limit_reinforcement && (o & border)) return o & border;
        return o;
    }
    case R6_ATTACK: {
        uint64_t aa = attack_army(&d, s->cur);
        if (r->limit_attack) return aa ? aa : SKIP_MASK;
        return aa | SKIP_MASK;
    }
    case R6_MOBILIZATION: return (1ull << s->mob_from) | (1ull << s->mob_to);
    case R6_FORTIFY:
        if (r->limit_reinforcement) return (oc & border) | SKIP_MASK;
        return oc | SKIP_MASK;
    }
    return 0;
}

static void goto_attack(r6_state* s)                              /* State::gotoAttack, state/state.cpp:20-40 */
{
    s->phase = R6_ATTACK; s->mob_from = R6_NONE; s->mob_to = R6_NONE; s->reinf = 0;
    derived6 d; derive(s, &d);
    if (attack_army(&d, s->cur) == 0) s->phase = R6_FORTIFY;
}

/* State::nextPlayerGameTurn (state/state.cpp:748-766): the next player in seat order that still owns a land; the round counter
   advances whenever the order wraps past seat 5 */
static void end_turn(r6_state* s)
{
    if (s->allow_draw) { s->cards[s->cur] = (uint8_t)(s->cards[s->cur] + 1); s->allow_draw = 0; }
    derived6 d; derive(s, &d);
    int next = s->cur;
    for (int k = 0; k < R6_PLAYERS; ++k) {
        next = (next + 1) % R6_PLAYERS;
        if (next == 0) s->round++;
        if (d.owned[next]) break;
    }
    s->cur = (uint8_t)next;
    s->attacks = 0;
    s->phase = R6_REINFORCEMENT;
    s->reinf = (uint8_t)ro_reinforcement_value(d.owned[next]);
}

static void dfs(int l, uint64_t owned, uint64_t* seen, int* order, int* n)   /* GameHelper::LandSetMovement::add, game_helper.cpp:51-82 */
{
    if (!((1ull << l) & owned & ~*seen)) return;
    *seen |= 1ull << l;
    order[(*n)++] = l;
    for (int k = 0; k < 6 && RO_NBR_LIST[l][k] >= 0; ++k) dfs(RO_NBR_LIST[l][k], owned, seen, order, n);
}

static void sort_desc(int* d, int n)
{
    for (int i = 0; i < n; ++i)
        for (int k = i + 1; k < n; ++k)
            if (d[k] > d[i]) { int t = d[i]; d[i] = d[k]; d[k] = t; }
}

/* UtilityNN::makeMove (alphazero_moves.cpp:72-233) for six players; dice = the die stream of (seed, game, ply, AZ_STREAM_REAL) */
int r6_make_move(r6_state* s, int action, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply)
{
    if (r6_game_status(s, r) != R6_NOT_ENDED) return RO_ERR_GAME_OVER;
    if (action < 0 || action > R6_SKIP || !((r6_valid_moves(s, r) >> action) & 1)) return RO_ERR_ILLEGAL_ACTION;
    const int cur = s->cur;
    uint32_t die = 0;
    if (action == R6_SKIP) {
        switch (s->phase) {
        case R6_REINFORCEMENT: goto_attack(s); break;
        case R6_ATTACK: s->phase = R6_FORTIFY; break;
        case R6_FORTIFY: end_turn(s); break;
        default: return RO_ERR_ILLEGAL_ACTION;                   /* SETUP with every land full cannot happen: 7 lands x 32 >> 20 armies */
        }
        return RO_OK;
    }
    const int li = action;
    switch (s->phase) {
    case R6_SETUP: {
        /* one army on an own land, then the next seat; when the seat that comes up has nothing left to place the game proper
           starts with it (pools run out in seat order, so that seat is 0) */
        s->army[li] = (uint8_t)(s->army[li] + 1);
        s->pool[cur] = (uint8_t)(s->pool[cur] - 1);
        s->cur = (uint8_t)((cur + 1) % R6_PLAYERS);
        if (s->cur == 0) s->round++;
        if (s->pool[s->cur] == 0) {
            derived6 d; derive(s, &d);
            s->phase = R6_REINFORCEMENT; s->reinf = (uint8_t)ro_reinforcement_value(d.owned[s->cur]);
        }
        break;
    }
    case R6_REINFORCEMENT: {                                      /* alphazero_moves.cpp:104-121; State::playCards (simple mode) state.cpp:1091-1117 */
        if (s->cards[cur] >= 3) {
            s->cards[cur] = (uint8_t)(s->cards[cur] - 3);
            s->card_sets = (uint8_t)(s->card_sets + 1);
            int cs = s->card_sets;
            s->reinf = (uint8_t)(s->reinf + (cs <= 5 ? 2 + 2 * cs : 15 + (cs - 6) * 5));
        }
        int rf = s->reinf / 2;
        if (rf < r->min_unit_move) rf = r->min_unit_move < s->reinf ? r->min_unit_move : s->reinf;
        int space = RO_ARMY_MAX - s->army[li];
        if (space < rf) rf = space;
        s->reinf = (uint8_t)(s->reinf - rf);
        s->army[li] = (uint8_t)(s->army[li] + rf);
        if (s->reinf == 0) goto_attack(s);
        break;
    }
    case R6_ATTACK: {                                             /* alphazero_moves.cpp:122-145, State::attackMove state.cpp:769-918 */
        int best = 0, from = -1;
        for (int k = 0; k < 6 && RO_NBR_LIST[li][k] >= 0; ++k) {
            int n = RO_NBR_LIST[li][k];
            if (s->owner[n] == cur && s->army[n] > 1) { int v = s->army[n] - 1; if (v > best) { best = v; from = n; } }
        }
        if (from < 0) return RO_ERR_ILLEGAL_ACTION;
        s->attacks = (uint8_t)(s->attacks + 1);
        int a = s->army[from], d = s->army[li], defender = s->owner[li], units = 1;
        {
            int na = a >= 4 ? 3 : a == 3 ? 2 : 1, nd = d >= 2 ? 2 : 1;
            units = na;
            int ad[3] = { 0, 0, 0 }, dd[3] = { 0, 0, 0 };
            for (int i = 0; i < na; ++i) ad[i] = az_rng_die(seed, game, ply, AZ_STREAM_REAL, die++);   /* attacker dice first */
            for (int i = 0; i < nd; ++i) dd[i] = az_rng_die(seed, game, ply, AZ_STREAM_REAL, die++);
            sort_desc(ad, na); sort_desc(dd, nd);
            if (ad[0] > dd[0]) d--; else { a--; units--; }
            if (na >= 2 && nd == 2) { if (ad[1] > dd[1]) d--; else { a--; units--; } }
        }
        if (d == 0) {
            a -= units;
            if (a > 1) { s->phase = R6_MOBILIZATION; s->mob_from = (uint8_t)from; s->mob_to = (uint8_t)li; }
            s->allow_draw = 1;
            s->army[from] = (uint8_t)a; s->army[li] = (uint8_t)units; s->owner[li] = (uint8_t)cur;
            derived6 dd6; derive(s, &dd6);
            if (dd6.owned[defender] == 0) {                       /* elimination: the eliminator takes the cards */
                s->cards[cur] = (uint8_t)(s->cards[cur] + s->cards[defender]);
                s->cards[defender] = 0;
            }
        } else { s->army[from] = (uint8_t)a; s->army[li] = (uint8_t)d; }
        if (s->phase == R6_ATTACK) { derived6 d6; derive(s, &d6); if (attack_army(&d6, cur) == 0) s->phase = R6_FORTIFY; }
        break;
    }
    case R6_MOBILIZATION:                                         /* alphazero_moves.cpp:146-171 */
        if (li == s->mob_from) goto_attack(s);
        else {
            int from = s->mob_from, to = s->mob_to;
            int v = s->army[from] - 1, rf = v / 2;
            if (rf < r->min_unit_move) rf = r->min_unit_move < v ? r->min_unit_move : v;
            s->army[from] = (uint8_t)(s->army[from] - rf);
            s->army[to] = (uint8_t)(s->army[to] + rf);
            if (s->army[from] == 1) goto_attack(s);
        }
        break;
    case R6_FORTIFY: {                                            /* alphazero_moves.cpp:172-231, game_helper.cpp:90-109 */
        if (s->army[li] != RO_ARMY_MAX) {
            derived6 d6; derive(s, &d6);
            uint64_t owned = d6.owned[cur], grouped = 0;
            for (int sd = 0; sd < RO_LANDS; ++sd) {
                if (!((1ull << sd) & owned & ~grouped)) continue;
                uint64_t seen = 0; int order[RO_LANDS], n = 0;
                dfs(sd, owned, &seen, order, &n);
                grouped |= seen;
                if (!((seen >> li) & 1)) continue;
                int best_i = 0, from_i = -1, best_b = 0, from_b = -1;
                for (int j = 0; j < n; ++j) {
                    int l = order[j];
                    if (l == li) continue;
                    int v = s->army[l] - 1;
                    if ((RO_NBR_MASK[l] & owned) == RO_NBR_MASK[l]) { if (v > best_i) { best_i = v; from_i = l; } }
                    else { if (v > best_b) { best_b = v; from_b = l; } }
                }
                if (from_i >= 0) { from_b = from_i; best_b = best_i; }
                if (from_b >= 0) {
                    int space = RO_ARMY_MAX - s->army[li];
                    int mv = space < best_b ? space : best_b;
                    s->army[from_b] = (uint8_t)(s->army[from_b] - mv);
                    s->army[li] = (uint8_t)(s->army[li] + mv);
                }
                break;
            }
        }
        end_turn(s);
        break;
    }
    default: return RO_ERR_ILLEGAL_ACTION;
    }
    return RO_OK;
}

/* State::newGame (state/state.cpp:137-167): the 42 draws of the deal stream go to seats 0,1,2,3,4,5,0,... (7 lands each, one army
   per land); every seat then has 20 - 7 = 13 armies to place */
void r6_new_game(r6_state* s, uint64_t seed, uint32_t game, uint32_t ply)
{
    memset(s, 0, sizeof *s);
    uint64_t avail = ALL_LANDS;
    for (uint32_t i = 0; i < 42; ++i) {
        uint32_t k = az_rng_deal_draw(seed, game, ply, i);
        uint64_t m = avail; for (uint32_t t = 0; t < k; ++t) m &= m - 1;
        int l = ctz(m); avail &= ~(1ull << l);
        s->army[l] = 1; s->owner[l] = (uint8_t)(i % R6_PLAYERS);
    }
    for (int p = 0; p < R6_PLAYERS; ++p) s->pool[p] = 13;
    s->round = 1; s->cur = 0; s->phase = R6_SETUP; s->mob_from = R6_NONE; s->mob_to = R6_NONE;
}

/* the rollout's action rule, as in the two-player oracle: k-th set bit of the legal mask, k = mulhi(word 1 of the real-move block, popcount) */
int r6_random_action(const r6_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply)
{
    uint64_t m = r6_valid_moves(s, r);
    az_u32x4 b = az_rng_block(seed, game, ply, AZ_STREAM_REAL, 0);
    uint32_t k = az_mulhi32(b.y, (uint32_t)popc(m));
    for (uint32_t i = 0; i < k; ++i) m &= m - 1;
    return ctz(m);
}
