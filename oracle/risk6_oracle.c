/*
 * TEST INFRASTRUCTURE — risk6_oracle.c: plain-C statement of the SIX-PLAYER extension of the Risk rules (BASELINE.json configs[3]),
 * the only checker of the CUDA six-player environment (az_env6_*, alphazero_risk_b200/csrc/az_env6.cu).
 *
 * PARITY UNPINNED: the reference has no six-player game (PLAYER_COUNT = 2 is a compile-time constant,
 * /root/reference/src/risk_game/state/state.h:13; two owner bits with 2 = neutral, state.h:24-41), so there is nothing to pin this
 * file against.  It states SIXPLAYER.md; everything that has a two-player counterpart follows that counterpart's arithmetic and
 * cites it (paths relative to /root/reference/src/risk_game).  Never linked into or called by the product.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "az_philox.h"
#include "az_pseudo_net.h"
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

/* UtilityNN::makeMove (alphazero_moves.cpp:72-233) for six players; dice = the die stream of (seed, game, ply, sim), consumed
   from index *die on (a real move starts its own stream AZ_STREAM_REAL at 0; a descent of the search keeps one running index) */
static int make_move_stream(r6_state* s, int action, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply, uint32_t sim, uint32_t* die_io)
{
    if (r6_game_status(s, r) != R6_NOT_ENDED) return RO_ERR_GAME_OVER;
    if (action < 0 || action > R6_SKIP || !((r6_valid_moves(s, r) >> action) & 1)) return RO_ERR_ILLEGAL_ACTION;
    const int cur = s->cur;
    uint32_t die = *die_io;
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
            for (int i = 0; i < na; ++i) ad[i] = az_rng_die(seed, game, ply, sim, die++);               /* attacker dice first */
            for (int i = 0; i < nd; ++i) dd[i] = az_rng_die(seed, game, ply, sim, die++);
            *die_io = die;
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

int r6_make_move(r6_state* s, int action, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply)
{
    uint32_t die = 0;
    return make_move_stream(s, action, r, seed, game, ply, AZ_STREAM_REAL, &die);
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


/* ------------------------------------------------------------------ network input (SIXPLAYER.md): the reference's 13 planes
   (neural_network/alphazero_nn_data.cpp:165-196, alphazero_nn.cpp:31-67) re-read for six seats so that the SAME tower runs:
   plane 0 own army / 32, plane 1 the armies of the seat that moves next / 32, plane 2 every other seat's armies / 32, plane 3
   own / (own + all opponents') armies, plane 4 own / (own + living opponents') reinforcement values, plane 5 min(attacks / 8, 1),
   plane 6 may-draw-a-card, planes 7..12 the phase one-hot (SETUP_NEUTRAL never set).  Tensor [7][6][13], cell (y, x) = land y*6+x */
static int next_alive(const derived6* d, int cur)
{
    int n = cur;
    for (int k = 0; k < R6_PLAYERS; ++k) { n = (n + 1) % R6_PLAYERS; if (d->owned[n]) break; }
    return n;
}
void r6_encode(const r6_state* s, float x[RO_INPUT_FLOATS])
{
    derived6 d; derive(s, &d);
    const int c = s->cur, nx = next_alive(&d, c);
    int own_total = 0, other_total = 0;
    for (int i = 0; i < RO_LANDS; ++i) { if (s->owner[i] == c) own_total += s->army[i]; else other_total += s->army[i]; }
    float ref = (float)ro_reinforcement_value(d.owned[c]), oref = 0.0f;
    for (int p = 0; p < R6_PLAYERS; ++p) if (p != c && d.owned[p]) oref += (float)ro_reinforcement_value(d.owned[p]);
    const float army_share = (float)own_total / ((float)own_total + (float)other_total);
    const float reinf_share = ref / (ref + oref);
    float att = (float)s->attacks / 8.0f; if (att > 1.0f) att = 1.0f;
    for (int i = 0; i < RO_LANDS; ++i) {
        float* f = x + i * 13;
        const float a = (float)s->army[i] / 32.0f;
        f[0] = s->owner[i] == c ? a : 0.0f;
        f[1] = (s->owner[i] == nx && nx != c) ? a : 0.0f;
        f[2] = (s->owner[i] != c && s->owner[i] != nx) ? a : 0.0f;
        f[3] = army_share; f[4] = reinf_share; f[5] = att; f[6] = s->allow_draw ? 1.0f : 0.0f;
        for (int k = 0; k < 6; ++k) f[7 + k] = s->phase == k ? 1.0f : 0.0f;
    }
}

/* ------------------------------------------------------------------ search (SIXPLAYER.md "search value").  The reference's
   AlphaZeroMCTS (player/alpha_zero/alphazero_mcts.cpp: PUCT :67-119, search :322-377, addValue :8-21, policy from visits
   :121-148) with two changes: the table is cleared before every search (no trimNodes carry-over), and the scalar value of a leaf
   belongs to the seat to move there — a node whose mover is that seat receives +v, every other node -v / 5. */
typedef struct r6_node {
    r6_state key; uint64_t valid; float value; uint32_t sumN;
    float P[RO_MOVES], Q[RO_MOVES]; uint32_t N[RO_MOVES];
} r6_node;
struct r6_mcts { r6_node* nodes; int n_nodes, cap; r6_eval_fn eval; void* user; uint64_t evals; };

static uint64_t pn_key6(const r6_state* s)
{
    uint8_t land[RO_LANDS];
    for (int i = 0; i < RO_LANDS; ++i) land[i] = (uint8_t)(s->army[i] + 40 * s->owner[i]);     /* army <= 32 < 40, seat <= 5: one byte, unique */
    return az_pn_key(land, s->cur, s->round, s->phase);
}
void r6_eval_pseudo(const r6_state* s, float policy[RO_MOVES], float* value, void* user)
{
    (void)user;
    uint64_t key = pn_key6(s);
    for (int i = 0; i < RO_MOVES; ++i) policy[i] = az_pn_policy(key, i);
    *value = az_pn_value(key);
}
r6_mcts* r6_mcts_new(r6_eval_fn eval, void* user)
{
    r6_mcts* m = (r6_mcts*)calloc(1, sizeof *m);
    m->cap = 256; m->nodes = (r6_node*)malloc(sizeof(r6_node) * (size_t)m->cap);
    m->eval = eval ? eval : r6_eval_pseudo; m->user = user;
    return m;
}
void r6_mcts_free(r6_mcts* m) { if (m) { free(m->nodes); free(m); } }
int r6_mcts_table_size(const r6_mcts* m) { return m->n_nodes; }

static int find6(const r6_mcts* m, const r6_state* s)
{
    for (int i = 0; i < m->n_nodes; ++i) if (memcmp(&m->nodes[i].key, s, sizeof(r6_state)) == 0) return i;
    return -1;
}
static void expand6(r6_mcts* m, const r6_state* s, uint64_t valid, float* value_out)
{
    float policy[RO_MOVES], value = 0.0f;
    m->eval(s, policy, &value, m->user); m->evals++;
    ro_normalize_policy(policy, valid);
    if (m->n_nodes == m->cap) { m->cap *= 2; m->nodes = (r6_node*)realloc(m->nodes, sizeof(r6_node) * (size_t)m->cap); }
    r6_node* n = &m->nodes[m->n_nodes++];
    memset(n, 0, sizeof *n);
    n->key = *s; n->valid = valid; n->value = value;
    for (int i = 0; i < RO_MOVES; ++i) n->P[i] = ((valid >> i) & 1) ? policy[i] : 0.0f;
    *value_out = value;
}
static int select6(const r6_node* n, const ro_rules* r)             /* getNextBestMoveAndSetVisited, one descent at a time */
{
    int best = -1; float best_u = -INFINITY;
    for (int i = 0; i < RO_MOVES; ++i) {
        if (!((n->valid >> i) & 1)) continue;
        float noiseP = (1 - r->dir_noise_epsi) * n->P[i] + r->dir_noise_epsi * r->dir_noise_value;
        float v = noiseP * r->cpuct * sqrtf(1.0f + (float)n->sumN);
        float u = n->Q[i] + (v / (1.0f + (float)n->N[i]));
        if (u > best_u) { best_u = u; best = i; }
    }
    return best;
}
/* returns the leaf's value in *v and the seat it belongs to in *seat */
static void search6(r6_mcts* m, r6_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply, uint32_t sim, uint32_t* die,
                    int* seat, float* v, int* err)
{
    int gs = r6_game_status(s, r);
    if (gs != R6_NOT_ENDED) { *seat = gs == R6_DRAW ? s->cur : gs; *v = gs == R6_DRAW ? 0.0f : 1.0f; return; }
    uint64_t valid = r6_valid_moves(s, r);
    int idx = find6(m, s);
    if (idx < 0) { *seat = s->cur; expand6(m, s, valid, v); return; }
    int mv = select6(&m->nodes[idx], r);
    int cur = s->cur;
    if (make_move_stream(s, mv, r, seed, game, ply, sim, die) != RO_OK) { *err = 1; *seat = cur; *v = 0.0f; return; }
    search6(m, s, r, seed, game, ply, sim, die, seat, v, err);
    r6_node* n = &m->nodes[idx];                                     /* re-fetched: expand may have moved the array */
    float val = cur == *seat ? *v : -*v / 5.0f;
    if (n->N[mv] == 0) n->Q[mv] = val; else n->Q[mv] = ((float)n->N[mv] * n->Q[mv] + val) / (float)(n->N[mv] + 1);
    n->N[mv]++; n->sumN++;
}
int r6_mcts_search(r6_mcts* m, const r6_state* root, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply,
                   uint32_t N[RO_MOVES], float Q[RO_MOVES], float P[RO_MOVES], float pi[RO_MOVES], uint32_t* sumN, float* root_value)
{
    m->n_nodes = 0;
    float v0; expand6(m, root, r6_valid_moves(root, r), &v0);
    int count = r->mcts_simulations - (r->mcts_simulations % r->threads_per_mcts), err = 0;
    for (int i = 0; i < count; ++i) {
        r6_state copy = *root; uint32_t die = 0; int seat; float v;
        search6(m, &copy, r, seed, game, ply, (uint32_t)i, &die, &seat, &v, &err);
        if (err) return RO_ERR_ILLEGAL_ACTION;
    }
    const r6_node* n = &m->nodes[find6(m, root)];
    float sum = 0.0f;
    for (int i = 0; i < RO_MOVES; ++i) {
        int ok = (int)((n->valid >> i) & 1);
        N[i] = ok ? n->N[i] : 0; Q[i] = ok ? n->Q[i] : 0.0f; P[i] = ok ? n->P[i] : 0.0f;
        pi[i] = ok ? (float)n->N[i] : 0.0f;
        if (ok) sum += pi[i];
    }
    for (int i = 0; i < RO_MOVES; ++i) pi[i] /= sum;
    *sumN = n->sumN; *root_value = n->value;
    return RO_OK;
}
