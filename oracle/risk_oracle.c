/*
 * risk_oracle.c — TEST INFRASTRUCTURE (see risk_oracle.h).  Plain-C restatement of the
 * reference's self-play hot path.  Written from the reference's behaviour, not its
 * code: the state here is the ~54-byte primary state and every bit mask the reference
 * maintains incrementally (State::setLandArmy, state/state.cpp:279-385) is recomputed
 * from scratch, which is exactly the invariant State::consistencyCheck
 * (state/state.cpp:1209-1429) asserts.
 */
#include "risk_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "az_philox.h"
#include "az_pseudo_net.h"

/* ------------------------------------------------------------------ tables */
/* land/land.cpp:246-297: neighbour lists IN DECLARATION ORDER (the order is the tie-break
   of attack-from selection and of the fortify DFS) */
const int8_t RO_NBR_LIST[RO_LANDS][6] = {
    /* 0 ALASKA        */ { 1, 3, 29, -1, -1, -1 },
    /* 1 NORTHWEST_T   */ { 0, 3, 4, 2, -1, -1 },
    /* 2 GREENLAND     */ { 1, 4, 5, 13, -1, -1 },
    /* 3 ALBERTA       */ { 0, 1, 4, 6, -1, -1 },
    /* 4 ONTARIO       */ { 1, 3, 6, 7, 5, 2 },
    /* 5 QUEBEC        */ { 4, 7, 2, -1, -1, -1 },
    /* 6 WESTERN_US    */ { 3, 4, 7, 8, -1, -1 },
    /* 7 EASTERN_US    */ { 8, 6, 4, 5, -1, -1 },
    /* 8 CENTRAL_AM    */ { 6, 7, 9, -1, -1, -1 },
    /* 9 VENEZUELA     */ { 8, 10, 11, -1, -1, -1 },
    /* 10 PERU         */ { 9, 11, 12, -1, -1, -1 },
    /* 11 BRAZIL       */ { 9, 10, 12, 20, -1, -1 },
    /* 12 ARGENTINA    */ { 10, 11, -1, -1, -1, -1 },
    /* 13 ICELAND      */ { 2, 14, 15, -1, -1, -1 },
    /* 14 GREAT_BRITAIN*/ { 13, 19, 15, 17, -1, -1 },
    /* 15 SCANDINAVIA  */ { 13, 14, 16, 17, -1, -1 },
    /* 16 UKRAINE      */ { 15, 17, 18, 35, 33, 26 },
    /* 17 N_EUROPE     */ { 15, 14, 18, 19, 16, -1 },
    /* 18 S_EUROPE     */ { 19, 17, 16, 20, 21, 35 },
    /* 19 W_EUROPE     */ { 20, 14, 18, 17, -1, -1 },
    /* 20 N_AFRICA     */ { 11, 19, 18, 21, 23, 22 },
    /* 21 EGYPT        */ { 18, 20, 23, 35, -1, -1 },
    /* 22 CONGO        */ { 20, 23, 24, -1, -1, -1 },
    /* 23 E_AFRICA     */ { 21, 20, 22, 24, 25, 35 },
    /* 24 S_AFRICA     */ { 22, 23, 25, -1, -1, -1 },
    /* 25 MADAGASKAR   */ { 24, 23, -1, -1, -1, -1 },
    /* 26 URAL         */ { 16, 33, 34, 27, -1, -1 },
    /* 27 SIBERIA      */ { 26, 34, 32, 30, 28, -1 },
    /* 28 YAKUTSK      */ { 27, 30, 29, -1, -1, -1 },
    /* 29 KAMCHATKA    */ { 28, 30, 32, 31, 0, -1 },
    /* 30 IRKUTSK      */ { 28, 29, 32, 27, -1, -1 },
    /* 31 JAPAN        */ { 29, 32, -1, -1, -1, -1 },
    /* 32 MONGOLIA     */ { 27, 30, 29, 31, 34, -1 },
    /* 33 AFGHANISTAN  */ { 16, 26, 34, 36, 35, -1 },
    /* 34 CHINA        */ { 32, 27, 26, 33, 36, 37 },
    /* 35 MIDDLE_EAST  */ { 21, 23, 18, 16, 33, 36 },
    /* 36 INDIA        */ { 35, 33, 34, 37, -1, -1 },
    /* 37 SIAM         */ { 36, 34, 38, -1, -1, -1 },
    /* 38 INDONESIA    */ { 37, 39, 40, -1, -1, -1 },
    /* 39 NEW_GUINEA   */ { 38, 41, 40, -1, -1, -1 },
    /* 40 W_AUSTRALIA  */ { 41, 39, 38, -1, -1, -1 },
    /* 41 E_AUSTRALIA  */ { 40, 39, -1, -1, -1, -1 },
};

const uint64_t RO_NBR_MASK[RO_LANDS] = {
    0x0002000000aull, 0x0000000001dull, 0x00000002032ull, 0x00000000053ull, 0x000000000eeull, 0x00000000094ull,
    0x00000000198ull, 0x00000000170ull, 0x000000002c0ull, 0x00000000d00ull, 0x00000001a00ull, 0x00000101600ull,
    0x00000000c00ull, 0x0000000c004ull, 0x000000aa000ull, 0x00000036000ull, 0x00a04068000ull, 0x000000dc000ull,
    0x008003b0000ull, 0x00000164000ull, 0x00000ec0800ull, 0x00800940000ull, 0x00001900000ull, 0x00803700000ull,
    0x00002c00000ull, 0x00001800000ull, 0x00608010000ull, 0x00554000000ull, 0x00068000000ull, 0x001d0000001ull,
    0x00138000000ull, 0x00120000000ull, 0x004e8000000ull, 0x01c04010000ull, 0x0330c000000ull, 0x01200a50000ull,
    0x02e00000000ull, 0x05400000000ull, 0x1a000000000ull, 0x34000000000ull, 0x2c000000000ull, 0x18000000000ull,
};

/* land/land_set.cpp:12-24 + land/land_index.h:5-10 — NA, SA, EU, AF, AS, AU */
const uint64_t RO_CONTINENT_MASK[6] = { 0x1ffull, 0x1e00ull, 0xfe000ull, 0x3f00000ull, 0x3ffc000000ull, 0x3c000000000ull };
const int RO_CONTINENT_BONUS[6] = { 5, 2, 5, 3, 7, 2 };

#define ALL_LANDS 0x3ffffffffffull   /* LandSet::ALL_LANDS_MASK, land/land_set.cpp:27-33 */
#define SKIP_MASK (1ull << RO_SKIP)  /* Land::SKIP_MOVE_MASK, land/land.cpp:313 */

static inline int army_of(const ro_state* s, int i) { return s->land[i] & 63; }
static inline int owner_of(const ro_state* s, int i) { return s->land[i] >> 6; }
static inline void set_land(ro_state* s, int i, int army, int owner) { s->land[i] = (uint8_t)((army & 63) | (owner << 6)); }
static inline int popc(uint64_t x) { return __builtin_popcountll(x); }
static inline int ctz(uint64_t x) { return __builtin_ctzll(x); }

void ro_default_rules(ro_rules* r)
{   /* settings.h:40-62 */
    r->allow_yield = 1; r->limit_reinforcement = 1; r->limit_attack = 0; r->max_game_rounds = 58; r->min_unit_move = 3;
    r->mcts_simulations = 32; r->threads_per_mcts = 2; r->cpuct = 1.1f; r->dir_noise_value = 0.3f; r->dir_noise_epsi = 0.25f;
    r->temperature_threshold = 43;
}

/* ------------------------------------------------------------------ dice */
void ro_dice_philox(ro_dice* d, uint64_t seed, uint32_t game, uint32_t ply, uint32_t sim)
{
    memset(d, 0, sizeof *d);
    d->seed = seed; d->game = game; d->ply = ply; d->sim = sim;
}
void ro_dice_tape(ro_dice* d, const int32_t* tape, int n)
{
    memset(d, 0, sizeof *d);
    d->use_tape = 1; d->tape = tape; d->tape_len = n;
}
static int next_die(ro_dice* d)
{
    if (d->use_tape) {
        if (d->tape_pos >= d->tape_len) abort();
        return d->tape[d->tape_pos++];
    }
    return az_rng_die(d->seed, d->game, d->ply, d->sim, d->j++);
}

/* ------------------------------------------------------------------ derived masks */
typedef struct derived {
    uint64_t owned[2], owned_army[2], full[2], attack[2], attack_army[2], neutral;
    int total_army[2];
} derived;

static uint64_t nbr_union(uint64_t set)
{
    uint64_t u = 0;
    while (set) { int i = ctz(set); set &= set - 1; u |= RO_NBR_MASK[i]; }
    return u;
}

/* definition of the five PlayerStatus masks + totalArmy as functions of landArmy[]
   (what State::setLandArmy maintains incrementally, state/state.cpp:279-385, and what
   State::consistencyCheck recomputes, state/state.cpp:1209-1429) */
static void derive(const ro_state* s, derived* d)
{
    memset(d, 0, sizeof *d);
    for (int i = 0; i < RO_LANDS; ++i) {
        int o = owner_of(s, i), a = army_of(s, i);
        uint64_t m = 1ull << i;
        if (o < 2) {
            d->owned[o] |= m; d->total_army[o] += a;
            if (a > 1) d->owned_army[o] |= m;
            if (a == RO_ARMY_MAX) d->full[o] |= m;
        } else d->neutral |= m;
    }
    for (int p = 0; p < 2; ++p) {
        d->attack[p] = nbr_union(d->owned[p]) & ~d->owned[p];
        d->attack_army[p] = nbr_union(d->owned_army[p]) & ~d->owned[p];
    }
}

/* ------------------------------------------------------------------ Data image */
static void put48(uint8_t* p, uint64_t v) { for (int i = 0; i < 6; ++i) p[i] = (uint8_t)(v >> (8 * i)); }
static uint64_t get48(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 6; ++i) v |= (uint64_t)p[i] << (8 * i); return v; }

/* layout of struct Data, state/state.h:86-105 as laid out by g++ 13.3 (x86-64):
   landArmy[42] @0; playerStatus[2] @48 (48 B each: five 48-bit masks in five 8-byte units
   @0,8,16,24,32; totalArmy i16 packed @38; playerCards u8 @40); round u16 @144;
   currentPlayerTurn @146; cardSetsPlayed @147; reinforcements @148; roundPhase @149;
   attackMobilizationFrom/To @150/151; playerAllowedDrawCard @152; attacksDuringTurn @153;
   drawnCardsBitMask u16 @154; everything else is padding. */
void ro_export_data(const ro_state* s, uint8_t data[RO_DATA_BYTES])
{
    derived d; derive(s, &d);
    memset(data, 0, RO_DATA_BYTES);
    memcpy(data, s->land, RO_LANDS);
    for (int p = 0; p < 2; ++p) {
        uint8_t* ps = data + 48 + 48 * p;
        put48(ps + 0, d.owned[p]); put48(ps + 8, d.owned_army[p]); put48(ps + 16, d.full[p]);
        put48(ps + 24, d.attack[p]); put48(ps + 32, d.attack_army[p]);
        ps[38] = (uint8_t)(d.total_army[p] & 0xff); ps[39] = (uint8_t)((d.total_army[p] >> 8) & 0xff);
        ps[40] = s->cards[p];
    }
    data[144] = (uint8_t)(s->round & 0xff); data[145] = (uint8_t)(s->round >> 8);
    data[146] = (uint8_t)s->cur; data[147] = s->card_sets; data[148] = s->reinf; data[149] = s->phase;
    data[150] = s->mob_from; data[151] = s->mob_to; data[152] = s->allow_draw; data[153] = s->attacks;
}

void ro_data_byte_mask(uint8_t mask[RO_DATA_BYTES])
{
    memset(mask, 0, RO_DATA_BYTES);
    memset(mask, 1, RO_LANDS);
    for (int p = 0; p < 2; ++p) {
        uint8_t* ps = mask + 48 + 48 * p;
        for (int w = 0; w < 5; ++w) memset(ps + 8 * w, 1, 6);
        ps[38] = ps[39] = 1; ps[40] = 1;
    }
    memset(mask + 144, 1, 12); /* 144..155 */
}

int ro_import_data(ro_state* s, const uint8_t data[RO_DATA_BYTES])
{
    memset(s, 0, sizeof *s);
    memcpy(s->land, data, RO_LANDS);
    s->cards[0] = data[48 + 40]; s->cards[1] = data[96 + 40];
    s->round = (uint16_t)(data[144] | (data[145] << 8));
    s->cur = (int8_t)data[146]; s->card_sets = data[147]; s->reinf = data[148]; s->phase = data[149];
    s->mob_from = data[150]; s->mob_to = data[151]; s->allow_draw = data[152]; s->attacks = data[153];
    /* the masks in the image must agree with landArmy[] */
    derived d; derive(s, &d);
    for (int p = 0; p < 2; ++p) {
        const uint8_t* ps = data + 48 + 48 * p;
        if (get48(ps) != d.owned[p] || get48(ps + 8) != d.owned_army[p] || get48(ps + 16) != d.full[p] ||
            get48(ps + 24) != d.attack[p] || get48(ps + 32) != d.attack_army[p]) return RO_ERR_BAD_STATE;
        if ((int16_t)(ps[38] | (ps[39] << 8)) != d.total_army[p]) return RO_ERR_BAD_STATE;
    }
    if (s->cur < 0 || s->cur > 1 || s->phase > RO_FORTIFY) return RO_ERR_BAD_STATE;
    return RO_OK;
}

/* ------------------------------------------------------------------ rules */
/* State::calculateReinforcementValue, state/state.cpp:457-491 */
int ro_reinforcement_value(uint64_t owned)
{
    int v = popc(owned) / 3;
    for (int c = 0; c < 6; ++c)
        if ((owned & RO_CONTINENT_MASK[c]) == RO_CONTINENT_MASK[c]) v += RO_CONTINENT_BONUS[c];
    return v < 3 ? 3 : v;
}

/* State::gameStatus, state/state.cpp:518-565 */
int ro_game_status(const ro_state* s, const ro_rules* r)
{
    int p0 = 0, p1 = 0;
    for (int i = 0; i < RO_LANDS; ++i) { int o = owner_of(s, i); p0 += o == 0; p1 += o == 1; }
    if (p0 == 0) return 1;
    if (p1 == 0) return 0;
    if (r->allow_yield) { if (p0 >= 30) return 0; if (p1 >= 30) return 1; }
    if ((int)s->round > r->max_game_rounds) return p0 > p1 ? 0 : p0 < p1 ? 1 : RO_DRAW;
    return RO_NOT_ENDED;
}

/* UtilityNN::getValidMoves, player/alpha_zero/alphazero_moves.cpp:3-70 */
uint64_t ro_valid_moves(const ro_state* s, const ro_rules* r)
{
    derived d; derive(s, &d);
    int c = s->cur, e = c ^ 1;
    switch (s->phase) {
    case RO_SETUP:
    case RO_REINFORCEMENT: {
        uint64_t o = d.owned[c] & ~d.full[c];
        if (o == 0) return SKIP_MASK;
        if (r->limit_reinforcement) {
            /* State::getNeutralPlayerAttackLands, state/state.cpp:1067-1083 */
            uint64_t neutral_attack = nbr_union(d.neutral) & ~d.neutral;
            uint64_t b = o & (d.attack[e] | neutral_attack);
            return b ? b : o;
        }
        return o;
    }
    case RO_SETUP_NEUTRAL: return ALL_LANDS & ~d.owned[c] & ~d.owned[e];
    case RO_ATTACK:
        if (r->limit_attack) return d.attack_army[c] ? d.attack_army[c] : SKIP_MASK;
        return d.attack_army[c] | SKIP_MASK;
    case RO_ATTACK_MOBILIZATION: return (1ull << s->mob_from) | (1ull << s->mob_to);
    case RO_FORTIFY:
        if (r->limit_reinforcement) return (d.owned[c] & d.attack[e]) | SKIP_MASK;
        return d.owned[c] | SKIP_MASK;
    }
    return 0;
}

static uint64_t attack_army_mask(const ro_state* s, int p) { derived d; derive(s, &d); return d.attack_army[p]; }

/* State::gotoAttack, state/state.cpp:20-40 (+ gotoFortify :42-49) */
static void goto_attack(ro_state* s)
{
    s->phase = RO_ATTACK; s->mob_from = RO_NONE; s->mob_to = RO_NONE; s->reinf = 0;
    if (attack_army_mask(s, s->cur) == 0) s->phase = RO_FORTIFY;
}

/* State::nextPlayerGameTurn, state/state.cpp:748-766 (+ drawCard simple mode :618-626) */
static void end_turn(ro_state* s)
{
    if (s->allow_draw) { s->cards[s->cur] = (uint8_t)(s->cards[s->cur] + 1); s->allow_draw = 0; }
    s->round++;
    s->cur ^= 1;
    s->attacks = 0;
    s->phase = RO_REINFORCEMENT;
    derived d; derive(s, &d);
    s->reinf = (uint8_t)ro_reinforcement_value(d.owned[s->cur]);
}

/* GameHelper::LandSetMovement::add, player/game_helper.cpp:51-82 — DFS pre-order */
static void dfs(int l, uint64_t owned, uint64_t* seen, int* order, int* n)
{
    if (!((1ull << l) & owned & ~*seen)) return;
    *seen |= 1ull << l;
    order[(*n)++] = l;
    for (int k = 0; k < 6 && RO_NBR_LIST[l][k] >= 0; ++k) dfs(RO_NBR_LIST[l][k], owned, seen, order, n);
}

static void sort_dice_desc(int* d, int n)
{
    for (int i = 0; i < n; ++i)
        for (int k = i + 1; k < n; ++k)
            if (d[k] > d[i]) { int t = d[i]; d[i] = d[k]; d[k] = t; }
}

/* UtilityNN::makeMove, player/alpha_zero/alphazero_moves.cpp:72-233.  The accepted actions
   are exactly the bits of ro_valid_moves (what every caller on the path produces); the
   reference throws for most other inputs, here the state is left untouched. */
int ro_make_move(ro_state* s, int action, const ro_rules* r, ro_dice* dice)
{
    if (ro_game_status(s, r) != RO_NOT_ENDED) return RO_ERR_GAME_OVER;
    if (action < 0 || action > RO_SKIP || !((ro_valid_moves(s, r) >> action) & 1)) return RO_ERR_ILLEGAL_ACTION;
    int cur = s->cur;
    if (action == RO_SKIP) { /* alphazero_moves.cpp:79-92 */
        switch (s->phase) {
        case RO_REINFORCEMENT: goto_attack(s); break;
        case RO_ATTACK: s->phase = RO_FORTIFY; break;
        case RO_FORTIFY: end_turn(s); break;
        default: /* SETUP with every owned land full: the reference throws logic_error */ return RO_ERR_ILLEGAL_ACTION;
        }
        return RO_OK;
    }
    int li = action;
    switch (s->phase) {
    case RO_SETUP: /* State::setupReinforcementMove, state/state.cpp:1009-1030 */
        s->reinf = (uint8_t)(s->reinf - 2);
        set_land(s, li, army_of(s, li) + 2, cur);
        s->phase = RO_SETUP_NEUTRAL;
        break;
    case RO_SETUP_NEUTRAL: /* setupReinforcementNeutralMove :1032-1053 + nextPlayerSetupTurn :725-746 */
        set_land(s, li, army_of(s, li) + 1, RO_NEUTRAL);
        s->phase = RO_SETUP; s->round++; s->cur ^= 1;
        if (s->reinf == 0) {
            derived d; derive(s, &d);
            s->phase = RO_REINFORCEMENT; s->reinf = (uint8_t)ro_reinforcement_value(d.owned[s->cur]);
        }
        break;
    case RO_REINFORCEMENT: { /* alphazero_moves.cpp:104-121; GameHelper::playCards game_helper.cpp:3-17; State::playCards state.cpp:1091-1117 */
        if (s->cards[cur] >= 3) {
            s->cards[cur] = (uint8_t)(s->cards[cur] - 3);
            s->card_sets = (uint8_t)(s->card_sets + 1);
            int gained;
            switch (s->card_sets) {
            case 1: gained = 4; break; case 2: gained = 6; break; case 3: gained = 8; break;
            case 4: gained = 10; break; case 5: gained = 12; break; case 6: gained = 15; break;
            default: gained = 15 + (s->card_sets - 6) * 5; break;
            }
            s->reinf = (uint8_t)(s->reinf + gained);
        }
        int rf = s->reinf / 2;                                           /* FAST_ATTACK_MOBILIZATION branch */
        if (rf < r->min_unit_move) rf = r->min_unit_move < s->reinf ? r->min_unit_move : s->reinf;
        int space = RO_ARMY_MAX - army_of(s, li);
        if (space < rf) rf = space;
        s->reinf = (uint8_t)(s->reinf - rf);                             /* State::reinforcementMove :976-998 */
        set_land(s, li, army_of(s, li) + rf, cur);
        if (s->reinf == 0) goto_attack(s);
        break;
    }
    case RO_ATTACK: { /* alphazero_moves.cpp:122-145 then State::attackMove state.cpp:769-918 */
        int best = 0, from = -1;
        for (int k = 0; k < 6 && RO_NBR_LIST[li][k] >= 0; ++k) {
            int n = RO_NBR_LIST[li][k];
            if (owner_of(s, n) == cur && army_of(s, n) > 1) {
                int v = army_of(s, n) - 1;
                if (v > best) { best = v; from = n; }
            }
        }
        if (from < 0) return RO_ERR_ILLEGAL_ACTION;
        s->attacks = (uint8_t)(s->attacks + 1);
        int a = army_of(s, from), d = army_of(s, li), defender = owner_of(s, li), units = 1;
        if (d > 0) {
            int na = a >= 4 ? 3 : a == 3 ? 2 : 1;
            int nd = d >= 2 ? 2 : 1;
            units = na;
            int ad[3] = { 0, 0, 0 }, dd[3] = { 0, 0, 0 };
            for (int i = 0; i < na; ++i) ad[i] = next_die(dice);         /* attacker dice first, state.cpp:832 */
            for (int i = 0; i < nd; ++i) dd[i] = next_die(dice);         /* then defender, :833 */
            sort_dice_desc(ad, na); sort_dice_desc(dd, nd);
            if (ad[0] > dd[0]) d--; else { a--; units--; }
            if (na >= 2 && nd == 2) { if (ad[1] > dd[1]) d--; else { a--; units--; } }
        }
        if (d == 0) {
            a -= units;
            if (a > 1) { s->phase = RO_ATTACK_MOBILIZATION; s->mob_from = (uint8_t)from; s->mob_to = (uint8_t)li; }
            s->allow_draw = 1;
            set_land(s, from, a, cur);
            set_land(s, li, units, cur);
        } else {
            set_land(s, from, a, cur);
            set_land(s, li, d, defender);
        }
        if (s->phase == RO_ATTACK && attack_army_mask(s, cur) == 0) s->phase = RO_FORTIFY;
        break;
    }
    case RO_ATTACK_MOBILIZATION: /* alphazero_moves.cpp:146-171; State::attackReinforcementMove state.cpp:920-947 */
        if (li == s->mob_from) goto_attack(s);
        else {
            int from = s->mob_from, to = s->mob_to;
            int v = army_of(s, from) - 1;
            int rf = v / 2;
            if (rf < r->min_unit_move) rf = r->min_unit_move < v ? r->min_unit_move : v;
            set_land(s, from, army_of(s, from) - rf, cur);
            set_land(s, to, army_of(s, to) + rf, cur);
            if (army_of(s, from) == 1) goto_attack(s);
        }
        break;
    case RO_FORTIFY: { /* alphazero_moves.cpp:172-231; GameHelper::PlayerMovement game_helper.cpp:90-109 */
        if (army_of(s, li) != RO_ARMY_MAX) {
            derived dd; derive(s, &dd);
            uint64_t owned = dd.owned[cur], grouped = 0;
            for (int seed = 0; seed < RO_LANDS; ++seed) {
                if (!((1ull << seed) & owned & ~grouped)) continue;
                uint64_t seen = 0; int order[RO_LANDS], n = 0;
                dfs(seed, owned, &seen, order, &n);
                grouped |= seen;
                if (!((seen >> li) & 1)) continue;
                int best_i = 0, from_i = -1, best_b = 0, from_b = -1;
                for (int j = 0; j < n; ++j) {
                    int l = order[j];
                    if (l == li) continue;
                    int v = army_of(s, l) - 1;
                    if ((RO_NBR_MASK[l] & owned) == RO_NBR_MASK[l]) { if (v > best_i) { best_i = v; from_i = l; } }
                    else { if (v > best_b) { best_b = v; from_b = l; } }
                }
                if (from_i >= 0) { from_b = from_i; best_b = best_i; }
                if (from_b >= 0) {
                    int space = RO_ARMY_MAX - army_of(s, li);
                    int mv = space < best_b ? space : best_b;            /* State::fortifyMove state.cpp:949-974 */
                    set_land(s, from_b, army_of(s, from_b) - mv, cur);
                    set_land(s, li, army_of(s, li) + mv, cur);
                }
                break;
            }
        }
        end_turn(s);
        break;
    }
    }
    return RO_OK;
}

/* State::newGame, state/state.cpp:137-167 with Utility::randomMask, land/land.cpp:100-112 */
static void deal(ro_state* s, const uint32_t k42[42])
{
    memset(s, 0, sizeof *s);
    for (int i = 0; i < RO_LANDS; ++i) s->land[i] = (uint8_t)(RO_NEUTRAL << 6);
    s->round = 1; s->phase = RO_SETUP; s->mob_from = RO_NONE; s->mob_to = RO_NONE;
    uint64_t avail = ALL_LANDS; int i = 0, cur = 0;
    while (avail) {
        uint64_t m = avail; for (uint32_t t = 0; t < k42[i]; ++t) m &= m - 1;
        int l = ctz(m); avail &= ~(1ull << l); i++;
        set_land(s, l, 1, cur);
        if (cur == 1) {
            m = avail; for (uint32_t t = 0; t < k42[i]; ++t) m &= m - 1;
            l = ctz(m); avail &= ~(1ull << l); i++;
            set_land(s, l, 1, RO_NEUTRAL);
        }
        cur ^= 1;
    }
    s->cur = (int8_t)cur;
    s->reinf = (40 - 14) * 2;
}

void ro_new_game(ro_state* s, uint64_t seed, uint32_t game, uint32_t ply)
{
    uint32_t k[42];
    for (uint32_t i = 0; i < 42; ++i) k[i] = az_rng_deal_draw(seed, game, ply, i);
    deal(s, k);
}

void ro_new_game_tape(ro_state* s, const int32_t* draws42)
{
    uint32_t k[42];
    for (int i = 0; i < 42; ++i) k[i] = (uint32_t)draws42[i] % (uint32_t)(42 - i);
    deal(s, k);
}

int ro_random_action(const ro_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply)
{
    uint64_t m = ro_valid_moves(s, r);
    az_u32x4 b = az_rng_block(seed, game, ply, AZ_STREAM_REAL, 0);
    uint32_t k = az_mulhi32(b.y, (uint32_t)popc(m));
    for (uint32_t i = 0; i < k; ++i) m &= m - 1;
    return ctz(m);
}

/* NNInputData(const State&), neural_network/alphazero_nn_data.cpp:165-196 and
   setInStateTensor, neural_network/alphazero_nn.cpp:31-67; channel indices
   alphazero_nn_data.h:13-39 (INPUT_VECTOR_TYPE_2): tensor [7][6][13], cell (y,x) = land y*6+x */
void ro_encode(const ro_state* s, float x[RO_INPUT_FLOATS])
{
    derived d; derive(s, &d);
    int c = s->cur, e = c ^ 1;
    float ref = (float)(int8_t)ro_reinforcement_value(d.owned[c]);
    float eref = (float)(int8_t)ro_reinforcement_value(d.owned[e]);
    float reinf_share = ref / (ref + eref);
    float att = (float)s->attacks / 8.0f; if (!(att < 1.0f)) att = 1.0f;
    float ta = (float)d.total_army[c], eta = (float)d.total_army[e];
    float army_share = ta / (ta + eta);
    for (int i = 0; i < RO_LANDS; ++i) {
        float* f = x + i * 13;
        float fa = (float)army_of(s, i) / (float)RO_ARMY_MAX;
        int o = owner_of(s, i);
        f[0] = o == c ? fa : 0.0f;
        f[1] = o == e ? fa : 0.0f;
        f[2] = o == RO_NEUTRAL ? fa : 0.0f;
        f[3] = army_share;
        f[4] = reinf_share;
        f[5] = att;
        f[6] = s->allow_draw ? 1.0f : 0.0f;
        for (int p = 0; p < 6; ++p) f[7 + p] = s->phase == p ? 1.0f : 0.0f;
    }
}


/* ------------------------------------------------------------------ scripted opponent
   ScriptPlayer::takeTurn, player/script/script_player.cpp:162-227 (and :19-160), GameHelper::PlayerMovement /
   LandSetMovement::add, player/game_helper.cpp:51-109.  One call plays the WHOLE turn (or, in the setup phase, the
   player's placement and the neutral placement).  Dice come from the sequential stream AZ_STREAM_OPP of (game, ply), the
   rInt() of Utility::randomMask from AZ_STREAM_OPP_INT (include/az_philox.h). */

/* land/land_set.cpp:12-24: member lands of each continent IN DECLARATION ORDER (first attackable one is attacked);
   continent ids here: 0 NA, 1 SA, 2 EU, 3 AF, 4 AS, 5 AU as in RO_CONTINENT_MASK */
static const int8_t RO_CONTINENT_LANDS[6][13] = {
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, -1 },
    { 9, 10, 11, 12, -1 },
    { 13, 14, 15, 16, 17, 19, 18, -1 },
    { 20, 21, 22, 24, 25, 23, -1 },
    { 26, 33, 35, 36, 27, 28, 29, 30, 31, 32, 34, 37, -1 },
    { 38, 39, 40, 41, -1 },
};

void ro_script_init(ro_script* sp) { sp->set = -1; sp->to = -1; sp->from = -1; sp->from_army = 0; }

typedef struct script_ctx {
    ro_state* s; ro_script* sp; const ro_rules* r;
    uint64_t seed; uint32_t game, ply, die_j, int_j;
    uint64_t owned_attack_mask, attack_mask;     /* ScriptPlayer::ownedAttackLandBitMask / attackLandBitMask */
    ro_turn_sink* sink;
} script_ctx;

/* Player::addTrainingSample(state, move), player/base/player.cpp:9-17 */
static void emit(script_ctx* c, int move)
{
    ro_turn_sink* k = c->sink;
    if (!k) return;
    if (k->n < k->cap) { k->states[k->n] = *c->s; k->moves[k->n] = (uint8_t)move; }
    k->n++;
}

static int script_die(script_ctx* c) { return az_rng_die(c->seed, c->game, c->ply, AZ_STREAM_OPP, c->die_j++); }

/* GameHelper::sortLandSet, game_helper.cpp:19-39: fewer lands still to conquer first, then more attackable ones, then the
   larger continent mask — a strict total order, so the previous order of ScriptPlayer::attackLandSetPriority is irrelevant */
static int set_before(int a, int na, int aa, int b, int nb, int ab)
{
    if (na != nb) return na < nb;
    if (aa != ab) return aa > ab;
    return RO_CONTINENT_MASK[a] > RO_CONTINENT_MASK[b];
}

/* updateAttackLandSetPriority / updateAttackLandSet / updateAttackLandTo / updateAttackLandFrom, script_player.cpp:19-70;
   members that find no candidate keep their previous value, like the reference's */
static void script_pick_target(script_ctx* c)
{
    derived d; derive(c->s, &d);
    int me = c->s->cur;
    int not_owned[6], not_owned_attack[6], order[6];
    for (int k = 0; k < 6; ++k) {
        uint64_t m = RO_CONTINENT_MASK[k] & ~d.owned[me];
        not_owned[k] = popc(m); not_owned_attack[k] = popc(m & c->attack_mask);
        order[k] = k;
    }
    for (int i = 1; i < 6; ++i)          /* any sort gives the same result for a strict total order */
        for (int j = i; j > 0 && set_before(order[j], not_owned[order[j]], not_owned_attack[order[j]],
                                            order[j - 1], not_owned[order[j - 1]], not_owned_attack[order[j - 1]]); --j) {
            int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
        }
    for (int i = 0; i < 6; ++i) if (not_owned_attack[order[i]] > 0) { c->sp->set = (int8_t)order[i]; break; }
    if (c->sp->set >= 0)
        for (int i = 0; RO_CONTINENT_LANDS[c->sp->set][i] >= 0; ++i) {
            int l = RO_CONTINENT_LANDS[c->sp->set][i];
            if ((1ull << l) & c->attack_mask) { c->sp->to = (int8_t)l; break; }
        }
    c->sp->from_army = 0;
    if (c->sp->to >= 0)
        for (int k = 0; k < 6 && RO_NBR_LIST[c->sp->to][k] >= 0; ++k) {
            int n = RO_NBR_LIST[c->sp->to][k];
            if (((1ull << n) & c->owned_attack_mask) && army_of(c->s, n) > c->sp->from_army) {
                c->sp->from_army = (uint8_t)army_of(c->s, n); c->sp->from = (int8_t)n;
            }
        }
}

/* State::attackMove, state/state.cpp:769-918; returns 1 when the land was captured */
static int script_attack(script_ctx* c, int from, int to)
{
    ro_state* s = c->s;
    int cur = s->cur;
    s->attacks = (uint8_t)(s->attacks + 1);
    int a = army_of(s, from), dd = army_of(s, to), defender = owner_of(s, to), units = 1;
    if (dd > 0) {
        int na = a >= 4 ? 3 : (a == 3 ? 2 : 1), nd = dd >= 2 ? 2 : 1;
        units = na;
        int ad[3] = { 0, 0, 0 }, df[2] = { 0, 0 };
        for (int i = 0; i < na; ++i) ad[i] = script_die(c);
        for (int i = 0; i < nd; ++i) df[i] = script_die(c);
        sort_dice_desc(ad, 3); sort_dice_desc(df, 2);
        if (ad[0] > df[0]) dd--; else { a--; units--; }
        if (na >= 2 && nd == 2) { if (ad[1] > df[1]) dd--; else { a--; units--; } }
    }
    int captured = 0;
    if (dd == 0) {
        a -= units;
        if (a > 1) { s->phase = RO_ATTACK_MOBILIZATION; s->mob_from = (uint8_t)from; s->mob_to = (uint8_t)to; }
        s->allow_draw = 1;
        set_land(s, from, a, cur); set_land(s, to, units, cur);
        captured = 1;
    } else { set_land(s, from, a, cur); set_land(s, to, dd, defender); }
    if (s->phase == RO_ATTACK && attack_army_mask(s, cur) == 0) s->phase = RO_FORTIFY;
    return captured;
}

/* ScriptPlayer::attackLand, script_player.cpp:71-136 */
static void script_attack_land(script_ctx* c)
{
    ro_state* s = c->s; ro_script* sp = c->sp;
    int me = s->cur;
    while (s->reinf > 0) {
        derived d; derive(s, &d);
        uint64_t not_full = d.owned[me] & ~d.full[me];
        int to = sp->from;
        if (((1ull << sp->from) & not_full) == 0) {
            uint64_t nb = RO_NBR_MASK[sp->to] & not_full;
            if (nb) to = ctz(nb);
            else {
                uint64_t neutral_attack = nbr_union(d.neutral) & ~d.neutral;
                nb = not_full & (d.attack[me ^ 1] | neutral_attack);
                to = nb ? ctz(nb) : ctz(not_full);
            }
        }
        int amount = RO_ARMY_MAX - army_of(s, to);
        if (s->reinf < amount) amount = s->reinf;
        while (amount > 0) {                       /* State::reinforcementMove, state.cpp:976-998 */
            int step = amount < c->r->min_unit_move ? amount : c->r->min_unit_move;
            emit(c, to);                                                              /* script_player.cpp:105 */
            s->reinf = (uint8_t)(s->reinf - step);
            set_land(s, to, army_of(s, to) + step, me);
            if (s->reinf == 0) goto_attack(s);
            amount -= step;
        }
    }
    sp->from_army = (uint8_t)army_of(s, sp->from);
    while (sp->from_army > 1) {
        emit(c, sp->to);                                                              /* :115 */
        int captured = script_attack(c, sp->from, sp->to);
        sp->from_army = (uint8_t)army_of(s, sp->from);
        if (captured && sp->from_army > 1) {       /* move everything but one army into the new land, MIN_UNIT_MOVE at a time */
            int left = sp->from_army - 1;
            while (left > 0) {                     /* State::attackReinforcementMove, state.cpp:920-947 */
                int step = left < c->r->min_unit_move ? left : c->r->min_unit_move;
                emit(c, sp->to);                                                      /* :125 */
                left -= step;
                set_land(s, s->mob_from, army_of(s, s->mob_from) - step, me);
                set_land(s, s->mob_to, army_of(s, s->mob_to) + step, me);
                if (army_of(s, s->mob_from) == 1) goto_attack(s);
            }
            break;
        }
    }
}

/* ScriptPlayer::fortify, script_player.cpp:138-160 over GameHelper::PlayerMovement, game_helper.cpp:84-109: the owned components in
   ascending seed order, each listed in DFS pre-order; per component the interior land with the largest army (first strict
   maximum) is the source, the border land with the most foreign neighbours (first strict maximum) the target; the component
   with the largest source army is used (std::sort on <= 16 elements is libstdc++'s stable insertion sort: first maximum). */
static void script_fortify(script_ctx* c)
{
    ro_state* s = c->s;
    int me = s->cur;
    derived d; derive(s, &d);
    if (d.owned_army[me] == 0) return;
    uint64_t owned = d.owned[me], seen = 0;
    int best_from = -1, best_to = -1, best_amount = -1;
    for (int i = 0; i < RO_LANDS; ++i) {
        if (!((1ull << i) & owned & ~seen)) continue;
        int order[RO_LANDS], n = 0;
        dfs(i, owned, &seen, order, &n);
        int from = -1, from_amount = 0, to = -1, to_nbrs = 0;
        for (int k = 0; k < n; ++k) {
            int l = order[k];
            uint64_t foreign = ~owned & RO_NBR_MASK[l];
            if (foreign == 0) { if (army_of(s, l) > from_amount) { from = l; from_amount = army_of(s, l); } }
            else if (popc(foreign) > to_nbrs) { to_nbrs = popc(foreign); to = l; }
        }
        if (from_amount > best_amount) { best_amount = from_amount; best_from = from; best_to = to; }
    }
    if (best_amount > 0 && best_to >= 0) {          /* State::fortifyMove, state.cpp:949-974 */
        emit(c, best_to);                                                             /* :151 */
        int amount = army_of(s, best_from) - 1, space = RO_ARMY_MAX - army_of(s, best_to);
        if (space < amount) amount = space;
        set_land(s, best_from, army_of(s, best_from) - amount, me);
        set_land(s, best_to, army_of(s, best_to) + amount, me);
    } else emit(c, RO_SKIP);                                                          /* :157 */
}

int ro_script_turn(ro_state* s, ro_script* sp, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply)
{
    return ro_script_turn_rec(s, sp, r, seed, game, ply, NULL);
}

int ro_script_turn_rec(ro_state* s, ro_script* sp, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply, ro_turn_sink* sink)
{
    if (ro_game_status(s, r) != RO_NOT_ENDED) return RO_ERR_GAME_OVER;
    script_ctx c; memset(&c, 0, sizeof c);
    c.s = s; c.sp = sp; c.r = r; c.seed = seed; c.game = game; c.ply = ply; c.sink = sink;
    int me = s->cur;
    derived d; derive(s, &d);
    c.owned_attack_mask = d.owned[me]; c.attack_mask = d.attack[me];
    if (s->phase == RO_SETUP) {
        script_pick_target(&c);
        emit(&c, sp->from);                                       /* script_player.cpp:176 */
        s->reinf = (uint8_t)(s->reinf - 2);                       /* setupReinforcementMove, state.cpp:1009-1030 */
        set_land(s, sp->from, army_of(s, sp->from) + 2, me);
        s->phase = RO_SETUP_NEUTRAL;
        uint64_t neutral = ALL_LANDS & ~d.owned[0] & ~d.owned[1];
        uint64_t near_enemy = neutral & d.attack[me ^ 1] & ~d.attack[me];
        if (near_enemy == 0) near_enemy = neutral & d.attack[me ^ 1];
        uint64_t pool = near_enemy ? near_enemy : neutral;
        int k = (int)(az_rng_opp_int(seed, game, ply, c.int_j++) % (uint32_t)popc(pool));   /* Utility::randomMask */
        while (k--) pool &= pool - 1;
        int land = ctz(pool);
        emit(&c, land);                                           /* :198 */
        set_land(s, land, army_of(s, land) + 1, RO_NEUTRAL);      /* setupReinforcementNeutralMove + nextPlayerSetupTurn */
        s->phase = RO_SETUP; s->round++; s->cur ^= 1;
        if (s->reinf == 0) {
            derive(s, &d);
            s->phase = RO_REINFORCEMENT; s->reinf = (uint8_t)ro_reinforcement_value(d.owned[s->cur]);
        }
        return RO_OK;
    }
    if (s->phase != RO_REINFORCEMENT) return RO_ERR_ILLEGAL_ACTION;   /* the script only ever starts a turn */
    if (s->cards[me] >= 3) {                                      /* GameHelper::playCards + State::playCards, state.cpp:1091-1117 */
        s->cards[me] = (uint8_t)(s->cards[me] - 3);
        s->card_sets = (uint8_t)(s->card_sets + 1);
        int cs = s->card_sets;
        s->reinf = (uint8_t)(s->reinf + (cs <= 5 ? 2 + 2 * cs : 15 + (cs - 6) * 5));
    }
    while (c.attack_mask > 0 || s->reinf > 0) {
        script_pick_target(&c);
        script_attack_land(&c);
        derive(s, &d);
        c.owned_attack_mask = d.owned_army[me]; c.attack_mask = d.attack_army[me];
    }
    script_fortify(&c);
    end_turn(s);
    return RO_OK;
}


/* RandomPlayer::takeTurn, player/random/random_player.cpp:22-111: uniformly random moves on the State primitives until the turn
   passes.  Every pickRandomMove is one rInt() (k-th lowest set bit, k = rInt() % count), the mobilisation coin is one rFloat(),
   both from the opponent word sequence; dice from the opponent dice stream. */
static int random_pick(script_ctx* c, uint64_t mask)
{
    int k = (int)(az_rng_opp_int(c->seed, c->game, c->ply, c->int_j++) % (uint32_t)popc(mask));
    while (k--) mask &= mask - 1;
    return ctz(mask);
}

int ro_random_turn(ro_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply)
{
    return ro_random_turn_rec(s, r, seed, game, ply, NULL);
}

int ro_random_turn_rec(ro_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply, ro_turn_sink* sink)
{
    if (ro_game_status(s, r) != RO_NOT_ENDED) return RO_ERR_GAME_OVER;
    script_ctx c; memset(&c, 0, sizeof c);
    c.s = s; c.r = r; c.seed = seed; c.game = game; c.ply = ply; c.sink = sink;
    const int me = s->cur;
    while (ro_game_status(s, r) == RO_NOT_ENDED && s->cur == me) {
        derived d; derive(s, &d);
        switch (s->phase) {
        case RO_SETUP: {
            int li = random_pick(&c, d.owned[me]);
            emit(&c, li);                                            /* random_player.cpp:29 */
            s->reinf = (uint8_t)(s->reinf - 2);
            set_land(s, li, army_of(s, li) + 2, me);
            s->phase = RO_SETUP_NEUTRAL;
            break;
        }
        case RO_SETUP_NEUTRAL: {
            int li = random_pick(&c, ALL_LANDS & ~d.owned[0] & ~d.owned[1]);
            emit(&c, li);                                            /* :35 */
            set_land(s, li, army_of(s, li) + 1, RO_NEUTRAL);
            s->phase = RO_SETUP; s->round++; s->cur ^= 1;
            if (s->reinf == 0) {
                derive(s, &d);
                s->phase = RO_REINFORCEMENT; s->reinf = (uint8_t)ro_reinforcement_value(d.owned[s->cur]);
            }
            break;
        }
        case RO_REINFORCEMENT: {
            if (s->cards[me] >= 3) {
                s->cards[me] = (uint8_t)(s->cards[me] - 3);
                s->card_sets = (uint8_t)(s->card_sets + 1);
                int cs = s->card_sets;
                s->reinf = (uint8_t)(s->reinf + (cs <= 5 ? 2 + 2 * cs : 15 + (cs - 6) * 5));
            }
            int li = random_pick(&c, d.owned[me] & ~d.full[me]);
            emit(&c, li);                                            /* :43, after playCards */
            s->reinf = (uint8_t)(s->reinf - 1);                      /* reinforcementMove(1, li) */
            set_land(s, li, army_of(s, li) + 1, me);
            if (s->reinf == 0) goto_attack(s);
            break;
        }
        case RO_ATTACK: {
            int li = random_pick(&c, d.attack_army[me] | SKIP_MASK);
            emit(&c, li);                                            /* :49 */
            if (li == RO_SKIP) s->phase = RO_FORTIFY;                 /* gotoFortify */
            else {
                int from = random_pick(&c, RO_NBR_MASK[li] & d.owned_army[me]);
                script_attack(&c, from, li);
            }
            break;
        }
        case RO_ATTACK_MOBILIZATION: {
            float f = az_rng_unit_float(az_rng_opp_word(seed, game, ply, c.int_j++));
            if (f > 0.5f) {
                int amount = army_of(s, s->mob_from) - 1;
                if (r->min_unit_move < amount) amount = r->min_unit_move;
                emit(&c, s->mob_to);                                 /* :68 */
                set_land(s, s->mob_from, army_of(s, s->mob_from) - amount, me);
                set_land(s, s->mob_to, army_of(s, s->mob_to) + amount, me);
                if (army_of(s, s->mob_from) == 1) goto_attack(s);
            } else { emit(&c, s->mob_from); goto_attack(s); }        /* :73 */
            break;
        }
        default: { /* FORTIFY */
            int to = random_pick(&c, (d.owned[me] & ~d.full[me]) | SKIP_MASK);
            emit(&c, to);                                            /* :82 */
            if (to != RO_SKIP) {
                uint64_t seen = 0; int order[RO_LANDS], n = 0;
                dfs(to, d.owned[me], &seen, order, &n);               /* the owned component that holds `to` */
                uint64_t pool = seen & ~(1ull << to) & d.owned_army[me];
                if (pool) {
                    int from = random_pick(&c, pool);
                    int amount = army_of(s, from) - 1, space = RO_ARMY_MAX - army_of(s, to);
                    if (space < amount) amount = space;
                    int moved = (int)(az_rng_opp_int(seed, game, ply, c.int_j++) % (uint32_t)amount);
                    set_land(s, from, army_of(s, from) - moved, me);
                    set_land(s, to, army_of(s, to) + moved, me);
                }
            }
            end_turn(s);
            break;
        }
        }
    }
    return RO_OK;
}

/* Game::newGame's mirror game, game/game.cpp:170-179: State::invertPlayers (state.cpp:493-516) + setCurrentPlayerTurn */
void ro_invert_players(ro_state* s)
{
    for (int i = 0; i < RO_LANDS; ++i) { int o = owner_of(s, i); if (o < 2) set_land(s, i, army_of(s, i), o ^ 1); }
    uint8_t t = s->cards[0]; s->cards[0] = s->cards[1]; s->cards[1] = t;
}

/* NNInputData(const State&), neural_network/alphazero_nn_data.cpp:165-196, as the 88-byte g++ x86-64 image of the class
   (alphazero_nn_data.h:73-101): land[42] @0, playerIndex @42, round u16 @44, floats @48: reinforcementShare, attackFrequency,
   canDrawCard, phase one-hot x 6, armyShare.  Padding bytes 43, 46, 47 are zero. */
void ro_nn_input(const ro_state* s, uint8_t out[RO_NN_INPUT_BYTES])
{
    derived d; derive(s, &d);
    int c = s->cur, e = c ^ 1;
    float ref = (float)(int8_t)ro_reinforcement_value(d.owned[c]);
    float eref = (float)(int8_t)ro_reinforcement_value(d.owned[e]);
    float att = (float)s->attacks / 8.0f; if (!(att < 1.0f)) att = 1.0f;
    float ta = (float)d.total_army[c], eta = (float)d.total_army[e];
    float f[10];
    f[0] = ref / (ref + eref); f[1] = att; f[2] = s->allow_draw ? 1.0f : 0.0f;
    for (int p = 0; p < 6; ++p) f[3 + p] = s->phase == p ? 1.0f : 0.0f;
    f[9] = ta / (ta + eta);
    memset(out, 0, RO_NN_INPUT_BYTES);
    memcpy(out, s->land, RO_LANDS);
    out[42] = (uint8_t)c;
    out[44] = (uint8_t)(s->round & 0xff); out[45] = (uint8_t)(s->round >> 8);
    memcpy(out + 48, f, sizeof f);
}

/* one training sample as NNTrainDataStorage::saveTrainingSamples writes it (alphazero_nn_data.cpp:115-138):
   int8 playerIndex | NNInputData | float value | float policy[43]; value per updateValues (:51-65) for final status `status` */
void ro_sample_record(const ro_state* s, const float pi[RO_MOVES], int status, uint8_t out[RO_SAMPLE_BYTES])
{
    float z = status == RO_DRAW ? 0.0f : (status == s->cur ? 1.0f : -1.0f);
    out[0] = (uint8_t)s->cur;
    ro_nn_input(s, out + 1);
    memcpy(out + 1 + RO_NN_INPUT_BYTES, &z, 4);
    memcpy(out + 1 + RO_NN_INPUT_BYTES + 4, pi, 4 * RO_MOVES);
}

/* NNOutputData::normalize, neural_network/alphazero_nn_data.cpp:3-27 */
void ro_normalize_policy(float policy[RO_MOVES], uint64_t valid)
{
    float sum = 0.0f;
    for (int i = 0; i < RO_MOVES; ++i) { if ((valid >> i) & 1) sum += policy[i]; else policy[i] = 0.0f; }
    for (int i = 0; i < RO_MOVES; ++i) if (policy[i] > 0.0f) policy[i] /= sum;
}

/* ------------------------------------------------------------------ evaluators */
void ro_eval_pseudo(const ro_state* s, float policy[RO_MOVES], float* value, void* user)
{
    (void)user;
    uint64_t key = az_pn_key(s->land, s->cur, s->round, s->phase);
    for (int i = 0; i < RO_MOVES; ++i) policy[i] = az_pn_policy(key, i);
    *value = az_pn_value(key);
}
void ro_eval_uniform(const ro_state* s, float policy[RO_MOVES], float* value, void* user)
{
    (void)s; (void)user;
    for (int i = 0; i < RO_MOVES; ++i) policy[i] = 1.0f / 43.0f;
    *value = 0.0f;
}

/* ------------------------------------------------------------------ MCTS */
ro_mcts* ro_mcts_new(ro_eval_fn eval, void* user)
{
    ro_mcts* m = (ro_mcts*)calloc(1, sizeof *m);
    m->cap = 256; m->nodes = (ro_node*)malloc(sizeof(ro_node) * (size_t)m->cap);
    m->eval = eval ? eval : ro_eval_pseudo; m->user = user;
    return m;
}
void ro_mcts_free(ro_mcts* m) { if (m) { free(m->nodes); free(m); } }
void ro_mcts_clear(ro_mcts* m) { m->n_nodes = 0; }
int ro_mcts_table_size(const ro_mcts* m) { return m->n_nodes; }
uint64_t ro_mcts_vl_skips(const ro_mcts* m) { return m->vl_skips; }
uint64_t ro_mcts_vl_duplicates(const ro_mcts* m) { return m->vl_duplicates; }

/* StateSimulationsStorage::trimNodes, alphazero_mcts.cpp:229-245 */
void ro_mcts_trim(ro_mcts* m)
{
    int k = 0;
    for (int i = 0; i < m->n_nodes; ++i)
        if (m->nodes[i].visited) { if (k != i) m->nodes[k] = m->nodes[i]; m->nodes[k].visited = 0; k++; }
    m->n_nodes = k;
}

/* StateSimulationsStorage::exist / getStateSimulation, alphazero_mcts.cpp:189-201,217-221:
   full-state equality (State::equalFields, state/state.cpp:111-135; every compared field
   is either primary or derived from primary state) */
static int find_node(const ro_mcts* m, const ro_state* s)
{
    for (int i = 0; i < m->n_nodes; ++i)
        if (memcmp(&m->nodes[i].key, s, sizeof(ro_state)) == 0) return i;
    return -1;
}

/* StateSimulations ctor, alphazero_mcts.cpp:26-42 */
static int expand(ro_mcts* m, const ro_state* s, uint64_t valid, float* value_out)
{
    float policy[RO_MOVES], value = 0.0f;
    m->eval(s, policy, &value, m->user);
    m->evals++;
    ro_normalize_policy(policy, valid);
    if (m->n_nodes == m->cap) { m->cap *= 2; m->nodes = (ro_node*)realloc(m->nodes, sizeof(ro_node) * (size_t)m->cap); }
    ro_node* n = &m->nodes[m->n_nodes++];
    memset(n, 0, sizeof *n);
    n->key = *s; n->valid = valid; n->value = value; n->visited = 1; n->sumN = 0;
    for (int i = 0; i < RO_MOVES; ++i) n->P[i] = ((valid >> i) & 1) ? policy[i] : 0.0f;
    if ((uint64_t)m->n_nodes > m->max_nodes) m->max_nodes = (uint64_t)m->n_nodes;
    *value_out = value;
    return m->n_nodes - 1;
}

/* StateSimulations::getNextBestMoveAndSetVisited, alphazero_mcts.cpp:67-119 with the
   ascending-index iteration order of the contract.  A move nobody has backed up yet (N == 0)
   that exactly one descent is exploring (active_N == 1) is passed over (:91-93) unless no
   other move can be chosen (:108-111); with one descent at a time active_N is 0 here. */
static int select_move(ro_mcts* m, ro_node* n, const ro_rules* r)
{
    n->visited = 1;
    int best = -1, dup = -1; float best_u = -INFINITY, dup_u = -INFINITY;
    for (int i = 0; i < RO_MOVES; ++i) {
        if (!((n->valid >> i) & 1)) continue;
        float P = n->P[i];
        float noiseP = (1 - r->dir_noise_epsi) * P + r->dir_noise_epsi * r->dir_noise_value;
        float v = noiseP * r->cpuct * sqrtf(1.0f + (float)n->sumN);
        float nn = 1.0f + (float)n->N[i];
        float u = n->Q[i] + (v / nn);
        if (u > best_u) {
            if (n->N[i] == 0 && n->active[i] == 1) { m->vl_skips++; if (u > dup_u) { dup_u = u; dup = i; } }
            else { best_u = u; best = i; }
        }
    }
    if (best < 0) { best = dup; m->vl_duplicates++; }
    n->active[best]++;
    return best;
}

/* SimulationValue::addValue + StateSimulations::addValue, alphazero_mcts.cpp:8-21, 56-61 */
static void add_value(ro_node* n, int mv, float v)
{
    if (n->N[mv] == 0) n->Q[mv] = v;
    else n->Q[mv] = ((float)n->N[mv] * n->Q[mv] + v) / (float)(n->N[mv] + 1);
    n->N[mv]++;
    n->active[mv]--;
    n->sumN++;
}

/* AlphaZeroMCTS::search, alphazero_mcts.cpp:322-377 */
static float search(ro_mcts* m, ro_state* s, const ro_rules* r, ro_dice* dice, int* err)
{
    int gs = ro_game_status(s, r);
    if (gs != RO_NOT_ENDED) return gs == RO_DRAW ? 0.0f : (gs == s->cur ? 1.0f : -1.0f);
    uint64_t valid = ro_valid_moves(s, r);
    int idx = find_node(m, s);
    if (idx < 0) { float v; expand(m, s, valid, &v); return v; }
    m->descents++;
    int mv = select_move(m, &m->nodes[idx], r);
    int cur = s->cur;
    if (ro_make_move(s, mv, r, dice) != RO_OK) { *err = 1; return 0.0f; }
    int next = s->cur;                 /* sampled BEFORE recursing: the recursion keeps mutating *s */
    float v = search(m, s, r, dice, err);
    if (next != cur) v = -v;
    add_value(&m->nodes[idx], mv, v);   /* re-fetched: expand may have moved the array */
    return v;
}

/* root statistics via calculateMoveProbability(1.0), alphazero_mcts.cpp:121-148 */
static void root_stats(const ro_mcts* m, const ro_state* root, uint32_t N[RO_MOVES], float Q[RO_MOVES], float P[RO_MOVES],
                       float pi[RO_MOVES], uint32_t* sumN, float* root_value)
{
    const ro_node* n = &m->nodes[find_node(m, root)];
    float sum = 0.0f;
    for (int i = 0; i < RO_MOVES; ++i) {
        int ok = (int)((n->valid >> i) & 1);
        N[i] = ok ? n->N[i] : 0; Q[i] = ok ? n->Q[i] : 0.0f; P[i] = ok ? n->P[i] : 0.0f;
        pi[i] = ok ? (float)pow((double)n->N[i], 1.0 / 1.0f) : 0.0f;
        if (ok) sum += pi[i];
    }
    for (int i = 0; i < RO_MOVES; ++i) pi[i] /= sum;
    *sumN = n->sumN; *root_value = n->value;
}


/* one descent of the lockstep schedule: AlphaZeroMCTS::search (alphazero_mcts.cpp:322-377) unrolled up to the point where the
   reference's thread would block in predictFuture (:350) */
#define RO_PATH_MAX 1024
typedef struct ro_descent {
    int node[RO_PATH_MAX]; uint8_t mv[RO_PATH_MAX], flip[RO_PATH_MAX];
    int len, pending;
    ro_state leaf; uint64_t valid;
} ro_descent;

static void backup(ro_mcts* m, const ro_descent* d, float v)
{
    for (int k = d->len - 1; k >= 0; --k) {
        if (d->flip[k]) v = -v;
        add_value(&m->nodes[d->node[k]], d->mv[k], v);
    }
}

int ro_mcts_search_lockstep(ro_mcts* m, const ro_state* root, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply, int K,
                            uint32_t N[RO_MOVES], float Q[RO_MOVES], float P[RO_MOVES], float pi[RO_MOVES], uint32_t* sumN, float* root_value)
{
    ro_mcts_trim(m);
    if (find_node(m, root) < 0) { float v; expand(m, root, ro_valid_moves(root, r), &v); }
    int count = r->mcts_simulations - (r->mcts_simulations % r->threads_per_mcts);
    if (K < 1 || count % K != 0) return RO_ERR_ILLEGAL_ACTION;
    ro_descent* ds = (ro_descent*)malloc(sizeof(ro_descent) * (size_t)K);
    int rc = RO_OK;
    for (int round = 0; round < count / K && rc == RO_OK; ++round) {
        for (int j = 0; j < K && rc == RO_OK; ++j) {                  /* selection, in thread order */
            ro_descent* d = &ds[j];
            d->len = 0; d->pending = 0;
            ro_dice dice; ro_dice_philox(&dice, seed, game, ply, (uint32_t)(round * K + j));
            ro_state s = *root;
            for (;;) {
                int gs = ro_game_status(&s, r);
                if (gs != RO_NOT_ENDED) { backup(m, d, gs == RO_DRAW ? 0.0f : (gs == s.cur ? 1.0f : -1.0f)); break; }
                uint64_t valid = ro_valid_moves(&s, r);
                int idx = find_node(m, &s);
                if (idx < 0) { d->leaf = s; d->valid = valid; d->pending = 1; break; }
                if (d->len == RO_PATH_MAX) { rc = RO_ERR_ILLEGAL_ACTION; break; }
                m->descents++;
                int mv = select_move(m, &m->nodes[idx], r);
                int cur = s.cur;
                if (ro_make_move(&s, mv, r, &dice) != RO_OK) { rc = RO_ERR_ILLEGAL_ACTION; break; }
                d->node[d->len] = idx; d->mv[d->len] = (uint8_t)mv; d->flip[d->len] = (uint8_t)(s.cur != cur); d->len++;
            }
        }
        for (int j = 0; j < K && rc == RO_OK; ++j) {                  /* completion, in thread order */
            ro_descent* d = &ds[j];
            if (!d->pending) continue;
            float v;
            if (find_node(m, &d->leaf) < 0) expand(m, &d->leaf, d->valid, &v);       /* StateSimulationsStorage::add */
            else {                                                                  /* ... which drops a duplicate (:210-213); the value still counts */
                float policy[RO_MOVES];
                m->eval(&d->leaf, policy, &v, m->user);
                m->evals++;
            }
            backup(m, d, v);
        }
    }
    free(ds);
    if (rc != RO_OK) return rc;
    root_stats(m, root, N, Q, P, pi, sumN, root_value);
    return RO_OK;
}

/* AlphaZeroMCTS::simulate + setRootState, alphazero_mcts.cpp:255-307;
   root statistics via calculateMoveProbability(1.0), :121-148 */
int ro_mcts_search(ro_mcts* m, const ro_state* root, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply,
                   uint32_t N[RO_MOVES], float Q[RO_MOVES], float P[RO_MOVES], float pi[RO_MOVES], uint32_t* sumN, float* root_value)
{
    ro_mcts_trim(m);
    if (find_node(m, root) < 0) { float v; expand(m, root, ro_valid_moves(root, r), &v); }
    int count = r->mcts_simulations - (r->mcts_simulations % r->threads_per_mcts);
    int err = 0;
    for (int i = 0; i < count; ++i) {
        ro_dice dice; ro_dice_philox(&dice, seed, game, ply, (uint32_t)i);
        ro_state copy = *root;
        search(m, &copy, r, &dice, &err);
        if (err) return RO_ERR_ILLEGAL_ACTION;
    }
    root_stats(m, root, N, Q, P, pi, sumN, root_value);
    return RO_OK;
}

/* pickHigestWeightedMove alphazero_mcts.cpp:397-412 / pickRandomWeightedMove :379-395 */
int ro_pick_move(const float pi[RO_MOVES], int sample, uint64_t seed, uint32_t game, uint32_t ply)
{
    if (!sample) {
        float best = 0.0f; int li = RO_NONE;
        for (int i = 0; i < RO_MOVES; ++i) if (pi[i] > best) { best = pi[i]; li = i; }
        return li;
    }
    float sum = 0.0f;
    for (int i = 0; i < RO_MOVES; ++i) sum += pi[i];
    az_u32x4 b = az_rng_block(seed, game, ply, AZ_STREAM_REAL, 0);
    float a = sum * az_rng_unit_float(b.z), it = 0.0f;
    for (int i = 0; i < RO_MOVES; ++i) { it += pi[i]; if (it >= a) return i; }
    return RO_NONE;
}

/* ------------------------------------------------------------------ timing loop (cpu_baseline "port") */
void ro_bench_env(uint64_t n_steps, uint64_t seed, ro_bench_out* out)
{
    ro_rules r; ro_default_rules(&r);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    uint64_t steps = 0, games = 0; uint32_t game = 0;
    while (steps < n_steps) {
        ro_state s; uint32_t ply = 0;
        ro_new_game(&s, seed, game, 0); games++;
        while (steps < n_steps && ro_game_status(&s, &r) == RO_NOT_ENDED) {
            int a = ro_random_action(&s, &r, seed, game, ply);
            ro_dice d; ro_dice_philox(&d, seed, game, ply, AZ_STREAM_REAL);
            ro_make_move(&s, a, &r, &d);
            ply++; steps++;
        }
        game++;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    memset(out, 0, sizeof *out);
    out->steps = steps; out->games = games;
    out->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ CRC32C for oracle/ckpt_oracle.py (checkpoint bundles)
   bit-serial restatement (reflected Castagnoli polynomial 0x82f63b78), 8 bits per byte, no table: independent of the product's
   table-driven az_crc32c */
uint32_t ro_crc32c(const uint8_t* p, uint64_t n, uint32_t crc)
{
    uint32_t c = crc ^ 0xffffffffu;
    for (uint64_t i = 0; i < n; ++i) {
        c ^= p[i];
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82f63b78u & (0u - (c & 1u)));
    }
    return c ^ 0xffffffffu;
}
