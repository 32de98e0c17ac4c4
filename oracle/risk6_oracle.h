/*
 * TEST INFRASTRUCTURE — risk6_oracle.h: the six-player extension (BASELINE.json configs[3], SIXPLAYER.md).  PARITY UNPINNED: the
 * reference has no six-player game; this oracle states SIXPLAYER.md and is the only checker of az_env6_*.
 */
#ifndef RISK6_ORACLE_H
#define RISK6_ORACLE_H
#include <stdint.h>
#include "risk_oracle.h"

#define R6_PLAYERS 6
#define R6_SKIP 42
#define R6_NONE 43
#define R6_NOT_ENDED (-1)
#define R6_DRAW (-2)
/* round phases: the reference's numbering (state/state.h:49-57); SETUP_NEUTRAL (1) is never entered */
#define R6_SETUP 0
#define R6_REINFORCEMENT 2
#define R6_ATTACK 3
#define R6_MOBILIZATION 4
#define R6_FORTIFY 5
#define R6_STATE_BYTES 108

/* the whole game state; also the 108-byte image az_env6_export / az_env6_import exchange */
typedef struct r6_state {
    uint8_t army[RO_LANDS];      /* 1..32 */
    uint8_t owner[RO_LANDS];     /* seat 0..5 (no neutral army) */
    uint8_t cards[R6_PLAYERS];   /* simple-card mode: a counter per seat (the reference's STATE_SIMPLE_CARDS default) */
    uint8_t pool[R6_PLAYERS];    /* set-up armies still to place */
    uint16_t round;              /* starts at 1 */
    uint8_t cur, card_sets, reinf, phase, mob_from, mob_to, allow_draw, attacks;
    uint8_t pad[2];
} r6_state;

void r6_new_game(r6_state* s, uint64_t seed, uint32_t game, uint32_t ply);
uint64_t r6_valid_moves(const r6_state* s, const ro_rules* r);
int r6_game_status(const r6_state* s, const ro_rules* r);          /* winner seat 0..5, R6_DRAW, R6_NOT_ENDED */
int r6_make_move(r6_state* s, int action, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply);
int r6_random_action(const r6_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply);
void r6_encode(const r6_state* s, float x[RO_INPUT_FLOATS]);

/* search: the reference's MCTS with the table cleared per search and the six-seat value rule of SIXPLAYER.md; one descent at a
   time; move choice = ro_pick_move (argmax / temperature sampling) as in the two-player game */
typedef void (*r6_eval_fn)(const r6_state* s, float policy[RO_MOVES], float* value, void* user);
typedef struct r6_mcts r6_mcts;
void r6_eval_pseudo(const r6_state* s, float policy[RO_MOVES], float* value, void* user);
r6_mcts* r6_mcts_new(r6_eval_fn eval, void* user);
void r6_mcts_free(r6_mcts* m);
int r6_mcts_table_size(const r6_mcts* m);
int r6_mcts_search(r6_mcts* m, const r6_state* root, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply,
                   uint32_t N[RO_MOVES], float Q[RO_MOVES], float P[RO_MOVES], float pi[RO_MOVES], uint32_t* sumN, float* root_value);
#endif
