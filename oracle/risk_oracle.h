/*
 * risk_oracle.h — TEST INFRASTRUCTURE.  CPU restatement (plain C99) of the reference's
 * self-play hot path: Risk game transition, legal-move mask, input encoding, policy
 * normalisation and the state-keyed MCTS.  It is the checker the CUDA path is compared
 * with; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call it.  The product (alphazero_risk_b200/) never does.
 *
 * Parity status: PINNED for env / moves / encoding / MCTS — the reference has no golden
 * vectors of its own, so this restatement is pinned against the compiled
 * reference (oracle/_ref/libref_oracle.so, built from /root/reference by
 * oracle/ref/build_ref.sh) in tests/test_oracle_vs_ref.py and against the fixtures that
 * library generated (tests/golden/, made by tests/golden/gen_golden.py).
 * UNPINNED for the network arithmetic (TensorFlow absent) — see oracle/nn_oracle.py.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/src/risk_game unless noted).
 */
#ifndef RISK_ORACLE_H
#define RISK_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RO_LANDS 42
#define RO_MOVES 43          /* 42 lands + skip (Land::SKIP_MOVE = LandIndex::Count, land/land.cpp:312) */
#define RO_SKIP 42
#define RO_NONE 43           /* LandIndex::None */
#define RO_ARMY_MAX 32       /* LAND_ARMY_MAX, state/state.h:22 */
#define RO_NEUTRAL 2         /* NEUTRAL_PLAYER, state/state.h:38 */
#define RO_DATA_BYTES 160    /* sizeof(Data), state/state.h:86-105 (g++ 13.3 x86-64) */
#define RO_INPUT_FLOATS 546  /* 7*6*13, alphazero_nn_data.h:66 with INPUT_VECTOR_TYPE_2 */

enum { RO_SETUP = 0, RO_SETUP_NEUTRAL = 1, RO_REINFORCEMENT = 2, RO_ATTACK = 3, RO_ATTACK_MOBILIZATION = 4, RO_FORTIFY = 5 };
enum { RO_NOT_ENDED = -1, RO_DRAW = -2 };
enum { RO_OK = 0, RO_ERR_ILLEGAL_ACTION = -1, RO_ERR_GAME_OVER = -2, RO_ERR_BAD_STATE = -3, RO_ERR_TABLE_FULL = -4 };

/* primary (non-derivable) game state; everything else in the reference's Data is a pure
   function of it (State::consistencyCheck, state/state.cpp:1209-1429) */
typedef struct ro_state {
    uint8_t land[RO_LANDS];      /* LandArmy byte: army (6 low bits) | owner << 6, state/state.h:24-34 */
    uint8_t cards[2];            /* PlayerStatus::playerCards (STATE_SIMPLE_CARDS), state/state.h:68-69 */
    uint16_t round;              /* Data::round, starts at 1 */
    int8_t cur;                  /* Data::currentPlayerTurn */
    uint8_t card_sets;           /* Data::cardSetsPlayed */
    uint8_t reinf;               /* Data::reinforcements */
    uint8_t phase;               /* Data::roundPhase */
    uint8_t mob_from, mob_to;    /* Data::attackMobilizationFrom/To, RO_NONE when unset */
    uint8_t allow_draw;          /* Data::playerAllowedDrawCard */
    uint8_t attacks;             /* Data::attacksDuringTurn */
} ro_state;

/* runtime options of settings.h:40-62 that change the path's results */
typedef struct ro_rules {
    int allow_yield;             /* ALLOW_YIELD = true */
    int limit_reinforcement;     /* LIMIT_REINFORCEMENT_MOVES = true */
    int limit_attack;            /* LIMIT_ATTACK_MOVES = false */
    int max_game_rounds;         /* MAX_GAME_ROUNDS = 58 */
    int min_unit_move;           /* MIN_UNIT_MOVE = 3 */
    int mcts_simulations;        /* MCTS_SIMULATIONS = 32 */
    int threads_per_mcts;        /* THREADS_PER_MCTS = 2 (only used for sims - sims % T) */
    float cpuct;                 /* HP_EXPLORATION = 1.1 */
    float dir_noise_value;       /* DIR_NOISE_VALUE = 0.3 */
    float dir_noise_epsi;        /* DIR_NOISE_EPSI = 0.25 */
    int temperature_threshold;   /* TEMPERATURE_TRESHOLD = 43 */
} ro_rules;

void ro_default_rules(ro_rules* r);

/* dice source: the include/az_philox.h contract or an explicit tape */
typedef struct ro_dice {
    int use_tape;
    uint64_t seed; uint32_t game, ply, sim, j;
    const int32_t* tape; int tape_len, tape_pos;
} ro_dice;

void ro_dice_philox(ro_dice* d, uint64_t seed, uint32_t game, uint32_t ply, uint32_t sim);
void ro_dice_tape(ro_dice* d, const int32_t* tape, int n);

/* map tables (land/land.cpp:246-297, land/land_set.cpp:12-33, land/land_index.h:5-10) */
extern const uint64_t RO_NBR_MASK[RO_LANDS];
extern const int8_t RO_NBR_LIST[RO_LANDS][6];
extern const uint64_t RO_CONTINENT_MASK[6];
extern const int RO_CONTINENT_BONUS[6];

/* state <-> the reference's 160-byte Data image (padding bytes written as zero) */
void ro_export_data(const ro_state* s, uint8_t data[RO_DATA_BYTES]);
int ro_import_data(ro_state* s, const uint8_t data[RO_DATA_BYTES]);
/* 1 for bytes of Data that carry information, 0 for compiler padding */
void ro_data_byte_mask(uint8_t mask[RO_DATA_BYTES]);

void ro_new_game(ro_state* s, uint64_t seed, uint32_t game, uint32_t ply);       /* State::newGame */
void ro_new_game_tape(ro_state* s, const int32_t* draws42);
uint64_t ro_valid_moves(const ro_state* s, const ro_rules* r);                   /* UtilityNN::getValidMoves */
int ro_game_status(const ro_state* s, const ro_rules* r);                        /* State::gameStatus */
int ro_reinforcement_value(uint64_t owned);                                      /* State::calculateReinforcementValue */
int ro_make_move(ro_state* s, int action, const ro_rules* r, ro_dice* dice);     /* UtilityNN::makeMove */
int ro_random_action(const ro_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply);
void ro_encode(const ro_state* s, float x[RO_INPUT_FLOATS]);                     /* NNInputData + setInStateTensor */
void ro_normalize_policy(float policy[RO_MOVES], uint64_t valid);                /* NNOutputData::normalize */
/* ---- scripted opponent (ScriptPlayer, player/script/script_player.cpp) ---- */
typedef struct ro_script { int8_t set, to, from; uint8_t from_army; } ro_script;   /* attackingLandSet / landAttackTo / landAttackFrom / attackFromArmy */
void ro_script_init(ro_script* sp);
int ro_script_turn(ro_state* s, ro_script* sp, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply);  /* ScriptPlayer::takeTurn */
int ro_random_turn(ro_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply);                 /* RandomPlayer::takeTurn */
/* Player::addTrainingSample (player/base/player.cpp:9-17) inside those turns: every call site of script_player.cpp / random_player.cpp
   hands over (state before the move, move); the sink keeps at most `cap` of them and counts all in `n`. */
typedef struct ro_turn_sink { ro_state* states; uint8_t* moves; int cap; int n; } ro_turn_sink;
int ro_script_turn_rec(ro_state* s, ro_script* sp, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply, ro_turn_sink* sink);
int ro_random_turn_rec(ro_state* s, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply, ro_turn_sink* sink);
void ro_invert_players(ro_state* s);                                             /* State::invertPlayers */
#define RO_NN_INPUT_BYTES 88   /* sizeof(NNInputData), alphazero_nn_data.h:73-101 */
#define RO_SAMPLE_BYTES 265    /* 1 + 88 + 4 + 43 * 4, alphazero_nn_data.cpp:115-138 */
void ro_nn_input(const ro_state* s, uint8_t out[RO_NN_INPUT_BYTES]);             /* NNInputData(const State&) image */
void ro_sample_record(const ro_state* s, const float pi[RO_MOVES], int status, uint8_t out[RO_SAMPLE_BYTES]);

/* ---- MCTS (alphazero_mcts.cpp) ---- */
typedef void (*ro_eval_fn)(const ro_state* s, float policy[RO_MOVES], float* value, void* user);
void ro_eval_pseudo(const ro_state* s, float policy[RO_MOVES], float* value, void* user);
void ro_eval_uniform(const ro_state* s, float policy[RO_MOVES], float* value, void* user);

typedef struct ro_node {
    ro_state key;
    uint64_t valid;
    float value;
    uint32_t sumN;
    uint8_t visited;
    float Q[RO_MOVES], P[RO_MOVES];
    uint32_t N[RO_MOVES];
    uint8_t active[RO_MOVES];    /* SimulationValue::active_N, alphazero_mcts.h:16-28: descents that selected the move and have not backed up yet */
} ro_node;

typedef struct ro_mcts {
    ro_node* nodes;
    int n_nodes, cap;
    ro_eval_fn eval; void* user;
    uint64_t evals, max_nodes, descents;
    uint64_t vl_skips, vl_duplicates;   /* times the active_N rule passed a move over / had to request a duplicate (alphazero_mcts.cpp:91-111) */
} ro_mcts;

ro_mcts* ro_mcts_new(ro_eval_fn eval, void* user);
void ro_mcts_free(ro_mcts* m);
void ro_mcts_clear(ro_mcts* m);                                                  /* StateSimulationsStorage::clearNodes */
void ro_mcts_trim(ro_mcts* m);                                                   /* StateSimulationsStorage::trimNodes */
int ro_mcts_table_size(const ro_mcts* m);
uint64_t ro_mcts_vl_skips(const ro_mcts* m);
uint64_t ro_mcts_vl_duplicates(const ro_mcts* m);
/* AlphaZeroMCTS::simulate with T = 1 semantics; outputs root statistics (zeros for illegal moves) */
int ro_mcts_search(ro_mcts* m, const ro_state* root, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply,
                   uint32_t N[RO_MOVES], float Q[RO_MOVES], float P[RO_MOVES], float pi[RO_MOVES], uint32_t* sumN, float* root_value);
/* AlphaZeroMCTS::simulate with K search threads in the LOCKSTEP schedule (one of the interleavings the reference's locks allow,
   pinned against the reference's own threads taking turns in oracle/ref/ref_shim.cpp ref_mcts_search_lockstep): simulations run in
   rounds of K; within a round descent j selects after descents 0..j-1 (seeing their active_N marks — the "virtual loss" rule,
   alphazero_mcts.cpp:91-107), a descent that ends in a terminal state backs up at once, the others are evaluated together and then
   expanded + backed up in order j = 0..K-1.  K = 1 is ro_mcts_search. */
int ro_mcts_search_lockstep(ro_mcts* m, const ro_state* root, const ro_rules* r, uint64_t seed, uint32_t game, uint32_t ply, int K,
                            uint32_t N[RO_MOVES], float Q[RO_MOVES], float P[RO_MOVES], float pi[RO_MOVES], uint32_t* sumN, float* root_value);
int ro_pick_move(const float pi[RO_MOVES], int sample, uint64_t seed, uint32_t game, uint32_t ply);

/* ---- bounded CPU timing loops for bench.py's cpu_baseline "port" ---- */
typedef struct ro_bench_out { uint64_t steps, games, sims, evals, moves; double seconds; } ro_bench_out;
void ro_bench_env(uint64_t n_steps, uint64_t seed, ro_bench_out* out);

uint32_t ro_crc32c(const uint8_t* p, uint64_t n, uint32_t crc);                  /* CRC32C, for oracle/ckpt_oracle.py */

#ifdef __cplusplus
}
#endif
#endif
