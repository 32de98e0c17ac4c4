// az_arena.cuh — device-side bookkeeping of the -m play arena, shared by az_env.cu (k_arena_advance) and az_mcts.cu (host loop)
#pragma once

#include <stdint.h>
#include <cuda_runtime.h>

#include "az_b200.h"

enum { ARENA_COUNT = 0, ARENA_DRAW = 1, ARENA_WIN0 = 2, ARENA_WIN1 = 3, ARENA_WAS0 = 4, ARENA_WAS1 = 5, ARENA_OPP_TURNS = 6,
       ARENA_CLAIMED = 7, ARENA_ACTIVE = 8, ARENA_TOMOVE0 = 9, ARENA_TOMOVE1 = 10, ARENA_N = 12 };

struct ArenaDev {
    int n;
    uint32_t first_game;
    uint64_t seed;
    uint32_t* state;            // env SoA [16][n]
    uint32_t* start_state;      // [16][n]  Game::previousStartState (game/game.h)
    uint32_t* script;           // [n][2]   packed ScriptPlayer members per side
    uint8_t* player_start;      // [n]      Game::playerStart
    uint8_t* fresh;             // [n]      slot has not started a game yet
    uint8_t* active;            // [n]      slot still plays
    uint8_t* last_mover;        // [n]      side that moved last in the running game (0xff = nobody yet)
    uint8_t* extra_trim;        // [n]      the MCTS handle's pending trimNodes counter
    uint8_t* extra_trim_opp;    // [n]      the same for the second searcher (AZ_OPPONENT_ALPHAZERO), else NULL
    uint8_t* ended;             // [n]      game that ended in this advance call: 0 none, 1 / 2 = winner 0 / 1, 3 = draw (sample flush)
    uint8_t* to_move;           // [n]      side whose searcher moves next in this slot (AZ_OPPONENT_ALPHAZERO; 0xff = slot is done)
    unsigned long long* res;    // [ARENA_N]
    unsigned long long total_games;
    int opponent, mirror;
};

int az_launch_arena_advance(const ArenaDev& a, const az_rules* rules, cudaStream_t s);   // az_env.cu
