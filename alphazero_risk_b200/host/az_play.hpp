// az_play.hpp — host-side C++ adapter for SURVEY.md §8b seam 2 / §8f N1: the reference's
//     GameResults GameGroup::playGames(std::shared_ptr<PlayerGroup> pg1, std::shared_ptr<PlayerGroup> pg2, int games)
// (/root/reference/src/risk_game/game/game.cpp:277-312) for the pairing executePlay sets up by default
// (src/alphazero_risk.cpp:4-47: an AlphaZeroPlayerGroup against a ScriptPlayerGroup), played entirely on the device through the
// az_arena_* entry points of libaz_b200.so instead of one std::thread per game.
//
//   azb200::DevicePlay play(settings);                       // settings: the SETTINGS fields the path reads, see PlaySettings
//   play.loadCheckpoint(path);                               // AlphaZeroNNGroup::loadCheckpoint semantics (missing file => random init + save)
//   GameResults gr = play.playGames<GameResults>(SETTINGS.COMPARE_GAMES);
//
// GameResultsT is the reference's own class (game/game.h:17-29: count, draw, players[i].win, players[i].winAndStartedGame); the
// template keeps this header compilable without the reference tree (tests/host/test_play.cpp uses a mirror type).
// Errors: std::runtime_error with az_last_error(), like az_nn_service.hpp.  No CPU fallback.
#pragma once

#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "az_b200.h"

namespace azb200 {

// the SETTINGS.* fields (src/settings.h:22-81) that change what the match computes, under their reference names
struct PlaySettings {
    int MCTS_SIMULATIONS = 32;            // --mcts
    int THREADS_PER_MCTS = 2;             // -t     (sims - sims % t simulations per move, alphazero_mcts.cpp:265)
    float HP_EXPLORATION = 1.1f;          // --hp
    float DIR_NOISE_VALUE = 0.3f;         // --dnv
    float DIR_NOISE_EPSI = 0.25f;         // --dne
    bool ALLOW_YIELD = true, LIMIT_REINFORCEMENT_MOVES = true, LIMIT_ATTACK_MOVES = false, MIRROR_GAMES = true;
    int MAX_GAME_ROUNDS = 58, MIN_UNIT_MOVE = 3;
    int BLOCKS = 5;                       // CMake BLOCKS / the GraphDef in use (model_bin_V2_5.pb = 5)
    int device = 0;                       // the reference's gpu index
    int slots = 0;                        // concurrent games; 0 = one slot per claimed pair (the reference: gpus x gpu-games threads)
    int precision = AZ_NN_BF16;
    uint64_t seed = 0x5EED0001ull;        // Philox contract seed (the reference seeds from random_device)
};

class DevicePlay {
public:
    explicit DevicePlay(const PlaySettings& s) : st(s)
    {
        if (az_device_count() == 0) throw std::runtime_error("azb200::DevicePlay: no CUDA device (libaz_b200 has no CPU fallback)");
        ck(az_nn_create(st.BLOCKS, st.device, &nn), "az_nn_create");
        ck(az_nn_init_random(nn, st.seed), "az_nn_init_random");
    }
    ~DevicePlay() { release(); if (nn) az_nn_destroy(nn); }
    DevicePlay(const DevicePlay&) = delete;
    DevicePlay& operator=(const DevicePlay&) = delete;

    // AlphaZeroNN::loadCheckpoint, alphazero_nn.cpp:189-204: restore, or (missing file) keep the random init and save it.
    // File format: the flat fp32 blob of az_nn_export_blob (TF checkpoint import is SURVEY §8f N4, not built).
    void loadCheckpoint(const std::string& path)
    {
        const size_t count = az_nn_num_params(nn);
        std::vector<float> blob(count);
        std::ifstream in(path, std::ios::binary);
        if (in && in.read(reinterpret_cast<char*>(blob.data()), (std::streamsize)(count * sizeof(float)))) {
            ck(az_nn_import_blob(nn, blob.data(), count), "az_nn_import_blob");
        } else {
            ck(az_nn_export_blob(nn, blob.data(), count), "az_nn_export_blob");
            std::ofstream out(path, std::ios::binary);
            out.write(reinterpret_cast<const char*>(blob.data()), (std::streamsize)(count * sizeof(float)));
        }
    }

    // GameGroup::playGames(alphaZeroGroup, scriptGroup, games): 2 * floor(games / 2) games in mirror pairs
    template <class GameResultsT>
    GameResultsT playGames(int games)
    {
        const az_arena_results r = play(games);
        GameResultsT gr;
        gr.count = (int)r.count; gr.draw = (int)r.draw;
        for (int i = 0; i < 2; ++i) { gr.players[i].win = (int)r.win[i]; gr.players[i].winAndStartedGame = (int)r.win_and_started[i]; }
        return gr;
    }

    az_arena_results play(int games)
    {
        if (games < 2) throw std::invalid_argument("azb200::DevicePlay: games are played in pairs, need at least 2");
        const int want = st.slots > 0 ? st.slots : games / 2;
        if (!arena || want != n_slots) { release(); build(want); }
        az_arena_results r;
        ck(az_arena_play(arena, (uint64_t)games, st.seed, &r, nullptr), "az_arena_play");
        if (r.errors) throw std::runtime_error("azb200::DevicePlay: MCTS node pool overflow");
        return r;
    }

private:
    PlaySettings st;
    az_nn* nn = nullptr; az_env* env = nullptr; az_mcts* mcts = nullptr; az_arena* arena = nullptr;
    int n_slots = 0;

    static void ck(int rc, const char* what) { if (rc != AZ_OK) throw std::runtime_error(std::string(what) + ": " + az_last_error()); }

    void build(int slots)
    {
        az_rules r; az_default_rules(&r);
        r.mcts_simulations = st.MCTS_SIMULATIONS; r.threads_per_mcts = st.THREADS_PER_MCTS; r.cpuct = st.HP_EXPLORATION;
        r.dir_noise_value = st.DIR_NOISE_VALUE; r.dir_noise_epsi = st.DIR_NOISE_EPSI; r.allow_yield = st.ALLOW_YIELD;
        r.limit_reinforcement = st.LIMIT_REINFORCEMENT_MOVES; r.limit_attack = st.LIMIT_ATTACK_MOVES;
        r.max_game_rounds = st.MAX_GAME_ROUNDS; r.min_unit_move = st.MIN_UNIT_MOVE;
        ck(az_env_create(slots, &r, st.device, 0, &env), "az_env_create");
        ck(az_mcts_create(env, nn, AZ_EVAL_NN, st.precision, &mcts), "az_mcts_create");
        ck(az_arena_create(mcts, AZ_OPPONENT_SCRIPT, st.MIRROR_GAMES ? 1 : 0, &arena), "az_arena_create");
        n_slots = slots;
    }
    void release()
    {
        if (arena) az_arena_destroy(arena);
        if (mcts) az_mcts_destroy(mcts);
        if (env) az_env_destroy(env);
        arena = nullptr; mcts = nullptr; env = nullptr; n_slots = 0;
    }
};

}  // namespace azb200
