// az_play.hpp — host-side C++ adapter for SURVEY.md §8b seam 2 / §8f N1: the reference's
//     GameResults GameGroup::playGames(std::shared_ptr<PlayerGroup> pg1, std::shared_ptr<PlayerGroup> pg2, int games)
// (/root/reference/src/risk_game/game/game.cpp:277-312) for the pairings the reference sets up — executePlay's AlphaZeroPlayerGroup
// against a ScriptPlayerGroup (src/alphazero_risk.cpp:4-47), the trainer's benchmark against a RandomPlayerGroup and its comparison
// match of the new against the old model (alphazero_trainer.cpp:121-166) — played entirely on the device through the az_arena_*
// entry points of libaz_b200.so instead of one std::thread per game.
//
//   azb200::DevicePlay play(settings);                       // settings: the SETTINGS fields the path reads, see PlaySettings
//   play.loadCheckpoint(path);                               // AlphaZeroNNGroup::loadCheckpoint semantics (missing file => random init + save)
//   GameResults gr = play.playGames<GameResults>(SETTINGS.COMPARE_GAMES);            // vs ScriptPlayer (setOpponent: RandomPlayer)
//   play.setOpponentNetwork(old.network());                  // player index 1 = a second AlphaZeroPlayer on `old`'s network
//   GameResults cmp = play.playGames<GameResults>(SETTINGS.COMPARE_GAMES);           // isModelImproved(cmp) as in the trainer
//
// GameResultsT is the reference's own class (game/game.h:17-29: count, draw, players[i].win, players[i].winAndStartedGame); the
// template keeps this header compilable without the reference tree (tests/host/test_play.cpp uses a mirror type).
// Errors: std::runtime_error with az_last_error(), like az_nn_service.hpp.  No CPU fallback.
#pragma once

#include <algorithm>
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
    uint32_t first_game_id = 0;           // global id of slot 0: shards of one match on several GPUs use disjoint id ranges (az_cluster.hpp)
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

    // AlphaZeroNN::loadCheckpoint, alphazero_nn.cpp:189-204: restore the TensorFlow checkpoint bundle <path>.index +
    // <path>.data-00000-of-00001, or (missing) keep the random init and save it (TF_OP_SAVE, :207-214)
    void loadCheckpoint(const std::string& path)
    {
        if (std::ifstream(path + ".index", std::ios::binary).good()) ck(az_nn_load_checkpoint(nn, path.c_str()), "az_nn_load_checkpoint");
        else ck(az_nn_save_checkpoint(nn, path.c_str()), "az_nn_save_checkpoint");
        release();                                            // the searcher re-packs the weights when it is rebuilt
    }
    az_nn* network() { return nn; }
    // global id of slot 0 (az_cluster.hpp gives every GPU's shard of a match its own id range)
    void setFirstGameId(uint32_t id) { if (id != st.first_game_id) { st.first_game_id = id; release(); } }

    // player index 1: AZ_OPPONENT_SCRIPT (default, ScriptPlayerGroup) or AZ_OPPONENT_RANDOM (RandomPlayerGroup, the trainer's benchmark)
    void setOpponent(int kind)
    {
        if (kind != AZ_OPPONENT_SCRIPT && kind != AZ_OPPONENT_RANDOM) throw std::invalid_argument("azb200::DevicePlay: unknown opponent kind");
        opponent = kind; opponent_nn = nullptr; release();
    }
    // player index 1 = a second AlphaZeroPlayer searching with `other` (not owned; same number of blocks not required): the
    // trainer's comparison match GameGroup::playGames(trainAZPG, generateAZPG, COMPARE_GAMES), alphazero_trainer.cpp:147-166
    void setOpponentNetwork(az_nn* other)
    {
        if (!other || other == nn) throw std::invalid_argument("azb200::DevicePlay: the opponent needs its own network");
        opponent = AZ_OPPONENT_ALPHAZERO; opponent_nn = other; release();
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
        last = r;
        return r;
    }

    az_arena_results last{};          // counters of the last match (moves, simulations, ticks)

private:
    PlaySettings st;
    az_nn* nn = nullptr; az_env* env = nullptr; az_mcts* mcts = nullptr; az_arena* arena = nullptr;
    az_nn* opponent_nn = nullptr; az_mcts* opponent_mcts = nullptr;
    int opponent = AZ_OPPONENT_SCRIPT;
    int n_slots = 0;

    static void ck(int rc, const char* what) { if (rc != AZ_OK) throw std::runtime_error(std::string(what) + ": " + az_last_error()); }

    void build(int slots)
    {
        az_rules r; az_default_rules(&r);
        r.mcts_simulations = st.MCTS_SIMULATIONS; r.threads_per_mcts = st.THREADS_PER_MCTS; r.cpuct = st.HP_EXPLORATION;
        r.dir_noise_value = st.DIR_NOISE_VALUE; r.dir_noise_epsi = st.DIR_NOISE_EPSI; r.allow_yield = st.ALLOW_YIELD;
        r.limit_reinforcement = st.LIMIT_REINFORCEMENT_MOVES; r.limit_attack = st.LIMIT_ATTACK_MOVES;
        r.max_game_rounds = st.MAX_GAME_ROUNDS; r.min_unit_move = st.MIN_UNIT_MOVE;
        ck(az_env_create(slots, &r, st.device, st.first_game_id, &env), "az_env_create");
        ck(az_mcts_create(env, nn, AZ_EVAL_NN, st.precision, &mcts), "az_mcts_create");
        if (opponent == AZ_OPPONENT_ALPHAZERO) {
            ck(az_mcts_create(env, opponent_nn, AZ_EVAL_NN, st.precision, &opponent_mcts), "az_mcts_create");
            ck(az_arena_create_versus(mcts, opponent_mcts, st.MIRROR_GAMES ? 1 : 0, &arena), "az_arena_create_versus");
        } else ck(az_arena_create(mcts, opponent, st.MIRROR_GAMES ? 1 : 0, &arena), "az_arena_create");
        n_slots = slots;
    }
    void release()
    {
        if (arena) az_arena_destroy(arena);
        if (opponent_mcts) az_mcts_destroy(opponent_mcts);
        if (mcts) az_mcts_destroy(mcts);
        if (env) az_env_destroy(env);
        arena = nullptr; mcts = nullptr; opponent_mcts = nullptr; env = nullptr; n_slots = 0;
    }
};

// The data generation of AlphaZeroTrainer::trainOnGeneratedData (alphazero_trainer.cpp:242-268): DATA_GAMES_SS Script-vs-Script and
// DATA_GAMES_SR Script-vs-Random games whose players carry a shared NNTrainDataStorage (Player::addTrainingSample, one-hot policies;
// values from gameFinished -> updateValues) — played in lockstep on the device.  play() returns packed AZ_SAMPLE_BYTES records, the
// layout NNTrainDataStorage::saveTrainingSamples writes, ready for az_nn_train / az_samples_write_file:
//
//   azb200::DeviceDataGames gen(settings, 4096);
//   auto recs = gen.play(SETTINGS.DATA_GAMES_SS, /*second_is_random*/ false);
//   auto more = gen.play(SETTINGS.DATA_GAMES_SR, true);          recs.insert(recs.end(), more.begin(), more.end());
//   az_nn_train(nn, recs.data(), recs.size() / AZ_SAMPLE_BYTES, 3, SETTINGS.BATCH_SIZE, seed, nullptr, nullptr, nullptr);
class DeviceDataGames {
public:
    DeviceDataGames(const PlaySettings& s, int slots, int max_samples_per_game = 4096) : st(s), n(slots)
    {
        if (az_device_count() == 0) throw std::runtime_error("azb200::DeviceDataGames: no CUDA device (libaz_b200 has no CPU fallback)");
        if (slots <= 0) throw std::invalid_argument("azb200::DeviceDataGames: slots must be positive");
        az_rules r; az_default_rules(&r);
        r.allow_yield = st.ALLOW_YIELD; r.limit_reinforcement = st.LIMIT_REINFORCEMENT_MOVES; r.limit_attack = st.LIMIT_ATTACK_MOVES;
        r.max_game_rounds = st.MAX_GAME_ROUNDS; r.min_unit_move = st.MIN_UNIT_MOVE;
        ck(az_env_create(n, &r, st.device, 0, &env), "az_env_create");
        ck(az_env_record_turns(env, (size_t)n * (size_t)max_samples_per_game, max_samples_per_game), "az_env_record_turns");
    }
    ~DeviceDataGames() { if (env) az_env_destroy(env); }
    DeviceDataGames(const DeviceDataGames&) = delete;
    DeviceDataGames& operator=(const DeviceDataGames&) = delete;

    // at least `games` games (whole batches of `slots`), player index 0 = ScriptPlayer, player index 1 = ScriptPlayer or RandomPlayer
    std::vector<uint8_t> play(int games, bool second_is_random)
    {
        std::vector<uint8_t> out;
        std::vector<uint32_t> script((size_t)n * 2);
        std::vector<int8_t> status((size_t)n);
        games_played = 0; dropped = 0;
        while (games_played < games) {
            ck(az_env_reset(env, st.seed + 0x9E3779B97F4A7C15ull * (uint64_t)(++batches), nullptr), "az_env_reset");
            std::fill(script.begin(), script.end(), (uint32_t)AZ_SCRIPT_INIT);      // new ScriptPlayer objects per Game, alphazero_trainer.cpp:246-249
            bool running = true;
            for (int turn = 0; running && turn < 4096; ++turn) {
                ck(az_env_play_turn(env, AZ_OPPONENT_SCRIPT, second_is_random ? AZ_OPPONENT_RANDOM : AZ_OPPONENT_SCRIPT, script.data(),
                                    status.data(), nullptr), "az_env_play_turn");
                running = false;
                for (int8_t v : status) if (v == AZ_STATUS_RUNNING) { running = true; break; }
            }
            size_t have = 0; uint64_t d = 0;
            ck(az_env_turn_samples(env, nullptr, 0, &have, &d, nullptr), "az_env_turn_samples");
            const size_t at = out.size();
            out.resize(at + have * (size_t)AZ_SAMPLE_BYTES);
            if (have) {                                   // an empty queue was drained (dropped count reset) by the query itself
                ck(az_env_turn_samples(env, out.data() + at, have, &have, &d, nullptr), "az_env_turn_samples");
                out.resize(at + have * (size_t)AZ_SAMPLE_BYTES);
            }
            dropped += d;
            games_played += n;
        }
        return out;
    }

    int games_played = 0;             // of the last play()
    uint64_t dropped = 0;             // samples of games that overflowed their staging area (0 with the default 4096 per game)

private:
    PlaySettings st;
    int n;
    az_env* env = nullptr;
    uint64_t batches = 0;
    static void ck(int rc, const char* what) { if (rc != AZ_OK) throw std::runtime_error(std::string(what) + ": " + az_last_error()); }
};

}  // namespace azb200
