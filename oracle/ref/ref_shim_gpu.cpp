/*
 * TEST INFRASTRUCTURE — C-ABI shim of oracle/_ref/libref_gpusvc.so: the UNMODIFIED reference search and game
 * loop (AlphaZeroMCTS, AlphaZeroPlayer, Game/GameGroup, ScriptPlayer) served by the B200 network through the
 * host adapter alphazero_risk_b200/host/az_nn_service.hpp.  This is the drop-in of SURVEY.md §8b seam 1 and
 * the shape of BASELINE config 1 (`-m play`: AlphaZero vs ScriptPlayer, executePlay, src/alphazero_risk.cpp:4-47).
 */
#include <cstring>
#include <cstdint>
#include <chrono>
#include <thread>
#include <vector>

#include "risk_game/player/alpha_zero/alphazero_player.h"
#include "risk_game/player/script/script_player.h"
#include "az_philox.h"

#define REF_API extern "C" __attribute__((visibility("default")))

static thread_local char g_err[512];
REF_API const char* refgpu_last_error() { return g_err; }

struct GpuRef {
    std::shared_ptr<AlphaZeroCluster> cluster;
    std::shared_ptr<AlphaZeroNNGroup> group;
};

REF_API void* refgpu_new(int blocks, int precision, uint64_t seed)
{
    try {
        GpuRef* g = new GpuRef();
        g->cluster = std::make_shared<AlphaZeroCluster>(precision, blocks);
        g->cluster->initGpus(1);
        g->group = g->cluster->initPlayerGroup("az1", "model_bin_V2_" + std::to_string(blocks) + ".pb");
        g->group->getNN(0)->service().initRandom(seed);
        return g;
    } catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return nullptr; }
}
REF_API void refgpu_free(void* h) { delete (GpuRef*)h; }

static void rng_philox(uint64_t seed, uint32_t game, uint32_t ply, uint32_t sim)
{
    RefRngCtx& c = ref_rng_ctx();
    c.mode = REF_RNG_PHILOX; c.seed = seed; c.game = game; c.ply = ply; c.sim = sim; c.die_j = 0; c.deal_i = 0;
}

/*
 * n_games self-play games, ONE HOST THREAD PER GAME sharing one AlphaZeroNNId exactly like the reference's
 * thread-per-game arena (game/game.cpp:289-296): each thread registers with the service and its leaf
 * evaluations are batched across threads by predictFuture.  Search = setRootState + sims x search(copy) on the
 * thread itself so every simulation gets its (game, ply, sim) dice stream.  Records, per game and move,
 * the root visit counts and the chosen move (moves_cap moves at most).
 */
REF_API int refgpu_selfplay_threads(void* h, int n_games, uint32_t first_game, uint64_t seed, int sims, int moves_cap,
                                    uint32_t* visits /*[n_games][moves_cap][43]*/, uint8_t* moves /*[n_games][moves_cap]*/,
                                    int32_t* n_moves /*[n_games]*/, uint8_t* final_data /*[n_games][160]*/, uint64_t* batch_stats /*[2]*/)
{
    GpuRef* g = (GpuRef*)h;
    std::shared_ptr<AlphaZeroNNId> nn = g->group->getNN(0);
    SETTINGS.MCTS_SIMULATIONS = sims; SETTINGS.THREADS_PER_MCTS = 1;
    std::vector<std::thread> th;
    std::vector<int> fail(n_games, 0);
    for (int t = 0; t < n_games; ++t)
        th.emplace_back([&, t]() {
            try {
                nn->registerThread();
                AlphaZeroMCTS mcts;
                State root; root.setLog(false);
                const uint32_t game = first_game + (uint32_t)t;
                rng_philox(seed, game, 0, AZ_STREAM_DEAL);
                root.newGame();
                int ply = 0;
                while (root.gameStatus() == State::NOT_ENDED && ply < moves_cap) {
                    rng_philox(seed, game, (uint32_t)ply, 0);
                    mcts.setRootState(root, nn);
                    for (int i = 0; i < sims; ++i) {
                        rng_philox(seed, game, (uint32_t)ply, (uint32_t)i);
                        State copy = root; copy.setLog(false);
                        mcts.search(copy, nn);
                    }
                    std::shared_ptr<StateSimulations> ss = mcts.getStorage()->getStateSimulation(root);
                    std::vector<float> policy = ss->calculateMoveProbability(1.0f);
                    for (int i = 0; i < ALL_MOVES; ++i) {
                        LandIndex li = Utility::i2li((uint8_t)i);
                        visits[((size_t)t * moves_cap + ply) * ALL_MOVES + i] = ss->moveValues.contains(li) ? ss->moveValues.at(li).N : 0;
                    }
                    rng_philox(seed, game, (uint32_t)ply, AZ_STREAM_REAL);
                    LandIndex li = root.getRound() > SETTINGS.TEMPERATURE_TRESHOLD ? mcts.pickHigestWeightedMove(policy) : mcts.pickRandomWeightedMove(policy);
                    moves[(size_t)t * moves_cap + ply] = Utility::li2i(li);
                    rng_philox(seed, game, (uint32_t)ply, AZ_STREAM_REAL);
                    UtilityNN::makeMove(root, li);
                    ply++;
                }
                n_moves[t] = ply;
                memcpy(final_data + (size_t)t * sizeof(Data), &root.getData(), sizeof(Data));
                nn->unregisterThread();
            } catch (const std::exception& e) { fail[t] = 1; snprintf(g_err, sizeof g_err, "%s", e.what()); nn->unregisterThread(); }
        });
    for (auto& x : th) x.join();
    for (int f : fail) if (f) return -1;
    return 0;
}

/* BASELINE config 1 shape: executePlay (src/alphazero_risk.cpp:4-47) — AlphaZeroPlayerGroup (random-init net) vs
   ScriptPlayerGroup through the reference's own GameGroup::playGames; returns wall seconds and the tally */
REF_API int refgpu_play_vs_script(void* h, int games, int mcts_sims, int threads_per_mcts, int concurrent_games,
                                  int32_t* out5 /*count, draw, az wins, script wins, az wins-and-started*/, double* seconds)
{
    try {
        GpuRef* g = (GpuRef*)h;
        SETTINGS.MCTS_SIMULATIONS = mcts_sims; SETTINGS.THREADS_PER_MCTS = threads_per_mcts;
        SETTINGS.NUMBER_OF_GPUS = 1; SETTINGS.NUMBER_OF_CONCURENT_GAMES_PER_GPU = concurrent_games;
        RefRngCtx& c = ref_rng_ctx(); c.mode = REF_RNG_ENGINE;
        std::shared_ptr<PlayerGroup> g1(new AlphaZeroPlayerGroup(g->group));
        std::shared_ptr<PlayerGroup> g2(new ScriptPlayerGroup(concurrent_games));
        auto t0 = std::chrono::steady_clock::now();
        GameResults gr = GameGroup::playGames(g1, g2, games);
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out5[0] = gr.count; out5[1] = gr.draw; out5[2] = gr.players[0].win; out5[3] = gr.players[1].win; out5[4] = gr.players[0].winAndStartedGame;
        return 0;
    } catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}
