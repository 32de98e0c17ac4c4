// az_cluster.hpp — host-side C++ adapter for the reference's multi-GPU model: ONE process, one network copy and one
// set of games per GPU, as AlphaZeroCluster::initGpus / initPlayerGroup set it up
// (/root/reference/src/risk_game/player/alpha_zero/neural_network/alphazero_gpu_cluster.cpp:147-193), on top of the az_dist_* entry
// points of libaz_b200.so:
//
//   azb200::DeviceCluster cluster(settings, SETTINGS.GPUS);         // one DevicePlay (network + arena) per GPU, one NCCL communicator
//   cluster.loadCheckpoint(path);                                   // GPU 0 restores, the others receive the weights over NVLink
//   ... az_nn_train(cluster.network(0), ...) ...
//   cluster.broadcastWeights();                                     // AlphaZeroNNGroup::train's hand-off (alphazero_gpu_cluster.cpp:221-231,
//                                                                   // a temporary checkpoint FILE in the reference) = az_dist_broadcast_weights
//   GameResults gr = cluster.playGames<GameResults>(games);         // the match sharded over the GPUs in whole mirror pairs, one host
//                                                                   // thread per GPU; GameResults::add over the shards (game.cpp:298-309)
//                                                                   // = az_dist_gather_results
//
// Games never migrate and nothing is exchanged while a match runs.  Errors: std::runtime_error with az_last_error().
#pragma once

#include <exception>
#include <memory>
#include <thread>

#include "az_play.hpp"

namespace azb200 {

class DeviceCluster {
public:
    DeviceCluster(const PlaySettings& s, int gpus) : st(s)
    {
        if (az_device_count() == 0) throw std::runtime_error("azb200::DeviceCluster: no CUDA device (libaz_b200 has no CPU fallback)");
        if (gpus < 1 || gpus > az_device_count()) throw std::invalid_argument("azb200::DeviceCluster: gpus must be in 1 .. az_device_count()");
        for (int g = 0; g < gpus; ++g) {
            PlaySettings sg = s;
            sg.device = g;                                           // the reference's gpu index
            members.emplace_back(new DevicePlay(sg));                // every copy starts from the same random init (same seed)
        }
        if (az_dist_init(gpus, nullptr, &dist) != AZ_OK) throw std::runtime_error(std::string("az_dist_init: ") + az_last_error());
    }
    ~DeviceCluster() { members.clear(); if (dist) az_dist_destroy(dist); }
    DeviceCluster(const DeviceCluster&) = delete;
    DeviceCluster& operator=(const DeviceCluster&) = delete;

    int size() const { return (int)members.size(); }                 // AlphaZeroNNGroup::size
    az_nn* network(int gpu) { return members.at((size_t)gpu)->network(); }   // AlphaZeroNNGroup::getNN(i)
    DevicePlay& member(int gpu) { return *members.at((size_t)gpu); }

    // AlphaZeroNNGroup::loadCheckpoint: the reference restores the file on every GPU; here GPU 0 reads it once and the weights travel
    // over NVLink
    void loadCheckpoint(const std::string& path) { members[0]->loadCheckpoint(path); broadcastWeights(0); }

    void broadcastWeights(int root = 0)
    {
        std::vector<az_nn*> nets;
        for (auto& m : members) nets.push_back(m->network());
        ck(az_dist_broadcast_weights(dist, nets.data(), (int)nets.size(), root), "az_dist_broadcast_weights");
    }

    // GameGroup::playGames over the whole cluster: 2 * floor(games / 2) games in mirror pairs, pairs dealt to the GPUs in contiguous
    // blocks; every shard uses its own range of global slot ids, so the match does not depend on which GPU plays which block
    az_arena_results play(int games)
    {
        const int g = size(), pairs = games / 2;
        if (pairs < g) throw std::invalid_argument("azb200::DeviceCluster: fewer mirror pairs than GPUs");
        std::vector<az_arena_results> local((size_t)g);
        std::vector<std::exception_ptr> err((size_t)g);
        std::vector<std::thread> threads;
        uint32_t first = 0;
        for (int i = 0; i < g; ++i) {
            const int my_pairs = pairs / g + (i < pairs % g ? 1 : 0);
            members[(size_t)i]->setFirstGameId(st.first_game_id + first);
            first += (uint32_t)(st.slots > 0 ? st.slots : my_pairs);
            threads.emplace_back([this, i, my_pairs, &local, &err] {
                try { local[(size_t)i] = members[(size_t)i]->play(2 * my_pairs); } catch (...) { err[(size_t)i] = std::current_exception(); }
            });
        }
        for (auto& t : threads) t.join();
        for (auto& e : err) if (e) std::rethrow_exception(e);
        az_arena_results total;
        ck(az_dist_gather_results(dist, local.data(), &total), "az_dist_gather_results");
        per_gpu = local; last = total;
        return total;
    }

    template <class GameResultsT>
    GameResultsT playGames(int games)
    {
        const az_arena_results r = play(games);
        GameResultsT gr;
        gr.count = (int)r.count; gr.draw = (int)r.draw;
        for (int i = 0; i < 2; ++i) { gr.players[i].win = (int)r.win[i]; gr.players[i].winAndStartedGame = (int)r.win_and_started[i]; }
        return gr;
    }

    az_arena_results last{};                       // totals of the last match
    std::vector<az_arena_results> per_gpu;         // its shards

private:
    PlaySettings st;
    std::vector<std::unique_ptr<DevicePlay>> members;
    az_dist* dist = nullptr;
    static void ck(int rc, const char* what) { if (rc != AZ_OK) throw std::runtime_error(std::string(what) + ": " + az_last_error()); }
};

}  // namespace azb200
