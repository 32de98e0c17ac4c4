// Standalone test of the az_dist_* entry points and of alphazero_risk_b200/host/az_cluster.hpp: the reference's multi-GPU model,
// one process driving every visible GPU (two are enough; with one GPU the world has one rank and the same calls must still work).
// Without a GPU: az_dist_init must fail loudly -> NO_DEVICE_OK.  With GPUs -> DIST_OK.
//   1. weight broadcast (alphazero_gpu_cluster.cpp:221-231): copies with different weights agree bit for bit afterwards, from
//      either root, also after a training step moved the root's device copy ahead of its host copy, and the receivers' forward
//      (re-packed bf16 tiles) equals the root's;
//   2. statistics gather (game.cpp:298-309): per-rank listing = the local vectors, totals = their sum; az_counters of sharded
//      rollouts add up to the single-GPU rollout of the same global game ids;
//   3. the rank model (one communicator per thread, az_dist_unique_id / az_dist_init_rank) gives the same answers;
//   4. azb200::DeviceCluster plays a sharded match: totals = sum of the shards, shards reproducible.
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "az_cluster.hpp"

struct PlayerGameResult { int win = 0; int winAndStartedGame = 0; };
struct GameResults { int count = 0; int draw = 0; std::vector<PlayerGameResult> players = std::vector<PlayerGameResult>(2); };

#define CK(call) do { if ((call) != AZ_OK) { printf("FAIL %s: %s\n", #call, az_last_error()); return 1; } } while (0)
#define EXPECT(cond) do { if (!(cond)) { printf("FAIL line %d: %s\n", __LINE__, #cond); return 1; } } while (0)

static std::vector<float> blob_of(az_nn* nn)
{
    std::vector<float> b(az_nn_num_params(nn));
    if (az_nn_export_blob(nn, b.data(), b.size()) != AZ_OK) b.clear();
    return b;
}

int main()
{
    const int have = az_device_count();
    if (have == 0) {
        az_dist* d = nullptr;
        if (az_dist_init(1, nullptr, &d) == AZ_OK) { printf("expected a failure without a GPU\n"); return 1; }
        printf("NO_DEVICE_OK %s\n", az_last_error());
        return 0;
    }
    const int G = have >= 2 ? 2 : 1;
    int ver = 0;
    CK(az_dist_nccl_version(&ver));
    az_dist* dist = nullptr;
    CK(az_dist_init(G, nullptr, &dist));
    EXPECT(az_dist_world_size(dist) == G && az_dist_local_count(dist) == G && az_dist_rank(dist, G - 1) == G - 1);

    // ---- 1. weights
    std::vector<az_nn*> nn((size_t)G);
    for (int g = 0; g < G; ++g) { CK(az_nn_create(2, g, &nn[(size_t)g])); CK(az_nn_init_random(nn[(size_t)g], 100 + (uint64_t)g)); }
    std::vector<float> root = blob_of(nn[0]);
    if (G > 1) EXPECT(blob_of(nn[1]) != root);
    CK(az_dist_broadcast_weights(dist, nn.data(), G, 0));
    for (int g = 0; g < G; ++g) EXPECT(blob_of(nn[(size_t)g]) == root);
    std::vector<float> x(8 * AZ_INPUT_FLOATS);
    for (size_t i = 0; i < x.size(); ++i) x[i] = (float)((i * 2654435761u) % 1000u) / 1000.0f;
    std::vector<float> p0(8 * AZ_MOVES), v0(8), p1(8 * AZ_MOVES), v1(8);
    CK(az_nn_forward(nn[0], x.data(), 8, p0.data(), v0.data(), AZ_NN_BF16, nullptr));
    CK(az_nn_forward(nn[(size_t)G - 1], x.data(), 8, p1.data(), v1.data(), AZ_NN_BF16, nullptr));
    EXPECT(p0 == p1 && v0 == v1);
    {   // the root trains: its device copy moves ahead of the host copy; the hand-off must ship the trained weights
        std::vector<float> tp(8 * AZ_MOVES, 1.0f / AZ_MOVES), tv(8, 0.5f);
        float lp = 0, lv = 0;
        az_nn* trainer = nn[(size_t)G - 1];
        CK(az_nn_train_step(trainer, x.data(), tp.data(), tv.data(), 8, &lp, &lv, nullptr));
        CK(az_dist_broadcast_weights(dist, nn.data(), G, G - 1));
        std::vector<float> trained = blob_of(trainer);
        EXPECT(trained != root);
        for (int g = 0; g < G; ++g) EXPECT(blob_of(nn[(size_t)g]) == trained);
        CK(az_nn_forward(nn[0], x.data(), 8, p0.data(), v0.data(), AZ_NN_BF16, nullptr));
        CK(az_nn_forward(trainer, x.data(), 8, p1.data(), v1.data(), AZ_NN_BF16, nullptr));
        EXPECT(p0 == p1 && v0 == v1);
    }

    // ---- 2. statistics
    {
        std::vector<uint64_t> local((size_t)G * 5), sum(5), per((size_t)G * 5);
        for (size_t i = 0; i < local.size(); ++i) local[i] = 1000003ull * (i + 1) + (1ull << 40);
        CK(az_dist_gather_stats(dist, local.data(), 5, sum.data(), per.data()));
        EXPECT(per == local);
        for (int k = 0; k < 5; ++k) { uint64_t s = 0; for (int g = 0; g < G; ++g) s += local[(size_t)g * 5 + k]; EXPECT(sum[(size_t)k] == s); }
        CK(az_dist_barrier(dist));
        // sharded rollouts: G shards of 512 games with contiguous global ids == one env of G * 512 games
        const int per_gpu = 512, steps = 300;
        az_rules r; az_default_rules(&r);
        std::vector<az_env*> env((size_t)G);
        std::vector<az_counters> cl((size_t)G);
        for (int g = 0; g < G; ++g) {
            CK(az_env_create(per_gpu, &r, g, (uint32_t)(g * per_gpu), &env[(size_t)g]));
            CK(az_env_reset(env[(size_t)g], 0x5EED0001ull, nullptr));
            CK(az_env_rollout(env[(size_t)g], steps, nullptr));
        }
        for (int g = 0; g < G; ++g) CK(az_env_counters(env[(size_t)g], &cl[(size_t)g], 0, nullptr));
        az_counters total, whole;
        CK(az_dist_gather_counters(dist, cl.data(), &total));
        az_env* all = nullptr;
        CK(az_env_create(G * per_gpu, &r, 0, 0, &all));
        CK(az_env_reset(all, 0x5EED0001ull, nullptr));
        CK(az_env_rollout(all, steps, nullptr));
        CK(az_env_counters(all, &whole, 0, nullptr));
        EXPECT(total.steps == whole.steps && total.games == whole.games && total.wins[0] == whole.wins[0] && total.wins[1] == whole.wins[1] &&
               total.draws == whole.draws && total.steps == (uint64_t)G * per_gpu * steps && total.games > 0);
        for (az_env* e : env) az_env_destroy(e);
        az_env_destroy(all);
    }
    CK(az_dist_destroy(dist));

    // ---- 3. rank model: one thread per rank, each with its own communicator handle
    {
        uint8_t id[AZ_DIST_ID_BYTES];
        CK(az_dist_unique_id(id));
        std::vector<int> rc((size_t)G, -1);
        std::vector<uint64_t> sums((size_t)G * 3);
        std::vector<std::vector<float>> got((size_t)G);
        for (int g = 0; g < G; ++g) CK(az_nn_init_random(nn[(size_t)g], 500 + (uint64_t)g));
        const std::vector<float> want = blob_of(nn[(size_t)G - 1]);
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g)
            th.emplace_back([&, g] {
                az_dist* d = nullptr;
                int r = az_dist_init_rank(G, g, id, g, &d);
                if (r == AZ_OK) { az_nn* mine[1] = { nn[(size_t)g] }; r = az_dist_broadcast_weights(d, mine, 1, G - 1); }
                uint64_t loc[3] = { (uint64_t)g + 1, 10ull * (uint64_t)(g + 1), 7 };
                if (r == AZ_OK) r = az_dist_gather_stats(d, loc, 3, &sums[(size_t)g * 3], nullptr);
                if (r == AZ_OK) r = az_dist_barrier(d);
                if (r != AZ_OK) printf("rank %d: %s\n", g, az_last_error());
                az_dist_destroy(d);
                rc[(size_t)g] = r;
            });
        for (auto& t : th) t.join();
        for (int g = 0; g < G; ++g) {
            EXPECT(rc[(size_t)g] == AZ_OK);
            EXPECT(blob_of(nn[(size_t)g]) == want);
            EXPECT(sums[(size_t)g * 3] == (uint64_t)G * (G + 1) / 2 && sums[(size_t)g * 3 + 1] == 10ull * G * (G + 1) / 2 && sums[(size_t)g * 3 + 2] == 7ull * G);
        }
    }
    for (az_nn* n : nn) az_nn_destroy(n);

    // ---- 4. the cluster adapter: a match sharded over the GPUs
    {
        azb200::PlaySettings s;
        s.MCTS_SIMULATIONS = 8; s.THREADS_PER_MCTS = 2; s.BLOCKS = 1;
        azb200::DeviceCluster cluster(s, G);
        CK(az_nn_init_random(cluster.network(G - 1), 999));          // make the copies differ, then hand GPU 0's weights to everyone
        cluster.broadcastWeights(0);
        GameResults a = cluster.playGames<GameResults>(24);
        const az_arena_results tot = cluster.last;
        uint64_t cnt = 0, mv = 0, w0 = 0;
        for (const az_arena_results& r : cluster.per_gpu) { cnt += r.count; mv += r.az_moves; w0 += r.win[0]; }
        EXPECT(a.count == 24 && cnt == 24 && tot.count == 24 && tot.az_moves == mv && tot.win[0] == w0 && tot.errors == 0);
        EXPECT(a.draw + a.players[0].win + a.players[1].win == 24);
        GameResults b = cluster.playGames<GameResults>(24);          // same weights, same ids: reproducible
        EXPECT(b.count == a.count && b.draw == a.draw && b.players[0].win == a.players[0].win && cluster.last.az_moves == tot.az_moves);
        printf("cluster match: %d games on %d GPU(s), az %d / script %d / draw %d, %llu AlphaZero moves\n", a.count, G, a.players[0].win,
               a.players[1].win, a.draw, (unsigned long long)tot.az_moves);
    }
    printf("DIST_OK gpus=%d nccl=%d\n", G, ver);
    return 0;
}
