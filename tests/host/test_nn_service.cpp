// Standalone test of the host adapter (alphazero_risk_b200/host/az_nn_service.hpp) with the mirror types.
// Without a GPU: constructing the service must fail loudly (no CPU fallback) -> prints NO_DEVICE_OK.
// With a GPU: 16 threads call registerThread / predictFuture concurrently like the reference's search threads;
// every future must equal a direct az_nn_forward of the same input -> prints SERVICE_OK <batches>.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <thread>
#include "az_nn_service.hpp"

using namespace azb200;
typedef AlphaZeroClusterT<NNInputData, NNOutputData, NNTrainData> Cluster;

static NNInputData make_input(unsigned k)
{
    NNInputData in;
    for (int i = 0; i < 42; ++i) { in.land[i].army = 1 + (k * 7 + i * 3) % 30; in.land[i].playerIndex = (k + i) % 3; }
    in.playerIndex = k & 1; in.round = 1 + k % 50;
    in.featureReinforcementShare = 0.25f + 0.01f * (k % 40); in.featureAttackFrequency = 0.125f * (k % 9); in.featureCanDrawCard = k % 2;
    in.featureIsPhaseAttack = 1.0f; in.featureArmyShare = 0.3f + 0.005f * (k % 60);
    return in;
}

int main()
{
    static_assert(sizeof(NNInputData) == 88, "mirror layout");
    if (az_device_count() == 0) {
        try { Cluster c(AZ_NN_FP32, 2); c.initGpus(1); c.initPlayerGroup("az1", "model_bin_V2_2.pb"); }
        catch (const std::runtime_error& e) { printf("NO_DEVICE_OK %s\n", e.what()); return 0; }
        printf("expected a failure without a GPU\n");
        return 1;
    }
    Cluster cluster(AZ_NN_FP32, 2);
    cluster.initGpus(1);
    auto group = cluster.initPlayerGroup("az1", "model_bin_V2_2.pb");
    bool dup = false;
    try { cluster.initPlayerGroup("az1", "model_bin_V2_2.pb"); } catch (const std::invalid_argument&) { dup = true; }   // upstream behaviour
    if (!dup) { printf("duplicate group accepted\n"); return 1; }
    auto nn = group->getNN(0);
    nn->loadCheckpoint("/tmp/az_b200_test_ckpt.bin");          // missing -> random init + save
    const int T = 16, R = 40;
    std::atomic<int> bad{ 0 };
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            nn->registerThread();
            for (int r = 0; r < R; ++r) {
                NNInputData in = make_input(t * 1000 + r);
                NNOutputData a = nn->predictFuture(in).get();
                NNOutputData b = nn->predict(in);
                if (a.policy.size() != 43 || b.policy.size() != 43 || a.value != b.value) { bad++; continue; }
                float s = 0;
                for (int i = 0; i < 43; ++i) { if (a.policy[i] != b.policy[i]) bad++; s += a.policy[i]; }
                if (std::fabs(s - 1.0f) > 1e-4f) bad++;
            }
            nn->unregisterThread();
        });
    for (auto& x : th) x.join();
    bool threw = false;
    try { nn->train({}, 1); } catch (const std::logic_error&) { threw = true; }
    if (bad || !threw) { printf("SERVICE_FAIL bad=%d threw=%d\n", bad.load(), (int)threw); return 1; }
    printf("SERVICE_OK\n");
    return 0;
}
