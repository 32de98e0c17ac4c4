#pragma once
/*
 * TEST INFRASTRUCTURE — second overlay for
 * /root/reference/src/risk_game/player/alpha_zero/neural_network/alphazero_gpu_cluster.h:
 * the reference's NN façade class names bound to the B200 adapter
 * (alphazero_risk_b200/host/az_nn_service.hpp over libaz_b200.so).  With this header in place the
 * reference's UNMODIFIED AlphaZeroMCTS / AlphaZeroPlayer / Game / ScriptPlayer compile and run against the
 * CUDA network — it is exactly the binding INTEGRATION.md tells a reference maintainer to add, used here to
 * test the drop-in (oracle/_ref/libref_gpusvc.so, tests/test_dropin_gpu.py).
 */
#include <cstdio>
#include <string>
#include <vector>
#include <memory>
#include <mutex>
#include <thread>
#include <future>
#include <unordered_map>
#include <condition_variable>
#include <algorithm>
#include <filesystem>
#include <chrono>

#include "alphazero_nn_data.h"
#include "az_nn_service.hpp"

typedef azb200::AlphaZeroNNIdT<NNInputData, NNOutputData, NNTrainData> AlphaZeroNNId;
typedef azb200::AlphaZeroNNGroupT<NNInputData, NNOutputData, NNTrainData> AlphaZeroNNGroup;
typedef azb200::AlphaZeroClusterT<NNInputData, NNOutputData, NNTrainData> AlphaZeroCluster;
