#pragma once
/*
 * TEST INFRASTRUCTURE — overlay for
 * /root/reference/src/risk_game/player/alpha_zero/neural_network/alphazero_gpu_cluster.h used when the reference's OWN main
 * (src/alphazero_risk.cpp: executePlay / executeTrain / ..., settings.h CLI) is compiled against the B200 adapters
 * (oracle/ref/build_ref_main.sh -> oracle/_ref/AlphaZero_risk).  Same class names as the reference header, bound to
 * alphazero_risk_b200/host/az_nn_service.hpp; AlphaZeroNN (the TensorFlow class, only named by executeAnalysis) still comes from
 * the reference's alphazero_nn.h compiled against oracle/ref/tf_stub.
 */
#include "alphazero_nn.h"
#include "overlay_gpu_cluster_b200.h"

class AlphaZeroGPU {            // only the static helper executeAnalysis names (alphazero_gpu_cluster.cpp:100-103)
public:
    static std::string getDevicePath(int index) { return "/device:GPU:" + std::to_string(index); }
};
