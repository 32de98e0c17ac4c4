#pragma once
/*
 * TEST INFRASTRUCTURE — TensorFlow-free overlay for
 * /root/reference/src/risk_game/player/alpha_zero/neural_network/alphazero_gpu_cluster.h
 * (and, transitively, alphazero_nn.h, whose 13 tensorflow includes at
 * alphazero_nn.h:13-25 cannot be satisfied here: TensorFlow is an un-vendored,
 * un-pinned dependency).
 *
 * Same class and method names as the reference façade (alphazero_gpu_cluster.h:14-58),
 * so the UNMODIFIED alphazero_mcts.cpp compiles against it; every prediction is
 * routed to a caller-supplied evaluator instead of session->Run
 * (alphazero_nn.cpp:247-248, :339-340).
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

/* evaluator: fills policy[43] (un-masked, as the network's softmax would) and *value */
typedef void (*RefEvalFn)(const NNInputData* in, float* policy43, float* value, void* user);

class AlphaZeroNNId
{
public:
	RefEvalFn fn = nullptr;
	void* user = nullptr;
	uint64_t evals = 0;
	/* lockstep schedule of THREADS_PER_MCTS search threads (ref_mcts_search_lockstep): called by a search
	   thread when it reaches predictFuture — where the reference's thread blocks until the batch runs
	   (alphazero_nn.cpp:282-320) — and returns when that thread may go on */
	void (*park)(void*) = nullptr;
	void* park_user = nullptr;

	AlphaZeroNNId() {}
	AlphaZeroNNId(RefEvalFn f, void* u) : fn(f), user(u) {}

	void loadCheckpoint(std::string) {}
	void saveCheckpoint(std::string) {}
	void train(const std::vector<NNTrainData>&, int) {}

	void registerThread() {}
	void unregisterThread() {}

	NNOutputData predict(const NNInputData& state)
	{
		NNOutputData out;
		out.policy.resize(TF_OUTPUT_POLICY_TENSOR_SIZE);
		fn(&state, out.policy.data(), &out.value, user);
		evals++;
		return out;
	}

	std::future<NNOutputData> predictFuture(const NNInputData& state)
	{
		if (park) park(park_user);
		std::promise<NNOutputData> p;
		p.set_value(predict(state));
		return p.get_future();
	}
};

class AlphaZeroNNGroup
{
	std::string name;
	std::vector<std::shared_ptr<AlphaZeroNNId>> neuralNetworkIds;
public:
	AlphaZeroNNGroup(std::string n) : name(n) {}
	void add(std::shared_ptr<AlphaZeroNNId> instance) { neuralNetworkIds.push_back(instance); }
	void loadCheckpoint(std::string) {}
	void saveCheckpoint(std::string) {}
	void train(const std::vector<NNTrainData>&, int) {}
	int size() { return (int)neuralNetworkIds.size(); }
	std::shared_ptr<AlphaZeroNNId> getNN(int i) { return neuralNetworkIds[i]; }
};
