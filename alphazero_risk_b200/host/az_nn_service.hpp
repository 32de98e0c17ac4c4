// az_nn_service.hpp — host-side C++ adapter that presents the reference's NN façade
// (AlphaZeroNNId / AlphaZeroNNGroup / AlphaZeroCluster,
//  /root/reference/src/risk_game/player/alpha_zero/neural_network/alphazero_gpu_cluster.h:14-111)
// on top of the C ABI of libaz_b200.so, so the reference's UNCHANGED AlphaZeroMCTS / AlphaZeroPlayer
// can be served by the B200 network (SURVEY.md §8b seam 1).
//
// Same names, argument meaning and threading contract as the reference:
//   predict(in)          synchronous single evaluation (root of a search), alphazero_nn.cpp:333-349
//   predictFuture(in)    called concurrently by games x THREADS_PER_MCTS threads; requests are queued and
//                        evaluated as ONE batch by a consumer thread per network, alphazero_nn.cpp:236-289,
//                        alphazero_gpu_cluster.cpp:13-71; queue capacity max(1, registeredThreads / 2),
//                        alphazero_nn.cpp:291-309; the call blocks while the queue is full
//   registerThread / unregisterThread   batch-size hint, as above
//   loadCheckpoint(path) `path` is a TensorFlow checkpoint prefix: restores <path>.index + <path>.data-00000-of-00001 when the
//                        index exists, else random init (the graph's "init" op) + saveCheckpoint, alphazero_nn.cpp:189-204
//   saveCheckpoint(path) creates the directory and writes the bundle the graph's Saver writes (variables, moving statistics,
//                        Adam slots, beta powers), alphazero_nn.cpp:207-214
//   train(data, epochs)  AlphaZeroNN::train, alphazero_nn.cpp:351-410: `epochs` shuffled passes in whole batches of
//                        SETTINGS.BATCH_SIZE (512; setBatchSize) through az_nn_train
// Errors: the reference aborts through TF_CHECK_OK; here every ABI failure throws std::runtime_error with
// az_last_error().
//
// The adapter is a template over the caller's NNInputData / NNOutputData so it can be compiled both
// inside the reference tree (with the reference's own types) and standalone (with the mirror types below).
#pragma once

#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <future>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "az_b200.h"

namespace azb200 {

inline void check(int rc, const char* what)
{
    if (rc != AZ_OK) throw std::runtime_error(std::string(what) + ": " + az_last_error());
}

// byte view of the reference's NNInputData (alphazero_nn_data.h:81-105, INPUT_VECTOR_TYPE_2, 88 bytes):
// LandArmy land[42] @0, uint8 playerIndex @42, uint16 round @44, nine floats @48..80, featureArmyShare @84
struct InputView {
    const uint8_t* p;
    float f(int off) const { float v; std::memcpy(&v, p + off, 4); return v; }
};

// NNInputData -> [7][6][13] fp32 tensor; restates setInStateTensor, alphazero_nn.cpp:31-67 (channel indices
// alphazero_nn_data.h:13-39)
inline void encode_input(const void* nn_input_data, float* x546)
{
    InputView in{ static_cast<const uint8_t*>(nn_input_data) };
    const int cur = in.p[42], enemy = cur == 0 ? 1 : 0;
    const float reinf = in.f(48), attack = in.f(52), draw = in.f(56);
    const float phase[6] = { in.f(60), in.f(64), in.f(68), in.f(72), in.f(76), in.f(80) };
    const float army_share = in.f(84);
    for (int i = 0; i < 42; ++i) {
        const int army = in.p[i] & 63, owner = in.p[i] >> 6;
        const float fa = float(army) / 32.0f;
        float* o = x546 + i * 13;
        o[0] = owner == cur ? fa : 0.0f; o[1] = owner == enemy ? fa : 0.0f; o[2] = owner == 2 ? fa : 0.0f;
        o[3] = army_share; o[4] = reinf; o[5] = attack; o[6] = draw;
        for (int k = 0; k < 6; ++k) o[7 + k] = phase[k];
    }
}

template <class In, class Out>
class NNService {
    static_assert(sizeof(In) == 88, "NNInputData must be the reference's 88-byte INPUT_VECTOR_TYPE_2 layout");

    struct Pending { In in; std::promise<Out> promise; explicit Pending(const In& i) : in(i) {} };

    az_nn* nn_ = nullptr;
    int precision_;
    std::mutex lock_, gpu_lock_;
    std::condition_variable cv_full_, cv_empty_;
    std::vector<Pending> accepting_, processing_;
    int registered_ = 0, queue_size_ = 1;
    int batch_size_ = 512;              // SETTINGS.BATCH_SIZE, settings.h:74
    uint64_t train_seed_ = 1;           // shuffle stream of the next train() call
    bool running_ = true;
    std::thread consumer_;
    std::vector<float> x_, pol_, val_;

    bool queue_full() const { return (int)accepting_.size() >= queue_size_; }

    Out make_out(const float* pol, float v) const { Out o; o.policy.assign(pol, pol + AZ_MOVES); o.value = v; return o; }

    void run_batch(std::vector<Pending>& batch)
    {
        const int n = (int)batch.size();
        x_.resize((size_t)n * AZ_INPUT_FLOATS); pol_.resize((size_t)n * AZ_MOVES); val_.resize(n);
        for (int i = 0; i < n; ++i) encode_input(&batch[i].in, x_.data() + (size_t)i * AZ_INPUT_FLOATS);
        try {
            std::lock_guard<std::mutex> g(gpu_lock_);         // one Run per GPU at a time, alphazero_gpu_cluster.cpp:33
            check(az_nn_forward(nn_, x_.data(), n, pol_.data(), val_.data(), precision_, nullptr), "az_nn_forward");
        } catch (...) {
            // a failed batch (e.g. out of memory in az_nn_reserve) must reach every MCTS thread blocked in future.get() instead of
            // unwinding out of the consumer thread (std::terminate, the waiters hang): the reference aborts with a message at the
            // call site (TF_CHECK_OK, alphazero_nn.cpp:268); here predictFuture().get() rethrows az_last_error's text
            for (int i = 0; i < n; ++i) batch[i].promise.set_exception(std::current_exception());
            batch.clear();
            return;
        }
        for (int i = 0; i < n; ++i) batch[i].promise.set_value(make_out(pol_.data() + (size_t)i * AZ_MOVES, val_[i]));
        batch.clear();
    }

    void consume()
    {
        for (;;) {
            {
                std::unique_lock<std::mutex> ul(lock_);
                cv_full_.wait(ul, [this] { return !running_ || (!accepting_.empty() && queue_full()); });
                if (!running_ && accepting_.empty()) return;
                accepting_.swap(processing_);
            }
            cv_empty_.notify_all();
            run_batch(processing_);
        }
    }

public:
    NNService(int blocks, int device, int precision) : precision_(precision)
    {
        check(az_nn_create(blocks, device, &nn_), "az_nn_create");
        consumer_ = std::thread([this] { consume(); });
    }
    ~NNService()
    {
        { std::lock_guard<std::mutex> g(lock_); running_ = false; }
        cv_full_.notify_all();
        if (consumer_.joinable()) consumer_.join();
        az_nn_destroy(nn_);
    }
    NNService(const NNService&) = delete;

    az_nn* handle() { return nn_; }
    uint64_t batches = 0, samples = 0;

    void initRandom(uint64_t seed) { std::lock_guard<std::mutex> g(gpu_lock_); check(az_nn_init_random(nn_, seed), "az_nn_init_random"); }

    void saveCheckpoint(const std::string& path)
    {
        const size_t cut = path.find_last_of("/\\");
        if (cut != std::string::npos) std::filesystem::create_directories(path.substr(0, cut + 1));
        std::lock_guard<std::mutex> g(gpu_lock_);
        check(az_nn_save_checkpoint(nn_, path.c_str()), "az_nn_save_checkpoint");
    }
    void loadCheckpoint(const std::string& path)
    {
        if (std::filesystem::exists(path + ".index")) {
            std::lock_guard<std::mutex> g(gpu_lock_);
            check(az_nn_load_checkpoint(nn_, path.c_str()), "az_nn_load_checkpoint");
        } else {
            printf("Checkpoint '%s' not found initialized random weights\n", path.c_str());
            initRandom(1234);
            saveCheckpoint(path);
        }
    }

    // AlphaZeroNN::train (alphazero_nn.cpp:351-410).  TrainData = the reference's NNTrainData {int8 playerIndex; NNInputData in;
    // NNOutputData out}: packed into the 265-byte records of the sample file layout (alphazero_nn_data.cpp:115-138)
    void setBatchSize(int b) { batch_size_ = b; }
    template <class TrainData> void train(const std::vector<TrainData>& data, int epochs)
    {
        if ((int)data.size() < batch_size_) return;             // batchCount == 0: the reference's loops do nothing
        std::vector<uint8_t> rec(data.size() * (size_t)AZ_SAMPLE_BYTES);
        for (size_t i = 0; i < data.size(); ++i) {
            uint8_t* r = rec.data() + i * (size_t)AZ_SAMPLE_BYTES;
            r[0] = (uint8_t)data[i].playerIndex;
            static_assert(sizeof(In) == 88, "NNInputData must be the reference's 88-byte INPUT_VECTOR_TYPE_2 layout");
            std::memcpy(r + 1, &data[i].in, 88);
            std::memcpy(r + 89, &data[i].out.value, 4);
            std::memcpy(r + 93, data[i].out.policy.data(), 43 * sizeof(float));
        }
        std::vector<float> lp((size_t)epochs), lv((size_t)epochs);
        printf("Started training\n");
        std::lock_guard<std::mutex> g(gpu_lock_);
        check(az_nn_train(nn_, rec.data(), data.size(), epochs, batch_size_, train_seed_++, lp.data(), lv.data(), nullptr), "az_nn_train");
        for (int e = 0; e < epochs; ++e) printf("EPOCH %d\nLoss Policy / Value: %f / %f\n", e, lp[(size_t)e], lv[(size_t)e]);
    }

    // contractions of the training step: AZ_NN_FP32 (default, parity path) or AZ_NN_BF16 (tcgen05 GEMMs, ~7x faster at batch 512)
    void setTrainPrecision(int precision) { std::lock_guard<std::mutex> g(gpu_lock_); check(az_nn_train_precision(nn_, precision), "az_nn_train_precision"); }

    // variables + optimizer state of `other` into this network (the group hand-off after train, alphazero_gpu_cluster.cpp:221-231)
    void copyStateFrom(NNService& other)
    {
        if (&other == this) return;
        std::scoped_lock g(gpu_lock_, other.gpu_lock_);
        check(az_nn_copy_state(nn_, other.nn_, nullptr), "az_nn_copy_state");
    }

    void registerThread()
    {
        { std::lock_guard<std::mutex> g(lock_); registered_++; queue_size_ = registered_ / 2 > 1 ? registered_ / 2 : 1; }
        cv_empty_.notify_all();
    }
    void unregisterThread()
    {
        { std::lock_guard<std::mutex> g(lock_); registered_--; queue_size_ = registered_ / 2 > 1 ? registered_ / 2 : 1; }
        cv_full_.notify_all();
    }

    std::future<Out> predictFuture(const In& state)
    {
        std::future<Out> f;
        bool full;
        {
            std::unique_lock<std::mutex> ul(lock_);
            cv_empty_.wait(ul, [this] { return !queue_full(); });
            accepting_.emplace_back(state);
            f = accepting_.back().promise.get_future();
            full = queue_full();
        }
        if (full) cv_full_.notify_one(); else cv_empty_.notify_one();
        return f;
    }

    Out predict(const In& state)
    {
        float x[AZ_INPUT_FLOATS], pol[AZ_MOVES], val;
        encode_input(&state, x);
        std::lock_guard<std::mutex> g(gpu_lock_);
        check(az_nn_forward(nn_, x, 1, pol, &val, precision_, nullptr), "az_nn_forward");
        return make_out(pol, val);
    }
};

// ---- the reference's class names on top of the service
template <class In, class Out, class TrainData>
class AlphaZeroNNIdT {
    std::shared_ptr<NNService<In, Out>> svc_;
    int gpuIndex_, nnId_;
public:
    AlphaZeroNNIdT(std::shared_ptr<NNService<In, Out>> s, int gpuIndex, int nnId) : svc_(std::move(s)), gpuIndex_(gpuIndex), nnId_(nnId)
    {
        printf("Created NN on gpuIndex %d with id %d\n", gpuIndex, nnId);
    }
    void loadCheckpoint(std::string filePath) { svc_->loadCheckpoint(filePath); }
    void saveCheckpoint(std::string filePath) { svc_->saveCheckpoint(filePath); }
    void train(const std::vector<TrainData>& data, int epochs) { svc_->train(data, epochs); }
    void registerThread() { svc_->registerThread(); }
    void unregisterThread() { svc_->unregisterThread(); }
    std::future<Out> predictFuture(const In& state) { return svc_->predictFuture(state); }
    Out predict(const In& state) { return svc_->predict(state); }
    NNService<In, Out>& service() { return *svc_; }
};

template <class In, class Out, class TrainData>
class AlphaZeroNNGroupT {
    std::string name_;
    std::vector<std::shared_ptr<AlphaZeroNNIdT<In, Out, TrainData>>> ids_;
public:
    explicit AlphaZeroNNGroupT(std::string name) : name_(std::move(name)) { printf("Created NN Group %s\n", name_.c_str()); }
    void add(std::shared_ptr<AlphaZeroNNIdT<In, Out, TrainData>> id) { ids_.push_back(std::move(id)); }
    void loadCheckpoint(std::string filePath) { for (auto& id : ids_) id->loadCheckpoint(filePath); }
    void saveCheckpoint(std::string filePath) { ids_[0]->saveCheckpoint(filePath); }
    // alphazero_gpu_cluster.cpp:221-231: train the first copy, hand the result to the others (the reference goes through a
    // temporary checkpoint file; here az_nn_copy_state moves variables + optimizer state device to device)
    void train(const std::vector<TrainData>& d, int e)
    {
        ids_[0]->train(d, e);
        for (size_t i = 1; i < ids_.size(); ++i) ids_[i]->service().copyStateFrom(ids_[0]->service());
    }
    int size() { return (int)ids_.size(); }
    std::shared_ptr<AlphaZeroNNIdT<In, Out, TrainData>> getNN(int i) { return ids_[i]; }
};

// AlphaZeroCluster::initGpus / initPlayerGroup (alphazero_gpu_cluster.cpp:105-113, 144-164).  `graphFilePath`
// selects the architecture the way the reference's GraphDef file does: "..._<blocks>.pb" -> that many blocks
// (model_bin_V2_5.pb -> 5), anything else -> `defaultBlocks`.
template <class In, class Out, class TrainData>
class AlphaZeroClusterT {
    int gpus_ = 0, precision_, defaultBlocks_;
    std::unordered_map<std::string, std::shared_ptr<AlphaZeroNNGroupT<In, Out, TrainData>>> groups_;
    std::vector<int> nnPerGpu_;
public:
    explicit AlphaZeroClusterT(int precision = AZ_NN_BF16, int defaultBlocks = 5) : precision_(precision), defaultBlocks_(defaultBlocks) { printf("Creating AZ Cluster\n"); }
    void initGpus(int numberOfGpus) { gpus_ = numberOfGpus; nnPerGpu_.assign(numberOfGpus, 0); printf("Initializing gpus: %d\n", numberOfGpus); }
    // the reference's signature (alphazero_gpu_cluster.h:104, called as nnCluster->initGpus(nnCluster, SETTINGS.NUMBER_OF_GPUS),
    // src/alphazero_risk.cpp:7): the self pointer its AlphaZeroGPU objects keep is not needed here
    template <class Self> void initGpus(const std::shared_ptr<Self>&, int numberOfGpus) { initGpus(numberOfGpus); }
    static int blocksFromGraphPath(const std::string& path, int dflt)
    {
        size_t dot = path.rfind(".pb"), us = path.rfind('_');
        if (dot == std::string::npos || us == std::string::npos || us > dot) return dflt;
        int b = atoi(path.substr(us + 1, dot - us - 1).c_str());
        return b >= 1 && b <= 26 ? b : dflt;
    }
    std::shared_ptr<AlphaZeroNNGroupT<In, Out, TrainData>> initPlayerGroup(std::string groupName, std::string graphFilePath)
    {
        if (groups_.count(groupName)) throw std::invalid_argument("Duplicated player group");
        auto group = std::make_shared<AlphaZeroNNGroupT<In, Out, TrainData>>(groupName);
        groups_[groupName] = group;
        for (int g = 0; g < gpus_; ++g) {
            auto svc = std::make_shared<NNService<In, Out>>(blocksFromGraphPath(graphFilePath, defaultBlocks_), g, precision_);
            group->add(std::make_shared<AlphaZeroNNIdT<In, Out, TrainData>>(svc, g, nnPerGpu_[g]++));
        }
        return group;
    }
    std::shared_ptr<AlphaZeroNNGroupT<In, Out, TrainData>> getGroup(std::string name) { return groups_[name]; }
};

// ---- mirror types for standalone builds (layout == the reference's, alphazero_nn_data.h:81-118)
struct LandArmy { uint8_t army : 6; uint8_t playerIndex : 2; };
struct NNInputData {
    LandArmy land[42];
    uint8_t playerIndex = 0;
    uint16_t round = 0;
    float featureReinforcementShare = 0, featureAttackFrequency = 0, featureCanDrawCard = 0;
    float featureIsPhaseSetup = 0, featureIsPhaseSetupNeutral = 0, featureIsPhaseReinforcement = 0, featureIsPhaseAttack = 0,
          featureIsPhaseAttackMobilization = 0, featureIsPhaseFortify = 0;
    float featureArmyShare = 0;
};
struct NNOutputData {
    std::vector<float> policy;
    float value = 0.0f;
    // NNOutputData::normalize, alphazero_nn_data.cpp:3-27
    void normalize(uint64_t validMoves)
    {
        float sum = 0.0f;
        for (size_t i = 0; i < policy.size(); ++i) { if ((validMoves >> i) & 1) sum += policy[i]; else policy[i] = 0.0f; }
        for (size_t i = 0; i < policy.size(); ++i) if (policy[i] > 0.0f) policy[i] /= sum;
    }
};
struct NNTrainData { int8_t playerIndex = 0; NNInputData in; NNOutputData out; };

}  // namespace azb200
