// az_dist.cu — the two cross-GPU exchanges of the self-play path, over NCCL (NVLink / NVSwitch), OUTSIDE the search loop:
//
//   * az_dist_broadcast_weights — the root's weight blob (every variable of the graph, fp32, device to device) to every other
//     network copy.  Replaces AlphaZeroNNGroup::train's hand-off, where nn[0] saves a temporary checkpoint FILE and the group's
//     copies on the other GPUs reload it (/root/reference/src/risk_game/player/alpha_zero/neural_network/
//     alphazero_gpu_cluster.cpp:221-231).
//   * az_dist_gather_stats (+ the typed az_dist_gather_counters / az_dist_gather_results) — per-GPU tallies summed (and optionally
//     listed per rank).  Replaces GameResults::add over the per-thread results after thread::join
//     (/root/reference/src/risk_game/game/game.cpp:298-309, game/game.h:17-29).
//
// Two process models:
//   az_dist_init       one process drives n GPUs — the reference's model (AlphaZeroCluster::initGpus, alphazero_gpu_cluster.cpp:
//                      147-158, one TF session per GPU inside one process): ncclCommInitAll, rank i = devices[i];
//   az_dist_init_rank  one process per GPU (torchrun-style launch; what bench.py does): ncclCommInitRank with an id created by
//                      az_dist_unique_id on rank 0 and shipped to the others by any side channel.
//
// Games shard by contiguous global id and never migrate, so nothing here is called inside a search; there is no
// compute step that feeds a collective, hence no fused compute + collective kernel.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy already in the process if the host program — e.g. PyTorch — brought
// one, else the system library), so libaz_b200.so itself loads and exports every symbol on a machine without NCCL; az_dist_init*
// then fail with a message.
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <vector>
#include <nccl.h>

#include "az_common.cuh"
#include "az_nn.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

template <class F> bool sym(void* lib, const char* name, F& f) { f = reinterpret_cast<F>(dlsym(lib, name)); return f != nullptr; }

int load_nccl()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return AZ_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { az_set_error("az_dist: cannot load NCCL (libnccl.so.2): %s", dlerror()); return AZ_ERR_NOT_READY; }
    NcclApi a; a.lib = lib;
    bool ok = sym(lib, "ncclGetVersion", a.GetVersion) && sym(lib, "ncclGetUniqueId", a.GetUniqueId) && sym(lib, "ncclCommInitAll", a.CommInitAll) &&
              sym(lib, "ncclCommInitRank", a.CommInitRank) && sym(lib, "ncclCommDestroy", a.CommDestroy) &&
              sym(lib, "ncclGetErrorString", a.GetErrorString) && sym(lib, "ncclGroupStart", a.GroupStart) && sym(lib, "ncclGroupEnd", a.GroupEnd) &&
              sym(lib, "ncclBroadcast", a.Broadcast) && sym(lib, "ncclAllReduce", a.AllReduce) && sym(lib, "ncclAllGather", a.AllGather);
    if (!ok) { az_set_error("az_dist: libnccl.so.2 lacks a required symbol"); dlclose(lib); return AZ_ERR_NOT_READY; }
    g_nccl = a;
    return AZ_OK;
}

#define AZ_NCCL(call)                                                                                   \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            az_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));     \
            return AZ_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

}  // namespace

struct az_dist {
    int world = 0;                        // ranks of the communicator
    int first_rank = 0;                   // rank of local member 0 (single-process model: 0; one process per GPU: this process's rank)
    std::vector<int> devices;             // local members' CUDA devices
    std::vector<ncclComm_t> comms;        // one communicator handle per local member
    std::vector<cudaStream_t> streams;    // one stream per local member: the collectives of a call run concurrently, then all are joined
    std::vector<unsigned long long*> d_buf;   // per local member: staging for the statistics ([n] in, [n] sum, [world][n] gathered)
    size_t buf_words = 0;
};

static int dist_alloc_streams(az_dist* d)
{
    d->streams.assign(d->devices.size(), nullptr);
    d->d_buf.assign(d->devices.size(), nullptr);
    for (size_t i = 0; i < d->devices.size(); ++i) {
        AzDeviceGuard guard(d->devices[i]);
        AZ_CUDA(cudaStreamCreateWithFlags(&d->streams[i], cudaStreamNonBlocking));
    }
    return AZ_OK;
}

extern "C" int az_dist_nccl_version(int* version)
{
    AZ_REQUIRE(version != nullptr, "NULL argument");
    int rc = load_nccl(); if (rc) return rc;
    AZ_NCCL(g_nccl.GetVersion(version));
    return AZ_OK;
}

extern "C" int az_dist_init(int n_devices, const int* devices, az_dist** out)
{
    AZ_REQUIRE(out != nullptr, "NULL argument");
    AZ_REQUIRE(n_devices >= 1, "n_devices must be >= 1");
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { az_set_error("no CUDA device visible"); return AZ_ERR_NO_DEVICE; }
    AZ_REQUIRE(n_devices <= have || devices != nullptr, "more devices requested than are visible");
    int rc = load_nccl(); if (rc) return rc;
    az_dist* d = new (std::nothrow) az_dist();
    AZ_REQUIRE(d != nullptr, "out of host memory");
    d->world = n_devices; d->first_rank = 0;
    for (int i = 0; i < n_devices; ++i) {
        const int dev = devices ? devices[i] : i;
        if (dev < 0 || dev >= have) { delete d; az_set_error("device %d is not visible (%d devices)", dev, have); return AZ_ERR_INVALID_ARG; }
        for (int v : d->devices) if (v == dev) { delete d; az_set_error("device %d listed twice", dev); return AZ_ERR_INVALID_ARG; }
        d->devices.push_back(dev);
    }
    d->comms.assign((size_t)n_devices, nullptr);
    ncclResult_t r = g_nccl.CommInitAll(d->comms.data(), n_devices, d->devices.data());
    if (r != ncclSuccess) { az_set_error("ncclCommInitAll: %s", g_nccl.GetErrorString(r)); delete d; return AZ_ERR_CUDA; }
    rc = dist_alloc_streams(d);
    if (rc) { az_dist_destroy(d); return rc; }
    *out = d;
    return AZ_OK;
}

extern "C" int az_dist_unique_id(uint8_t* id128)
{
    AZ_REQUIRE(id128 != nullptr, "NULL argument");
    int rc = load_nccl(); if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == AZ_DIST_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    AZ_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return AZ_OK;
}

extern "C" int az_dist_init_rank(int world_size, int rank, const uint8_t* id128, int device, az_dist** out)
{
    AZ_REQUIRE(out && id128, "NULL argument");
    AZ_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "rank outside the world");
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { az_set_error("no CUDA device visible"); return AZ_ERR_NO_DEVICE; }
    AZ_REQUIRE(device >= 0 && device < have, "device is not visible");
    int rc = load_nccl(); if (rc) return rc;
    az_dist* d = new (std::nothrow) az_dist();
    AZ_REQUIRE(d != nullptr, "out of host memory");
    d->world = world_size; d->first_rank = rank;
    d->devices.push_back(device);
    d->comms.assign(1, nullptr);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    {
        AzDeviceGuard guard(device);
        ncclResult_t r = g_nccl.CommInitRank(&d->comms[0], world_size, id, rank);
        if (r != ncclSuccess) { az_set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r)); delete d; return AZ_ERR_CUDA; }
    }
    rc = dist_alloc_streams(d);
    if (rc) { az_dist_destroy(d); return rc; }
    *out = d;
    return AZ_OK;
}

extern "C" int az_dist_destroy(az_dist* d)
{
    if (!d) return AZ_OK;
    for (size_t i = 0; i < d->devices.size(); ++i) {
        AzDeviceGuard guard(d->devices[i]);
        if (i < d->streams.size() && d->streams[i]) { cudaStreamSynchronize(d->streams[i]); cudaStreamDestroy(d->streams[i]); }
        if (i < d->d_buf.size()) cudaFree(d->d_buf[i]);
        if (i < d->comms.size() && d->comms[i]) g_nccl.CommDestroy(d->comms[i]);
    }
    delete d;
    return AZ_OK;
}

extern "C" int az_dist_world_size(const az_dist* d) { return d ? d->world : 0; }
extern "C" int az_dist_local_count(const az_dist* d) { return d ? (int)d->devices.size() : 0; }
extern "C" int az_dist_rank(const az_dist* d, int local_index)
{
    return (d && local_index >= 0 && local_index < (int)d->devices.size()) ? d->first_rank + local_index : -1;
}

static int dist_join(az_dist* d)
{
    for (size_t i = 0; i < d->devices.size(); ++i) {
        AzDeviceGuard guard(d->devices[i]);
        AZ_CUDA(cudaStreamSynchronize(d->streams[i]));
    }
    return AZ_OK;
}

// nn[i] = the network copy of local member i (on that member's device).  After the call every copy holds the root's variables
// (weights and BatchNorm moving statistics); the packed bf16 tiles are rebuilt on the next forward / search (az_nn_finalize).
extern "C" int az_dist_broadcast_weights(az_dist* d, az_nn* const* nn, int n_local, int root_rank)
{
    AZ_REQUIRE(d && nn, "NULL argument");
    AZ_REQUIRE(n_local == (int)d->devices.size(), "one network per local member is required");
    AZ_REQUIRE(root_rank >= 0 && root_rank < d->world, "root rank outside the world");
    size_t count = 0;
    for (int i = 0; i < n_local; ++i) {
        AZ_REQUIRE(nn[i] != nullptr, "NULL network");
        AZ_REQUIRE(nn[i]->device == d->devices[(size_t)i], "network i must live on local member i's device");
        if (i == 0) count = nn[i]->blob.size();
        AZ_REQUIRE(nn[i]->blob.size() == count, "networks differ in architecture");
        AzDeviceGuard guard(nn[i]->device);
        const bool is_root = d->first_rank + i == root_rank;
        if (is_root) {
            // the current weights must be on the device: after training they already are (host copy stale), otherwise upload them
            if (!nn[i]->host_stale && !nn[i]->finalized) { int rc = az_nn_finalize(nn[i]); if (rc) return rc; }
        } else if (!nn[i]->d_blob) {
            AZ_CUDA(cudaMalloc(&nn[i]->d_blob, sizeof(float) * count));
        }
    }
    for (int i = 0; i < n_local; ++i) { AzDeviceGuard guard(d->devices[(size_t)i]); AZ_CUDA(cudaDeviceSynchronize()); }   // uploads / earlier work on other streams
    AZ_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < n_local; ++i) {
        AzDeviceGuard guard(d->devices[(size_t)i]);
        ncclResult_t r = g_nccl.Broadcast(nn[i]->d_blob, nn[i]->d_blob, count, ncclFloat, root_rank, d->comms[(size_t)i], d->streams[(size_t)i]);
        if (r != ncclSuccess) { g_nccl.GroupEnd(); az_set_error("ncclBroadcast: %s", g_nccl.GetErrorString(r)); return AZ_ERR_CUDA; }
    }
    AZ_NCCL(g_nccl.GroupEnd());
    int rc = dist_join(d); if (rc) return rc;
    for (int i = 0; i < n_local; ++i) {
        if (d->first_rank + i == root_rank) continue;
        nn[i]->host_stale = true; nn[i]->finalized = false;          // the device copy is now ahead of the host blob (as after training)
    }
    return AZ_OK;
}

// local: [n_local][n] values, member-major.  sum: [n] (may be NULL).  per_rank: [world][n] (may be NULL).
extern "C" int az_dist_gather_stats(az_dist* d, const uint64_t* local, int n, uint64_t* sum, uint64_t* per_rank)
{
    AZ_REQUIRE(d && local, "NULL argument");
    AZ_REQUIRE(n >= 1 && n <= 4096, "n must be in 1..4096");
    const size_t nl = d->devices.size(), words = (size_t)n * (2 + (size_t)d->world);
    if (words > d->buf_words) {
        for (size_t i = 0; i < nl; ++i) {
            AzDeviceGuard guard(d->devices[i]);
            cudaFree(d->d_buf[i]); d->d_buf[i] = nullptr;
            AZ_CUDA(cudaMalloc(&d->d_buf[i], sizeof(unsigned long long) * words));
        }
        d->buf_words = words;
    }
    for (size_t i = 0; i < nl; ++i) {
        AzDeviceGuard guard(d->devices[i]);
        AZ_CUDA(cudaMemcpyAsync(d->d_buf[i], local + i * (size_t)n, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, d->streams[i]));
    }
    AZ_NCCL(g_nccl.GroupStart());
    for (size_t i = 0; i < nl; ++i) {
        AzDeviceGuard guard(d->devices[i]);
        unsigned long long* in = d->d_buf[i];
        ncclResult_t r = g_nccl.AllReduce(in, in + n, (size_t)n, ncclUint64, ncclSum, d->comms[i], d->streams[i]);
        if (r == ncclSuccess && per_rank) r = g_nccl.AllGather(in, in + 2 * (size_t)n, (size_t)n, ncclUint64, d->comms[i], d->streams[i]);
        if (r != ncclSuccess) { g_nccl.GroupEnd(); az_set_error("nccl collective: %s", g_nccl.GetErrorString(r)); return AZ_ERR_CUDA; }
    }
    AZ_NCCL(g_nccl.GroupEnd());
    {   // every member holds the same result: read member 0's
        AzDeviceGuard guard(d->devices[0]);
        if (sum) AZ_CUDA(cudaMemcpyAsync(sum, d->d_buf[0] + n, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, d->streams[0]));
        if (per_rank) AZ_CUDA(cudaMemcpyAsync(per_rank, d->d_buf[0] + 2 * (size_t)n, sizeof(uint64_t) * (size_t)n * (size_t)d->world, cudaMemcpyDeviceToHost, d->streams[0]));
    }
    return dist_join(d);
}

// GameResults::add for the rollout / self-play counters: local[n_local] -> total (sum over every rank)
extern "C" int az_dist_gather_counters(az_dist* d, const az_counters* local, az_counters* total)
{
    AZ_REQUIRE(d && local && total, "NULL argument");
    static_assert(sizeof(az_counters) == 9 * sizeof(uint64_t), "az_counters is nine 64-bit counters");
    return az_dist_gather_stats(d, reinterpret_cast<const uint64_t*>(local), 9, reinterpret_cast<uint64_t*>(total), nullptr);
}

// GameResults::add for arena matches (`ticks` becomes the sum of the members' ticks)
extern "C" int az_dist_gather_results(az_dist* d, const az_arena_results* local, az_arena_results* total)
{
    AZ_REQUIRE(d && local && total, "NULL argument");
    static_assert(sizeof(az_arena_results) == 12 * sizeof(uint64_t), "az_arena_results is twelve 64-bit counters");
    return az_dist_gather_stats(d, reinterpret_cast<const uint64_t*>(local), 12, reinterpret_cast<uint64_t*>(total), nullptr);
}

// all ranks have reached this point and their earlier device work is complete
extern "C" int az_dist_barrier(az_dist* d)
{
    AZ_REQUIRE(d != nullptr, "NULL argument");
    for (size_t i = 0; i < d->devices.size(); ++i) { AzDeviceGuard guard(d->devices[i]); AZ_CUDA(cudaDeviceSynchronize()); }
    std::vector<uint64_t> one(d->devices.size(), 1), s(1);
    int rc = az_dist_gather_stats(d, one.data(), 1, s.data(), nullptr); if (rc) return rc;
    if (s[0] != (uint64_t)d->world) { az_set_error("az_dist_barrier: %llu of %d ranks arrived", (unsigned long long)s[0], d->world); return AZ_ERR_BAD_STATE; }
    return AZ_OK;
}
