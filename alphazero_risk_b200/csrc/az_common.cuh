// az_common.cuh — error handling and small host/device helpers shared by the translation
// units of libaz_b200.so.
#pragma once

#include <cstdio>
#include <cstdarg>
#include <cstdint>
#include <cuda_runtime.h>

#include "az_b200.h"

void az_set_error(const char* fmt, ...);

#define AZ_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            az_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));        \
            return AZ_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

#define AZ_REQUIRE(cond, msg)                                                                           \
    do {                                                                                                \
        if (!(cond)) { az_set_error("%s:%d: %s", __FILE__, __LINE__, msg); return AZ_ERR_INVALID_ARG; } \
    } while (0)

// RAII device guard: every entry point runs on the handle's device and restores the caller's
struct AzDeviceGuard {
    int prev = -1; bool ok = true;
    explicit AzDeviceGuard(int dev) { ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess; }
    ~AzDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// uploads the map tables to the current device (idempotent per device); defined in az_env.cu
int az_upload_tables();
// device pointer to the AZ_TABLE_U64 table block on the current device
const uint64_t* az_device_tables();
