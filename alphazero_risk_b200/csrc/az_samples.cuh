// az_samples.cuh — the head of one training-sample record, shared by the self-play recorder (az_mcts.cu) and the recorder of
// scripted / random turns (az_env.cu).
//
// One record = what NNTrainDataStorage::saveTrainingSamples writes per sample (/root/reference/src/risk_game/player/alpha_zero/
// neural_network/alphazero_nn_data.cpp:115-138):
//   int8 playerIndex | NNInputData (88 B, alphazero_nn_data.h:73-101, g++ x86-64 layout) | float value | float policy[43]
// NNInputData: land[42] @0, playerIndex @42, round u16 @44, then 10 floats @48: reinforcementShare, attackFrequency, canDrawCard,
// phase one-hot x 6, armyShare (alphazero_nn_data.cpp:165-196); the three padding bytes (43, 46, 47) are written as zero.
#pragma once

#include "az_game.cuh"

// Reserves `len` consecutive records of an output queue of `cap` records.  count[0] = records committed (never moves past cap, so
// [0, count[0]) is exactly what was written), count[1] = samples dropped.  A game that does not fit is dropped whole and counted;
// it leaves no hole (the reservation is a compare-and-swap that only succeeds when the whole game fits).  One lane calls this.
__device__ __forceinline__ bool az_rec_reserve(unsigned long long* count, unsigned long long cap, uint32_t len, bool fits_staging,
                                               unsigned long long* base)
{
    if (fits_staging) {
        unsigned long long old = *(volatile unsigned long long*)&count[0];
        while (old + len <= cap) {
            const unsigned long long prev = atomicCAS(&count[0], old, old + len);
            if (prev == old) { *base = old; return true; }
            old = prev;
        }
    }
    atomicAdd(&count[1], (unsigned long long)len);
    return false;
}

// Fills s_rec[0..92] (player, NNInputData image, value target for a game that ended with `status`: NNTrainDataStorage::updateValues,
// alphazero_nn_data.cpp:51-65) from the packed primary state st[14].  Called by all 32 lanes of a warp; the caller adds the policy
// (bytes 93..264) and a __syncwarp() before reading s_rec.
__device__ __forceinline__ void az_sample_head(const uint32_t* __restrict__ st, int status, uint8_t* s_rec, int lane)
{
    const unsigned AZ_SAMPLE_FULL = 0xffffffffu;
    uint32_t w = lane < 14 ? st[lane] : 0u;
    // land bytes: lanes 0..10 hold words 0..10
    const uint32_t w10 = __shfl_sync(AZ_SAMPLE_FULL, w, 10), w11 = __shfl_sync(AZ_SAMPLE_FULL, w, 11), w12 = __shfl_sync(AZ_SAMPLE_FULL, w, 12), w13 = __shfl_sync(AZ_SAMPLE_FULL, w, 13);
    AzGame g; g.own0 = g.own1 = g.gt1 = g.full = 0;
    az_unpack_scalars(g, w10, w11, w12, w13);
    int t0 = 0, t1 = 0;
    uint64_t o0 = 0, o1 = 0;
    if (lane < 11) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int l = lane * 4 + b;
            if (l < AZ_LANDS) {
                const uint32_t v = (w >> (8 * b)) & 0xffu;
                if ((v >> 6) == 0) { t0 += (int)(v & 63u); o0 |= 1ull << l; }
                if ((v >> 6) == 1) { t1 += (int)(v & 63u); o1 |= 1ull << l; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t0 += __shfl_xor_sync(AZ_SAMPLE_FULL, t0, o); t1 += __shfl_xor_sync(AZ_SAMPLE_FULL, t1, o);
        o0 |= ((uint64_t)__shfl_xor_sync(AZ_SAMPLE_FULL, (uint32_t)(o0 >> 32), o) << 32) | __shfl_xor_sync(AZ_SAMPLE_FULL, (uint32_t)o0, o);
        o1 |= ((uint64_t)__shfl_xor_sync(AZ_SAMPLE_FULL, (uint32_t)(o1 >> 32), o) << 32) | __shfl_xor_sync(AZ_SAMPLE_FULL, (uint32_t)o1, o);
    }
    const uint32_t cur = g.cur;
    const float ref = (float)az_reinforcement_value(cur ? o1 : o0), eref = (float)az_reinforcement_value(cur ? o0 : o1);
    const float ta = (float)(cur ? t1 : t0), eta = (float)(cur ? t0 : t1);
    float att = __fdiv_rn((float)g.attacks, 8.0f); att = att < 1.0f ? att : 1.0f;
    // assemble the 265 bytes in shared memory (4-byte fields of the record are not 4-byte aligned in the file)
    if (lane < 11) {
#pragma unroll
        for (int b = 0; b < 4; ++b) { const int l = lane * 4 + b; if (l < AZ_LANDS) s_rec[1 + l] = (uint8_t)(w >> (8 * b)); }
    }
    if (lane == 11) {
        s_rec[0] = (uint8_t)cur;
        s_rec[1 + 42] = (uint8_t)cur; s_rec[1 + 43] = 0;
        s_rec[1 + 44] = (uint8_t)(g.round & 0xffu); s_rec[1 + 45] = (uint8_t)(g.round >> 8); s_rec[1 + 46] = 0; s_rec[1 + 47] = 0;
    }
    float f = 0.0f; int foff = -1;
    if (lane == 12) { f = __fdiv_rn(ref, __fadd_rn(ref, eref)); foff = 1 + 48; }
    if (lane == 13) { f = att; foff = 1 + 52; }
    if (lane == 14) { f = g.allow_draw ? 1.0f : 0.0f; foff = 1 + 56; }
    if (lane >= 15 && lane <= 20) { f = g.phase == (uint32_t)(lane - 15) ? 1.0f : 0.0f; foff = 1 + 60 + 4 * (lane - 15); }
    if (lane == 21) { f = __fdiv_rn(ta, __fadd_rn(ta, eta)); foff = 1 + 84; }
    if (lane == 22) { f = status == AZ_STATUS_DRAW ? 0.0f : ((uint32_t)status == cur ? 1.0f : -1.0f); foff = 1 + 88; }
    if (foff >= 0) { const uint32_t u = __float_as_uint(f); for (int b = 0; b < 4; ++b) s_rec[foff + b] = (uint8_t)(u >> (8 * b)); }
}
