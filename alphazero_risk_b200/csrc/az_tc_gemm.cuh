// az_tc_gemm.cuh — the bf16 tcgen05 GEMM of the training step and its operand producers (az_tc_gemm.cu)
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

int az_tg_init();                                             // per device: tables + kernel attributes
int az_tg_splits(int Kp, int want);                           // largest K-split count <= want with no empty split
// C[split][Mp][256] = A[Mp][Kp] * B[256][Kp]^T over the split's K blocks; chunked K-major bf16 operands ([K/8][rows][8])
int az_tg_gemm(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C, int Mp, int Kp, int splits, cudaStream_t s);
int az_tg_reduce(const float* part, int splits, int Mp, int rows_out, float* out, cudaStream_t s);
// scratch16: [cpad/8][rows][8] bf16 (the chunked copy of src the gather reads)
int az_tg_im2col(const float* src, int rows, int cin, int cpad, int Mp, int Kp, __nv_bfloat16* scratch16, __nv_bfloat16* out, cudaStream_t s);
int az_tg_im2col_t(const float* src, int rows, int cin, int Mp, int Kp, __nv_bfloat16* scratch16, __nv_bfloat16* out, cudaStream_t s);
int az_tg_rows_t(const float* dz, int rows, int Kp, __nv_bfloat16* out, cudaStream_t s);
int az_tg_weights(const float* w, int cin, int cpad, int Kp, int flip, __nv_bfloat16* out, cudaStream_t s);
