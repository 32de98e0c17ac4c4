// az_nn_tc.cu — bf16 tcgen05 tensor-core tower (placeholder until the kernel lands in this round).
#include "az_common.cuh"
#include "az_nn.cuh"

int az_nn_tc_prepare(az_nn*) { return AZ_OK; }
void az_nn_tc_release(az_nn*) {}
int az_nn_tc_forward(az_nn*, const float*, const uint32_t*, int, float*, float*, cudaStream_t)
{
    az_set_error("bf16 tensor-core path not built yet");
    return AZ_ERR_NOT_READY;
}
