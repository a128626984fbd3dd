// TEST INFRASTRUCTURE ONLY: stands in for csrc/softmax_net.cu while that file is not part of the emulated build.
#define SGM_EMU 1
#include "../../sgmcmc_b200/csrc/softmax_net.cuh"
namespace sgm {
int softmax_net_create(SoftmaxNetWork**, int, int, int, int, int, int) { return cudaErrorInvalidValue; }
void softmax_net_destroy(SoftmaxNetWork*) {}
void softmax_net_tables(const SoftmaxNetWork*, int*, int*, int) {}
int softmax_net_set_option(SoftmaxNetWork*, const char*, int) { return -1; }
int softmax_net_launch(SoftmaxNetWork*, const ChainArgs&, cudaStream_t, int*) { return cudaErrorInvalidValue; }
int softmax_net_fullgrad(SoftmaxNetWork*, const ChainArgs&, int64_t, int64_t, int, const float*, double*, cudaStream_t) { return cudaErrorInvalidValue; }
int softmax_net_fullgrad_finish(SoftmaxNetWork*, const ChainArgs&, int, const double*, const float*, float*, float*, cudaStream_t) { return cudaErrorInvalidValue; }
int softmax_net_logpost(SoftmaxNetWork*, const ChainArgs&, uint32_t, const float*, double*, cudaStream_t) { return cudaErrorInvalidValue; }
}
