#include "softmax_net.cuh"
namespace sgm {
struct SoftmaxNetWork { int dummy; };
int softmax_net_create(SoftmaxNetWork**, int, int, int, int, int, int) { return (int)cudaErrorNotSupported; }
void softmax_net_destroy(SoftmaxNetWork*) {}
int softmax_net_launch(SoftmaxNetWork*, const ChainArgs&, cudaStream_t, int*) { return (int)cudaErrorNotSupported; }
int softmax_net_fullgrad(SoftmaxNetWork*, const ChainArgs&, int64_t, int64_t, int, const float*, double*, cudaStream_t) { return (int)cudaErrorNotSupported; }
int softmax_net_fullgrad_finish(SoftmaxNetWork*, const ChainArgs&, int, const double*, const float*, float*, float*, cudaStream_t) { return (int)cudaErrorNotSupported; }
int softmax_net_logpost(SoftmaxNetWork*, const ChainArgs&, uint32_t, const float*, double*, cudaStream_t) { return (int)cudaErrorNotSupported; }
}
