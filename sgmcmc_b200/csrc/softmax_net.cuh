// 784-100-10 softmax-softmax network family (vignettes/nn.Rmd:58-79) -- interface used by api.cu.
#pragma once
#include "common.cuh"

namespace sgm {
struct SoftmaxNetWork;
int softmax_net_create(SoftmaxNetWork** out, int d_in, int d_hidden, int d_out, int n, int chains, int device);
void softmax_net_destroy(SoftmaxNetWork* w);
// internal element -> flat (ABI) index / tensor id tables (A and B are stored row-major internally)
void softmax_net_tables(const SoftmaxNetWork* w, int* flat_of, int* tensor_of, int Pst);
// what sgm_debug_plan reports for this family (computed on the host, no device needed)
struct SoftmaxNetPlan { int threads; size_t smem_bytes; int tmem_cols; int Pst; int offB; int tma_fed; int s1_box_inner; size_t s1_ring_bytes; size_t ring_capacity_bytes; };
int softmax_net_plan(int d_in, int d_hidden, int d_out, SoftmaxNetPlan* out);
// kernel selectors (sgm_set_option keys "nn_update", "nn_stage1"); returns 0 when the key / value is known
int softmax_net_set_option(SoftmaxNetWork* w, const char* key, int value);
int softmax_net_launch(SoftmaxNetWork* w, const ChainArgs& a, cudaStream_t stream, int* launches);
int softmax_net_fullgrad(SoftmaxNetWork* w, const ChainArgs& a, int64_t row0, int64_t row1, int M,
                         const float* theta_hat, double* gsum, cudaStream_t stream);
int softmax_net_fullgrad_finish(SoftmaxNetWork* w, const ChainArgs& a, int M, const double* gsum,
                                const float* theta_hat, float* g_hat, float* theta, cudaStream_t stream);
int softmax_net_logpost(SoftmaxNetWork* w, const ChainArgs& a, uint32_t word2, const float* theta, double* out,
                        cudaStream_t stream);
}  // namespace sgm
