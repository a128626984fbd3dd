// Softmax-net family (softmax_net.cu): sizes, per-chain shared-memory carving, tensor maps, small block-level helpers.
#pragma once
#include "softmax_net.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>   // CUtensorMap

namespace sgm {

constexpr int NN_THREADS = 512;
constexpr int NN_WARPS = NN_THREADS / 32;
constexpr int NN_RB = 128;      // rows per block
constexpr int NN_MAXH = 128;    // hidden width limit (lanes x 4)
constexpr int NN_MAXO = 16;     // output width limit

// optional phase timing (debug, ChainArgs::prof): thread 0 of every CTA adds clock64() deltas per stage
__device__ unsigned long long* nn_prof = nullptr;
__device__ int nn_stream_updates = 1;   // 0: register-staged update of every tensor (sgm_set_option "nn_update" = 1)
__device__ int nn_s5_tmem = 0;          // 1: S5 with the X tile on the tensor-memory side (sgm_set_option "nn_stage5" = 1 in builds with -DNN_ENABLE_S5_TMEM; NOT YET RUN ON A GPU)
__device__ long long nn_t_prev[1024];
#define NN_TICK(slot)                                                                                   \
    if (nn_prof && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(nn_prof + (slot), (unsigned long long)(t_ - nn_t_prev[blockIdx.x & 1023])); nn_t_prev[blockIdx.x & 1023] = t_; }

struct NNDims {
    int din, dh, dop;  // dop = output width padded to a multiple of 4
    int dout;
    int offA, offB, offa, offb, offl;  // internal == flat tensor offsets
    int P;
};

struct SoftmaxNetWork {
    NNDims dm;
    int n, chains, device;
    float* G = nullptr;      // [chains][Pst] gradient sums at theta
    float* Ghat = nullptr;   // [chains][Pst] gradient sums at theta_hat (cv)
    float* part = nullptr;   // full-gradient partials
    size_t part_cap = 0;
    size_t smem = 0;
    // TMA descriptors, encoded on first use and cached by (base pointer, extent, stride)
    struct MapEntry { const void* base; long long rows; long long stride; CUtensorMap map; };
    MapEntry maps[8];
    int n_maps = 0;
    // kernel selectors (softmax_net_set_option); 0 = default
    int opt_update = 0;   // 1: register-staged update of every tensor
    int opt_stage1 = 0;   // 1: register-staged forward-GEMM operands (no TMA)
    int opt_stage5 = 0;   // 1: S5 with the X tile on the tensor-memory side (only in builds with -DNN_ENABLE_S5_TMEM)
};

// TMA descriptors of one gradient evaluation: the dataset (rows gathered four at a time) and the parameter array that
// holds this evaluation's B (3-D: input, hidden unit, chain); x == nullptr selects the register-staged path
struct NNMaps {
    const CUtensorMap* x;
    const CUtensorMap* b;
    int bcoord;
};

// internal element order == flat (ABI) order: every tensor in R's column-major element order (A[k + dh*l], B[j + din*n])
__device__ __forceinline__ int nn_flat_of_internal(const NNDims&, int e) { return e; }
__device__ __forceinline__ int nn_tensor_of(const NNDims& dm, int e) {
    if (e < dm.offB) return 0;
    if (e < dm.offa) return 1;
    if (e < dm.offb) return 2;
    if (e < dm.offl) return 3;
    return 4 + (e - dm.offl);
}

// tcgen05 state of one CTA: TMEM base, two stage barriers (operand buffers) and their commit counts
struct TcCtx {
    uint32_t tmem;      // TMEM base address (512 columns allocated for the kernel's lifetime)
    uint32_t bar[2];    // shared-space addresses of the stage mbarriers
    uint32_t commits[2];
};

struct NNSmem {
    float* H;     // [NN_RB][dh]
    float* dZ1;   // [NN_RB][dh]
    float* dZ2;   // [NN_RB][dop]
    float* A;     // [dh][dop] (zero padded)
    float* At;    // [dop][dh] transposed copy (stage 4 reads it with conflict-free float4 loads)
    float* a;     // [dop]
    float* b;     // [dh]
    float* tile;  // max(S1 tiles, S5 tile)
    uint32_t ring;      // shared-space address (1024-byte aligned) of the H | dZ1 | tile region = the S1 operand ring
    uint32_t ring_bytes;
    uint64_t* s1_bars;  // [16] barriers of the S1 pipeline
    float* gb;    // [NN_WARPS][dh] partial column sums
    float* gA;    // [dh*dop] gradient of A (row block accumulation)
    float* ga;    // [dop]
    float* red;   // [64]
    int* rows;    // [NN_RB]
    uint32_t* tc_slot;  // [4]: TMEM base written by tcgen05.alloc; 16-byte aligned
    uint64_t* tc_bars;  // [2]
};

__device__ __forceinline__ float nn_block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = (threadIdx.x < NN_WARPS) ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) { t = warp_sum(t); if (threadIdx.x == 0) red[32] = t; }
    __syncthreads();
    return red[32];
}

// load A (padded), a, b of one parameter vector into shared memory; `which` selects the tensors that changed since the last
// load (bit 0: A, bit 1: a, bit 2: b) -- inside a Gauss-Seidel sub-step only one small tensor moves between evaluations
enum : int { LOAD_A = 1, LOAD_a = 2, LOAD_b = 4, LOAD_ALL = 7 };
__device__ void nn_load_small(const NNDims& dm, const NNSmem& sm, const float* __restrict__ th, int which = LOAD_ALL) {
    if (which & LOAD_A)
        for (int i = threadIdx.x; i < dm.dh * dm.dop; i += NN_THREADS) {
            const int l = i / dm.dh, k = i % dm.dh;   // coalesced over k (A is stored [k + dh * l])
            const float av = (l < dm.dout) ? th[dm.offA + k + dm.dh * l] : 0.f;
            sm.A[k * dm.dop + l] = av;
            sm.At[l * dm.dh + k] = av;
        }
    if (which & LOAD_a)
        for (int i = threadIdx.x; i < dm.dop; i += NN_THREADS) sm.a[i] = (i < dm.dout) ? th[dm.offa + i] : 0.f;
    if (which & LOAD_b)
        for (int i = threadIdx.x; i < dm.dh; i += NN_THREADS) sm.b[i] = th[dm.offb + i];
    __syncthreads();
}

}  // namespace sgm
