// Shared device-side types for the sampler kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
#include "rng.cuh"

namespace sgm {

constexpr int MAX_TENSORS = 8;   // parameter tensors per model (nn vignette has 8)
constexpr int SMALL_MAXD = 8;    // Gaussian / mixture families: dimension handled in registers
constexpr int SMALL_MAXP = 2 * SMALL_MAXD;

enum Algo : int { ALGO_SGLD = 1, ALGO_SGHMC = 2, ALGO_SGNHT = 3, ALGO_OPT = 4 /* mode finder: plain SGD ascent */,
                  ALGO_GRAD = 5 /* validation: write the gradient estimate into the momentum slot */ };

// Per-tensor constants, rounded to float32 once on the host exactly where the
// reference's graph holds a float32 constant (R/dynamics.R:14-27,42-65,79-116).
struct TensorConsts {
    float half_eps[MAX_TENSORS];     // 0.5 * epsilon                   (sgld drift)
    float eps[MAX_TENSORS];          // epsilon / eta
    float sqrt_2eps[MAX_TENSORS];    // sqrt(2 * epsilon)               (getNoise, dynamics.R:135-142)
    float alpha[MAX_TENSORS];        // sghmc friction
    float sqrt_alpha[MAX_TENSORS];
    float sqrt_a[MAX_TENSORS];       // sgnht
    float sqrt_eps[MAX_TENSORS];     // momentum init sqrt(eta)
    int offset[MAX_TENSORS + 1];     // flat offsets of the tensors
};

// Arguments of one launch of a chain kernel: `iters` sgmcmcSteps for every chain.
struct ChainArgs {
    // dataset (row-major float32, row stride ld floats, 16-byte aligned rows)
    const float* X;
    uint32_t N;
    uint32_t perm_bits;
    int ld;
    int d;
    int n;            // minibatch rows
    float corr;       // float32(N / n), R/setup.R:117
    // chains
    int C;
    int Pst;          // state stride per chain (floats)
    int P;            // scalars per chain
    int T;            // parameter tensors
    float* theta;     // [C][Pst]
    float* mom;       // [C][Pst]
    float* thermo;    // [C][MAX_TENSORS]
    const float* theta_hat;  // [C or 1][Pst] (cv)
    const float* g_hat;      // [C or 1][Pst] (cv)
    int hat_stride;   // Pst, or 0 when the mode is shared by all chains
    int algo;
    int cv;
    int L;            // sub-steps per step (1 unless sghmc)
    int use_prior;
    float prior_inv_var;          // Gaussian families: 1 / prior_sd^2
    float prec[SMALL_MAXD * SMALL_MAXD];  // GAUSS_MEAN precision, row-major
    float logdet_prec;            // log det precision (density constant)
    float opt_lr;     // mode finder step
    TensorConsts tc;
    // rng
    uint32_t k0, k1;
    uint32_t chain0;     // global index of chain 0 of this launch (RNG stream id = chain0 + c)
    uint32_t index_tag;  // TAG_INDEX or TAG_OPT_INDEX
    uint64_t step0;   // step counter of the first iteration of this launch
    int iters;
    // deterministic validation mode (iters == 1)
    int det;
    const int32_t* idx;   // [C][L][n]
    const float* noise;   // [C][noise_len]
    int noise_len;
    // device-resident chain storage: [iter][C][P] float32, flat parameter order
    float* store;
    int64_t store_pos;
    // divergence flag (set to 1 if any state element is non-finite)
    int* diverged;
    // optional phase timing (debug): thread 0 of every CTA adds clock64() deltas, [phase] slots
    unsigned long long* prof;
    // dataset row-sharded over ranks and peer-mapped (loads over NVLink / NVSwitch): global row r lives in shard r / shard_rows.
    // n_shards <= 1: the whole dataset is X (every path except the logistic long-row kernel ignores these fields)
    const float* const* shard_ptr;   // device table [n_shards] of the shards' base pointers (own shard = X)
    uint32_t shard_rows;
    int n_shards;
    int rhat_cached;  // logistic *cv with one shared mode: y - s(x.beta_hat + bias_hat) is stored at column d + 1 of every row
    int gather_tma;   // logistic short-row kernel: 1 = one TMA bulk copy per row, 0 = 16-byte cp.async
};

// base pointer of dataset row `row` (global index when the dataset is sharded over ranks)
__device__ __forceinline__ const float* row_ptr(const ChainArgs& a, uint32_t row) {
    if (a.n_shards > 1) {
        const uint32_t s = row / a.shard_rows;
        return a.shard_ptr[s] + (size_t)(row - s * a.shard_rows) * a.ld;
    }
    return a.X + (size_t)row * a.ld;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t word2_of(uint64_t step, int L, int sub) {
    return (uint32_t)(step * (uint64_t)(L + 1) + (uint64_t)sub);
}

// sigmoid as the reference writes it: 1 / (1 + exp(-z))  (logisticRegression.Rmd:46)
__device__ __forceinline__ float sigmoidf(float z) { return 1.0f / (1.0f + expf(-z)); }

}  // namespace sgm
