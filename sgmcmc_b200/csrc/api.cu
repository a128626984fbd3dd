// C-ABI implementation (include/sgmcmc_b200.h): handle lifecycle, dataset upload,
// kernel dispatch, chain storage, NCCL row-sharded control-variate pass.
// No CPU fallback: every entry point needs a CUDA device and fails with SGM_ECUDA
// otherwise.
#include <climits>
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <algorithm>
#include <vector>
#include <dlfcn.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/sgmcmc_b200.h"
#include "common.cuh"
#include "small_models.cuh"
#include "logistic.cuh"
#include "logistic_staged.cuh"
#include "logistic_split.cuh"
#include "logistic_fullgrad_stream.cuh"
#include "logistic_cv_gemm.cuh"
#include "softmax_net.cuh"

using namespace sgm;

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(SGM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* sgm_last_error(void) { return g_err.c_str(); }
extern "C" const char* sgm_version(void) { return "sgmcmc_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------
// NCCL through dlopen (the library must load on hosts without NCCL)
// ------------------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
struct NcclId { char internal[SGM_NCCL_ID_BYTES]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, /*ncclUniqueId by value*/ NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
// communicators are process-level resources (creating one costs hundreds of milliseconds): handles that are given the same unique
// id share one ncclComm, which then lives until the process exits
struct SharedComm { ncclComm_t comm; int refs; int rank, world; };
static std::map<std::string, SharedComm> g_comms;

static int load_nccl() {
    if (g_nccl.lib) return SGM_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
#ifdef SGM_EMU   // the CPU-emulated test build talks to tests/emu/fake_nccl.c (multi-process collectives over shared memory)
    if (const char* fake = getenv("SGM_EMU_NCCL")) names[0] = fake;
#endif
    void* lib = nullptr;
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return fail(SGM_ECOMM, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
    g_nccl.ReduceScatter = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclReduceScatter");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.ReduceScatter || !g_nccl.CommDestroy)
        return fail(SGM_ECOMM, "libnccl lacks required symbols");
    g_nccl.lib = lib;
    return SGM_OK;
}

// ------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------
struct TensorInfo {
    std::string name;
    int ndim;
    int64_t dims[4];
    int64_t size;
    int64_t offset;  // flat
};

struct sgm_handle {
    int model = 0, algo = 0, cv = 0;
    int64_t N = 0;
    int n = 0, d = 0, ld = 0, P = 0, Pst = 0, T = 0, C = 1, L = 1;
    int device = 0;
    uint32_t flags = 0;
    int use_prior = 0;
    double prior_sd = 0;
    double opt_stepsize = 0;
    int64_t n_iters_opt = 0;
    uint64_t seed = 0;
    int64_t chain_offset = 0;
    std::vector<TensorInfo> tensors;
    std::vector<double> eps, alpha, a;
    std::vector<float> init_state;   // [Pst] start values in the internal layout (restored by every sgm_init)
    // device
    float* X = nullptr;
    float *theta = nullptr, *mom = nullptr, *thermo = nullptr, *theta_hat = nullptr, *g_hat = nullptr, *scratch = nullptr;
    double* gsum = nullptr;
    double* dout = nullptr;   // [C] logpost values
    int* d_flag = nullptr;
    int32_t* d_idx = nullptr; size_t d_idx_cap = 0;
    float* d_noise = nullptr; size_t d_noise_cap = 0;
    float* store = nullptr; int64_t store_cap = 0, store_len = 0;
    cudaStream_t stream = nullptr;
    // sgm_sample: chain streamed to the host while later iterations run (double-buffered blocks)
    struct Pipe {
        cudaStream_t copy = nullptr;
        float* d_store[2] = {nullptr, nullptr};   // [blk][C][P]   written by the chain kernel
        float* d_T[2] = {nullptr, nullptr};       // [C][P][blk]   transposed
        float* h_T[2] = {nullptr, nullptr};       // pinned host copies of d_T
        double* d_lp[2] = {nullptr, nullptr};     // [C] divergence-check log posteriors
        double* h_lp[2] = {nullptr, nullptr};
        int* h_flag[2] = {nullptr, nullptr};
        cudaEvent_t ran[2] = {nullptr, nullptr}, transposed[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
        int64_t blk = 0;
    } pipe;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0; int last_launches = 0;
    float fullgrad_ms = 0; int64_t fullgrad_bytes = 0;
    float collective_ms = 0, gather_ms = 0;   // device time of the last full-data pass's collectives (reduction; all-gather of the modes)
    uint64_t step = 0, check_ctr = 0;
    bool inited = false;
    CUtensorMap tmap;           // 2-D map of the dataset (box = one row) for TMA tile::gather4
    bool tmap_ok = false;
    float* Rt = nullptr; size_t rt_cap = 0;         // residuals of one row block, transposed [M][rt_pitch]
    int* d_flat_of = nullptr; double* d_state_out = nullptr; size_t d_state_out_cap = 0;   // getParams / getState: index table, conversion buffer
    int lsp_asked = 0, lsp_n = 0, lsp_granted = 0;  // cluster kernel: the cluster size asked for last, at which minibatch size, and what the device granted
    // dataset row-sharded over the ranks, peers mapped through CUDA IPC (sgm_attach_peer_shards): every rank holds N rows, the
    // chains sample from world x N rows and gather remote rows over NVLink; the control-variate pass sums the local shard
    int n_shards = 1;
    void* peer_base[16] = {nullptr};
    const float** d_shard_ptr = nullptr;
    float* hat_all = nullptr; double* gsum_all = nullptr;   // row-sharded pass with one mode per chain: the modes of ALL ranks / their sums
    double fullgrad_flops = 0;                      // 4 N d M of the last tensor-core full gradient (0: not used)
    bool rhat_cached = false;   // logistic *cv, shared mode: y - s(x.beta_hat) stored in the rows' spare slot
    sgm_progress_fn progress = nullptr; void* progress_user = nullptr;   // every 100 iterations of the mode finder (checkOptDivergence)
    ChainArgs base;  // template of kernel arguments
    // kernel selectors (sgm_set_option; 0 = the default dispatch).  They exist so that tests can pin a path through the ABI;
    // nothing on the dispatch path reads the environment.
    int opt_logistic_kernel = 0;   // 1: generic warp-group kernel even for short rows
    int opt_gather = 0;            // short-row kernel: 1 cp.async, 2 TMA bulk copy per row, 3 mixed, 4 TMA tile::gather4
    int opt_fullgrad = 0;          // logistic full-data gradient: 1 tensor-core contractions, 2 SIMT row-chunk x mode grid
    int opt_small_kernel = 0;      // Gaussian families: 1 = rows from global memory even when the dataset fits in shared memory
    int opt_sample_out = 0;        // sgm_sample: 1 = staged float32 copy + host widening even into a page-locked caller array
    int opt_mode_finder_split = 0; // logistic mode finder: 1 = always one CTA per mode; 2, 4, 8 = a cluster of that many CTAs per mode (default: by occupancy)
    int opt_chain_split = 0;       // logistic sampling kernels: 1 = always one CTA per chain; 2, 4, 8 = a cluster of that many CTAs per chain (default: by occupancy)
    int opt_sharded_finder = 0;    // 1: mode finder sharded by data rows (each rank draws its part of every minibatch from its own rows, the
                                   // P-vector is all-reduced every SGD step); needs sgm_comm_init + SGM_FLAG_SHARED_MODE
    int opt_row0 = 0, opt_row1 = 0;   // row1 > 0: the full-data pass sums rows [row0, row1) only -- what rank r of a row-sharded job does,
                                      // without a communicator (single-GPU pre-flight of shard starts; the sums are then partial sums)
    SoftmaxNetWork* nn = nullptr;
    // comm
    ncclComm_t comm = nullptr; int rank = 0, world = 1;
    std::string comm_key;
};

static int hat_count(const sgm_handle* h) { return (h->flags & SGM_FLAG_SHARED_MODE) ? 1 : h->C; }

// ------------------------------------------------------------------------------------
// small kernels used by the host layer
// ------------------------------------------------------------------------------------
// double [N, cols] (column-major, or row-major when row_major != 0) -> row-major float rows with stride ld at column offset col0;
// reads are coalesced in either layout
__global__ void pack_rows_kernel(const double* __restrict__ src, int64_t N, int cols, float* __restrict__ dst, int ld, int col0, int row_major) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * cols) return;
    const int64_t r = row_major ? i / cols : i % N; const int c = (int)(row_major ? i % cols : i / N);
    dst[r * ld + col0 + c] = (float)src[i];
}

// one slab of a dataset entry as float32, in the caller's layout ([rows][cols] row-major, or [cols][rows] column-major: the host only
// narrows, the transposition of R's layout happens here, reads coalesced either way) -> rows [row0, row0 + rows) of the dataset at
// column offset col0
__global__ void pack_slab_kernel(const float* __restrict__ src, int64_t rows, int cols, int row_major, float* __restrict__ dst, int64_t row0, int ld, int col0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = row_major ? i / cols : i % rows; const int c = (int)(row_major ? i - r * cols : i / rows);
    dst[(row0 + r) * ld + col0 + c] = src[i];
}

__global__ void unpack_rows_kernel(const float* __restrict__ src, int64_t row0, int64_t rows, int cols, int ld, int col0, double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i % rows; const int c = (int)(i / rows);
    dst[i] = (double)src[(row0 + r) * ld + col0 + c];
}

// synthetic logistic dataset: x ~ N(0,1) from Philox(TAG_DATA), planted beta* ~ N(0, 1/d),
// y ~ Bernoulli(sigmoid(x . beta*)).  One warp per row.
__global__ void synth_logistic_kernel(float* __restrict__ X, int64_t N, int d, int ld, uint32_t k0, uint32_t k1) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    const float inv_sqrt_d = rsqrtf((float)d);
    float z = 0.f;
    float* xr = X + row * ld;
    for (int b = lane; 4 * b < ld; b += 32) {
        float v[4], bs[4];
        normal4_at((uint32_t)b, (uint32_t)(row & 0xffffffffu), (uint32_t)(row >> 32), TAG_DATA, k0, k1, v);
        normal4_at((uint32_t)b, 0xffffffffu, 0xffffffffu, TAG_DATA, k0 ^ 0x5bd1e995u, k1, bs);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int col = 4 * b + k;
            if (col < d) { z += v[k] * bs[k] * inv_sqrt_d; } else v[k] = 0.f;
        }
        *reinterpret_cast<float4*>(xr + 4 * b) = make_float4(v[0], v[1], v[2], v[3]);
    }
    z = warp_sum(z);
    if (lane == 0) {
        const U4 r = philox4x32(0xfffffffeu, (uint32_t)(row & 0xffffffffu), (uint32_t)(row >> 32), TAG_DATA, k0, k1);
        const float u = (float)(r.x >> 8) * 5.9604644775390625e-8f;
        xr[d] = (u < 1.0f / (1.0f + expf(-z))) ? 1.0f : 0.0f;
    }
}

__global__ void fill_kernel(float* p, float v, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

// momentum init: sqrt(eta_t) * N(0,1) (R/dynamics.R:47,84); internal index e -> (flat f, tensor t) via tables
__global__ void momentum_init_kernel(float* __restrict__ mom, int C, int Pst, int P, const int* __restrict__ flat_of,
                                     const int* __restrict__ tensor_of, TensorConsts tc, const float* __restrict__ noise,
                                     uint32_t k0, uint32_t k1, uint32_t chain0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)C * Pst) return;
    const int c = (int)(i / Pst), e = (int)(i % Pst);
    const int f = flat_of[e];
    if (f < 0) { mom[i] = 0.f; return; }
    const float z = noise ? noise[(size_t)c * P + f] : normal_at((uint32_t)f, chain0 + (uint32_t)c, 0u, TAG_MOMENTUM, k0, k1);
    mom[i] = tc.sqrt_eps[tensor_of[e]] * z;
}

// [iter][C][P] float -> [C][P][iter] (R layout: every tensor a column-major (iters, dims...) array); OutT = double for the one-shot
// copy-out, float for the streamed one (widened on the host: half the PCIe bytes)
template <class OutT>
__global__ void chain_transpose_kernel(const float* __restrict__ store, int64_t iters, int C, int P, int c0, int cn,
                                       OutT* __restrict__ out) {
    __shared__ float tile[32][33];
    const int c = c0 + blockIdx.z;
    const int64_t it0 = (int64_t)blockIdx.x * 32; const int f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t it = it0 + r; const int f = f0 + threadIdx.x;
        tile[r][threadIdx.x] = (it < iters && f < P) ? store[((size_t)it * C + c) * P + f] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int f = f0 + r; const int64_t it = it0 + threadIdx.x;
        if (f < P && it < iters) out[((size_t)blockIdx.z * P + f) * iters + it] = (OutT)tile[threadIdx.x][r];
    }
    (void)cn;
}

// sgm_sample into a PAGE-LOCKED caller array: block [cnt][C][P] float -> out[(c * P + f) * n_iters + it0 + i] as doubles, written by the
// GPU straight into host memory (coalesced 256-byte runs along the iterations): no staging copy, no host thread touches the chain
__global__ void chain_block_to_host_kernel(const float* __restrict__ store, int64_t cnt, int C, int P, int c0, int64_t n_iters, int64_t it0,
                                           double* __restrict__ out) {
    __shared__ float tile[32][33];
    const int c = c0 + blockIdx.z;
    if (threadIdx.y == 0) SGM_EMU_NOTE("sample:direct_to_host");
    const int64_t i0 = (int64_t)blockIdx.x * 32; const int f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t i = i0 + r; const int f = f0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < cnt && f < P) ? store[((size_t)i * C + c) * P + f] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int f = f0 + r; const int64_t i = i0 + threadIdx.x;
        if (f < P && i < cnt) out[((size_t)c * P + f) * n_iters + it0 + i] = (double)tile[threadIdx.x][r];
    }
}

__global__ void state_to_double_kernel(const float* __restrict__ src, int C, int Pst, int P, const int* __restrict__ flat_of, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)C * Pst) return;
    const int c = (int)(i / Pst), e = (int)(i % Pst);
    const int f = flat_of[e];
    if (f >= 0) out[(size_t)c * P + f] = (double)src[i];
}

// row-sharded mode finder: this rank's part of one minibatch = n_r distinct rows of ITS row range [row0, row0 + rows), as absolute row numbers
__global__ void opt_indices_kernel(int32_t* __restrict__ idx, int n_r, uint32_t row0, uint32_t rows, uint32_t bits, uint32_t stream_id, uint32_t word2,
                                   uint32_t k0, uint32_t k1) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_r) return;
    const Perm p = make_perm(rows, bits, stream_id, word2, TAG_OPT_INDEX, k0, k1);
    idx[j] = (int32_t)(row0 + perm_at(p, (uint32_t)j));
}
// theta_hat += lr * g   (plain SGD ascent, R/controlVariates.R:87-88), g = all-reduced estimate (+ the prior gradient computed apart)
__global__ void opt_apply_kernel(float* __restrict__ theta_hat, const float* __restrict__ g, const float* __restrict__ g_prior, float lr, int count, int* __restrict__ diverged) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const float v = theta_hat[e] + lr * (g[e] + (g_prior ? g_prior[e] : 0.f));
    theta_hat[e] = v;
    if (!isfinite(v)) *diverged = 1;
}

__global__ void broadcast_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int C, int Pst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)C * Pst) dst[i] = src[i % Pst];
}

// Gaussian families: full-data sums at the mode(s), double atomics into gsum[m][Pst]
template <int MODEL>
__global__ void small_fullsum_kernel(ChainArgs a, int64_t row0, int64_t row1, double* __restrict__ gsum) {
    const int m = blockIdx.y;
    const int lane = threadIdx.x & 31;
    float hat[SMALL_MAXP], s[SMALL_MAXP];
#pragma unroll
    for (int k = 0; k < SMALL_MAXP; ++k) {
        const int f = small_flat_of_slot<MODEL>(k, a.d);
        hat[k] = f >= 0 ? a.theta_hat[(size_t)m * a.Pst + f] : 0.f;
        s[k] = 0.f;
    }
    const int d = a.d;
    for (int64_t j = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < row1; j += (int64_t)gridDim.x * blockDim.x) {
        const float* xr = a.X + (size_t)j * a.ld;
        float x[SMALL_MAXD];
#pragma unroll
        for (int k = 0; k < SMALL_MAXD; ++k) x[k] = (k < d) ? __ldg(xr + k) : 0.f;
        if (MODEL == SMALL_MODEL_GAUSS) {
#pragma unroll
            for (int k = 0; k < SMALL_MAXD; ++k) if (k < d) s[k] += x[k] - hat[k];
        } else {
            float q1 = 0.f, q2 = 0.f;
#pragma unroll
            for (int k = 0; k < SMALL_MAXD; ++k)
                if (k < d) { const float e1 = x[k] - hat[k], e2 = x[k] - hat[SMALL_MAXD + k]; q1 += e1 * e1; q2 += e2 * e2; }
            const float p1 = 1.0f / (1.0f + expf(0.5f * (q1 - q2))), p2 = 1.0f - p1;
#pragma unroll
            for (int k = 0; k < SMALL_MAXD; ++k)
                if (k < d) { s[k] += p1 * (x[k] - hat[k]); s[SMALL_MAXD + k] += p2 * (x[k] - hat[SMALL_MAXD + k]); }
        }
    }
#pragma unroll
    for (int k = 0; k < SMALL_MAXP; ++k) {
        const int f = small_flat_of_slot<MODEL>(k, d);
        if (f < 0) continue;
        const float t = warp_sum(s[k]);
        if (lane == 0) atomicAdd(gsum + (size_t)m * a.Pst + f, (double)t);
    }
}

template <int MODEL>
__global__ void small_fullgrad_finish(ChainArgs a, int M, const double* __restrict__ gsum, float* __restrict__ g_hat,
                                      float* __restrict__ theta) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int d = a.d;
    for (int f = 0; f < a.P; ++f) {
        const int kk = f % d;
        float ll;
        if (MODEL == SMALL_MODEL_GAUSS) {
            ll = 0.f;
            for (int q = 0; q < d; ++q) ll += a.prec[kk * SMALL_MAXD + q] * (float)gsum[(size_t)m * a.Pst + q];
        } else {
            ll = (float)gsum[(size_t)m * a.Pst + f];
        }
        const float th = a.theta_hat[(size_t)m * a.Pst + f];
        g_hat[(size_t)m * a.Pst + f] = (a.use_prior ? -th * a.prior_inv_var : 0.f) + ll;
        theta[(size_t)m * a.Pst + f] = th;
    }
}

// ------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// scratch device buffers / events of one ABI call: released on every return path
template <class T>
struct DevBuf {
    T* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;              // (a launch site must name the raw pointer, never the owner)
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { return cudaMalloc(&p, count * sizeof(T)); }
};
struct EventPair {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~EventPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

static int set_device(const sgm_handle* h) {
    CU(cudaSetDevice(h->device));
    return SGM_OK;
}

static void host_tables(const sgm_handle* h, std::vector<int>& flat_of, std::vector<int>& tensor_of) {
    flat_of.assign(h->Pst, -1);
    tensor_of.assign(h->Pst, 0);
    if (h->model == SGM_MODEL_LOGISTIC) {
        for (int e = 0; e < h->d; ++e) { flat_of[e] = e + 1; tensor_of[e] = 1; }
        flat_of[h->d] = 0; tensor_of[h->d] = 0;
    } else if (h->model == SGM_MODEL_SOFTMAX_NET && h->nn) {
        softmax_net_tables(h->nn, flat_of.data(), tensor_of.data(), h->Pst);
    } else {
        for (int t = 0; t < h->T; ++t)
            for (int64_t k = 0; k < h->tensors[t].size; ++k) {
                flat_of[h->tensors[t].offset + k] = (int)(h->tensors[t].offset + k);
                tensor_of[h->tensors[t].offset + k] = t;
            }
    }
}

struct LogiShape { int G, V; };
static bool logi_shape(int ld, LogiShape* s) {
    const int q = ld / 4;
    const int Gs[] = {1, 2, 4, 8, 16, 32};
    for (int G : Gs) if (G >= q) { s->G = G; s->V = 1; return true; }
    const int Vs[] = {2, 3, 4, 5, 6, 8, 9, 12, 16};
    for (int V : Vs) if (32 * V >= q) { s->G = 32; s->V = V; return true; }
    return false;
}

static int logi_threads(const sgm_handle* h, const LogiShape& s) {
    // rows in flight per CTA = threads / G; keep it <= minibatch, and keep register use in mind
    int thr = (s.V >= 5) ? 512 : 256;
    while (thr > 32 && thr / s.G > h->n) thr >>= 1;
    return thr;
}

static size_t logi_smem_bytes(int threads, const LogiShape& s, int n) {
    const int ldp = s.V * s.G * 4;
    return sizeof(float) * ((size_t)(5 + threads / 32) * ldp + 32 + 3 * (size_t)n);
}

template <int G, int V>
static int launch_logistic_gv(sgm_handle* h, const ChainArgs& a, int threads, size_t smem) {
    auto k = logistic_chain_kernel<G, V>;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    int grid = a.C < sms * (per_sm > 0 ? per_sm : 1) ? a.C : sms * (per_sm > 0 ? per_sm : 1);
    SGM_LAUNCH(k, grid, threads, smem, h->stream, a);
    CU(cudaGetLastError());
    return SGM_OK;
}

template <int G, int V>
static int launch_logistic_fullgrad_gv(sgm_handle* h, int64_t row0, int64_t row1, int M, const float* theta_hat) {
    auto k = logistic_fullgrad_kernel<G, V>;
    const int threads = (V >= 5) ? 512 : 256;
    const int ldp = V * G * 4;
    const size_t smem = sizeof(float) * (size_t)(1 + threads / 32) * ldp;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const int64_t rows = row1 - row0;
    // >= 4 CTAs per SM worth of chunks, each at least one sweep of the CTA's row groups
    int64_t chunks = (int64_t)sms * 8;
    int64_t rows_per_cta = cdiv(rows, chunks);
    const int64_t min_rows = 4 * (threads / G);
    if (rows_per_cta < min_rows) rows_per_cta = min_rows;
    chunks = cdiv(rows, rows_per_cta);
    if (chunks < 1) chunks = 1;
    dim3 grid((unsigned)chunks, (unsigned)M);
    SGM_LAUNCH(k, grid, threads, smem, h->stream, h->X, h->ld, h->d, row0, row1, (int)rows_per_cta, theta_hat, h->Pst, h->gsum);
    CU(cudaGetLastError());
    return SGM_OK;
}

#define LOGI_DISPATCH(FN, ...)                                                                 \
    do {                                                                                       \
        LogiShape s_;                                                                          \
        if (!logi_shape(h->ld, &s_)) return fail(SGM_EINVAL, "logistic: d = %d too large", h->d); \
        switch (s_.G * 100 + s_.V) {                                                           \
            case 101: return FN<1, 1>(__VA_ARGS__);                                            \
            case 201: return FN<2, 1>(__VA_ARGS__);                                            \
            case 401: return FN<4, 1>(__VA_ARGS__);                                            \
            case 801: return FN<8, 1>(__VA_ARGS__);                                            \
            case 1601: return FN<16, 1>(__VA_ARGS__);                                          \
            case 3201: return FN<32, 1>(__VA_ARGS__);                                          \
            case 3202: return FN<32, 2>(__VA_ARGS__);                                          \
            case 3203: return FN<32, 3>(__VA_ARGS__);                                          \
            case 3204: return FN<32, 4>(__VA_ARGS__);                                          \
            case 3205: return FN<32, 5>(__VA_ARGS__);                                          \
            case 3206: return FN<32, 6>(__VA_ARGS__);                                          \
            case 3208: return FN<32, 8>(__VA_ARGS__);                                          \
            case 3209: return FN<32, 9>(__VA_ARGS__);                                          \
            case 3212: return FN<32, 12>(__VA_ARGS__);                                         \
            case 3216: return FN<32, 16>(__VA_ARGS__);                                         \
        }                                                                                      \
        return fail(SGM_EINVAL, "logistic: no kernel for ld = %d", h->ld);                     \
    } while (0)

// short rows: whole minibatch staged in shared memory by TMA bulk copies (logistic_staged.cuh)
static bool logistic_use_staged(const sgm_handle* h, const ChainArgs& a) {
    if (h->opt_logistic_kernel == 1) return false;
    if (h->n_shards > 1) return false;   // the TMA row map covers the local shard only
    return h->ld <= 128 && a.n <= LS_MAX_R * LS_THREADS && ls_smem_bytes(h->ld, a.n) <= 227 * 1024;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
static int make_row_tensor_map(sgm_handle* h) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { cudaGetLastError(); return -1; }
    const cuuint64_t gdim[2] = {(cuuint64_t)h->ld, (cuuint64_t)h->N};
    const cuuint64_t gstride[1] = {(cuuint64_t)h->ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)h->ld, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = ((EncodeFn)fn)(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->X, gdim, gstride, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? 0 : -1;
}

template <bool HAT_DOT, int R>
static int launch_logistic_staged_t(sgm_handle* h, const ChainArgs& a) {
    auto k = logistic_staged_kernel<HAT_DOT, R>;
    const size_t smem = ls_smem_bytes(h->ld, a.n);
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0, per_sm = 1;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, LS_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = a.C < sms * per_sm ? a.C : sms * per_sm;
    SGM_LAUNCH(k, grid, LS_THREADS, smem, h->stream, a, h->tmap);
    CU(cudaGetLastError());
    return SGM_OK;
}

// tile::gather4 drops four rows contiguously at a 128-byte aligned shared-memory address: the rows must be unpadded there (stride == ld)
// and four of them a whole number of 128-byte lines (ld % 8 == 0), or every second group of four would start 64 bytes off -- a
// misaligned TMA destination (d = 72: ld = 76, found by the shape fuzz on the emulated build; every tested shape had ld % 8 == 0)
static bool logistic_can_gather4(const sgm_handle* h) {
    return h->tmap_ok && (ls_row_stride4(h->ld) == h->ld / 4) && h->ld % 8 == 0 && h->ld <= 256;
}

static int launch_logistic_staged(sgm_handle* h, const ChainArgs& a) {
    // control variates need x.beta_hat per row unless the residual at the (shared) mode is cached in the rows
    const bool hat_dot = a.cv && !h->rhat_cached;
    {   // gather engine: TMA tile::gather4 when the shared-memory rows are unpadded, else one TMA bulk copy per row
        const bool can_gather4 = logistic_can_gather4(h);
        int mode = can_gather4 ? 3 : 1;
        switch (h->opt_gather) {
            case 1: mode = 0; break;
            case 2: mode = 1; break;
            case 3: mode = 2; break;
            case 4: if (!can_gather4) return fail(SGM_EINVAL, "gather = tile::gather4 needs unpadded shared-memory rows of ld %% 8 == 0 floats and a tensor map (ld = %d)", h->ld); mode = 3; break;
        }
        const_cast<ChainArgs&>(a).gather_tma = mode;
    }
    const int R = (a.n + LS_THREADS - 1) / LS_THREADS;
    switch ((hat_dot ? 10 : 0) + R) {
        case 1: return launch_logistic_staged_t<false, 1>(h, a);
        case 2: return launch_logistic_staged_t<false, 2>(h, a);
        case 3: return launch_logistic_staged_t<false, 3>(h, a);
        case 4: return launch_logistic_staged_t<false, 4>(h, a);
        case 11: return launch_logistic_staged_t<true, 1>(h, a);
        case 12: return launch_logistic_staged_t<true, 2>(h, a);
        case 13: return launch_logistic_staged_t<true, 3>(h, a);
        case 14: return launch_logistic_staged_t<true, 4>(h, a);
    }
    return fail(SGM_EINVAL, "logistic staged kernel: minibatch %d too large", a.n);
}

// ---- fewer chains (or modes) than SMs: a thread-block cluster per chain (logistic_split.cuh)
// CTAs per chain: the largest power of two <= 8 that still leaves every CTA LSP_MIN_ROWS minibatch rows and fits chains x CTAs on the
// device; 1 = keep one CTA per chain (many chains: the SMs are busy anyway).  opt: 0 = that rule, 1 = never, 2 / 4 / 8 = that cluster size
constexpr int LSP_MIN_ROWS = 16;
static int split_of(int opt, int chains, int n, int sms, int ld) {
    if (opt == 1) return 1;
    if (opt >= 2) { int split = 1; while (split * 2 <= opt) split *= 2; if (n / split < 1) split = 1; return split; }
    int split = 1;
    for (int s = 2; s <= LSP_MAX_SPLIT; s *= 2)
        if ((int64_t)chains * s <= sms && n / s >= LSP_MIN_ROWS) split = s;
    // short rows: a cluster of 2 spends about as many issue slots per step as the short-row kernel does on one SM (G lanes per row
    // against a lane per row); the split pays from 4 CTAs on.  Long rows are bound by one SM's load path: 2 CTAs already halve it
    if (split == 2 && ld <= 128) split = 1;
    return split;
}
// the cluster size a launch of this handle would use (1: one CTA per chain).  Same rule for the mode finder ("mode_finder_split") and
// the sampling kernels ("chain_split"): a few chains are the reference's way of running -- they get a cluster each --, many chains
// are this library's, and then every SM already has a chain of its own
static int logistic_split_for(const sgm_handle* h, const ChainArgs& a, int sms) {
    if (a.X != h->X) return 1;
    const bool finder = a.algo == ALGO_OPT && !a.cv && a.L == 1;
    const int opt = finder ? h->opt_mode_finder_split : h->opt_chain_split;
    if (opt == 0 && (h->opt_logistic_kernel || h->opt_gather)) return 1;   // a pinned one-CTA kernel path is what the caller wants to run
    return split_of(opt, a.C, a.n, sms, h->ld);
}

// *launched = false: the device cannot co-schedule such a cluster (or refuses the launch configuration) -- the caller falls back to
// one CTA per chain; nothing has run in that case
template <int G, int V>
static int launch_logistic_split_gv(sgm_handle* h, const ChainArgs& a, int split, int sms, bool* launched) {
    *launched = false;
    if (h->lsp_asked == split && h->lsp_n == a.n) {
        split = h->lsp_granted;        // asked before (a step-by-step loop launches once per sgmcmcStep: no occupancy query per step)
    } else {
        const int asked = split;
        for (; split > 1; split /= 2) {   // the largest cluster the device can hold
            const size_t smem = lsp_smem_bytes(G, V, h->ld, a.n, split);
            if (smem > 227 * 1024) continue;
            CU(cudaFuncSetAttribute((logistic_split_kernel<G, V>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int fit = 0;
            cudaError_t qe;
            SGM_CLUSTER_OCCUPANCY(qe, &fit, (logistic_split_kernel<G, V>), (unsigned)split, LSP_THREADS, smem, split);
            if (qe == cudaSuccess && fit > 0) break;
            cudaGetLastError();
        }
        h->lsp_asked = asked; h->lsp_n = a.n; h->lsp_granted = split;
    }
    if (split <= 1) return SGM_OK;
    const size_t smem = lsp_smem_bytes(G, V, h->ld, a.n, split);
    CU(cudaFuncSetAttribute((logistic_split_kernel<G, V>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per function, not per handle: set for THIS launch
    int clusters = sms / split;
    if (clusters > a.C) clusters = a.C;
    if (clusters < 1) clusters = 1;
    cudaError_t err;
    SGM_LAUNCH_CLUSTER(err, (logistic_split_kernel<G, V>), (unsigned)(clusters * split), LSP_THREADS, smem, h->stream, split, a);
    if (err != cudaSuccess) { cudaGetLastError(); return SGM_OK; }   // a refused configuration is not a fault: nothing was launched
    CU(cudaGetLastError());
    *launched = true;
    return SGM_OK;
}

static int launch_logistic_split(sgm_handle* h, const ChainArgs& a, int split, int sms, bool* launched) {
    LOGI_DISPATCH(launch_logistic_split_gv, h, a, split, sms, launched);
}

static int launch_logistic(sgm_handle* h, const ChainArgs& a) {
    const_cast<ChainArgs&>(a).rhat_cached = h->rhat_cached ? 1 : 0;
    {
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
        const int split = logistic_split_for(h, a, sms);
        LogiShape s;
        if (split > 1 && logi_shape(h->ld, &s)) {
            bool launched = false;
            const int rc = launch_logistic_split(h, a, split, sms, &launched);
            if (rc != SGM_OK || launched) return rc;
        }
    }
    if (logistic_use_staged(h, a)) return launch_logistic_staged(h, a);
    LogiShape s;
    if (!logi_shape(h->ld, &s)) return fail(SGM_EINVAL, "logistic: d = %d too large", h->d);
    const int threads = logi_threads(h, s);
    const size_t smem = logi_smem_bytes(threads, s, a.n);
    if (smem > 227 * 1024) return fail(SGM_EINVAL, "logistic: minibatch %d needs %zu B of shared memory", a.n, smem);
    LOGI_DISPATCH(launch_logistic_gv, h, a, threads, smem);
}

template <bool LONG, int V>
static int launch_fullgrad_stream_t(sgm_handle* h, int64_t row0, int64_t row1, const float* theta_hat) {
    auto k = logistic_fullgrad_stream_kernel<LONG, V>;
    const int tile_rows = LONG ? FG_LONG_ROWS : FG_SHORT_ROWS;
    const size_t smem = fg_smem_bytes(h->ld, tile_rows);
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0, per_sm = 1;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, FG_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = cdiv(row1 - row0, tile_rows);
    const int grid = (int)(n_tiles < (int64_t)sms * per_sm ? n_tiles : (int64_t)sms * per_sm);
    if (grid < 1) return SGM_OK;
    SGM_LAUNCH(k, grid, FG_THREADS, smem, h->stream, h->X, h->ld, h->d, row0, row1, tile_rows, (int)n_tiles, theta_hat, h->gsum);
    CU(cudaGetLastError());
    return SGM_OK;
}

// ---- many modes: two tensor-core contractions per row block (logistic_cv_gemm.cuh)
static int cvg_encode_2d(CUtensorMap* map, const float* base, uint64_t inner, uint64_t rows, uint64_t pitch_floats, uint32_t box_inner = CVG_KC,
                         uint32_t box_rows = CVG_TILE, bool swizzle128 = true) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { cudaGetLastError(); return -1; }
        fn = (EncodeFn)p;
    }
    const cuuint64_t gdim[2] = {inner, rows};
    const cuuint64_t gstride[1] = {pitch_floats * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

static bool logistic_fullgrad_use_tc(const sgm_handle* h, int M) {
    if (h->opt_fullgrad == 1) return M >= 1 && (h->Pst % 4) == 0 && (h->ld % 4) == 0;
    if (h->opt_fullgrad == 2) return false;
    return M >= 16 && (h->Pst % 4) == 0 && (h->ld % 4) == 0;   // below that the SIMT row-chunk x mode grid is HBM-bound anyway
}

// ---- launch plan of the many-mode pass: every number the launches and tensor maps use, computed without touching the device.
// launch_logistic_fullgrad_tc() executes exactly this plan; sgm_debug_plan() prints and checks it for any configuration
// (full BASELINE sizes, any world size, any shard start) in a process that has no GPU.
struct TmapPlan {          // a 2-D float32 tensor map; box (CVG_KC x CVG_TILE) with the 128-byte swizzle unless stated otherwise
    const char* what;
    uint64_t inner, rows, pitch_floats;      // globalDim[0], globalDim[1], globalStrides[0] / 4
    uint64_t alloc_floats;                   // floats available behind the base pointer
    uint32_t box_inner = CVG_KC, box_rows = CVG_TILE;
    bool swizzle128 = true;
};
struct CvgBlockPlan {
    int64_t b0, b1;        // global rows of the block
    CvgArgs p1, p2;        // kernel arguments of the two passes (device pointers left null)
    int grid1, grid2;
    int64_t rowsum_warps;
    TmapPlan mRt;
};
struct CvgPlan {
    int64_t blk = 0, pitch = 0;
    int shift = 0, tiles_m = 0;              // shift: always 0 (kept in the arithmetic: the column of a block's first row in R^T)
    int a_box_w = 0;                         // pass 2: features per row of the A tile (<= 128)
    size_t rt_bytes = 0;
    TmapPlan mX{}, mTh{}, mX2{};
    std::vector<CvgBlockPlan> blocks;
};

static CvgPlan cvg_plan(int64_t N, int d, int ld, int Pst, int M, int64_t row0, int64_t row1, int sms) {
    CvgPlan P;
    const int64_t rows_all = row1 - row0;
    // row blocks: R^T of one block stays <= ~2 GB
    const int Mpad = (int)cdiv(M, CVG_TILE) * CVG_TILE;
    int64_t blk = ((int64_t)1 << 29) / Mpad;
    blk = blk / 256 * 256; if (blk < 256) blk = 256;
    if (blk > rows_all) blk = cdiv(rows_all, 256) * 256;
    P.blk = blk;
    P.shift = 0;                  // pass 2's K coordinate is a ROW of X (outer dimension of its tensor map) and a block-relative column of R^T: no alignment rule binds the row a range starts at
    P.a_box_w = ld < CVG_TILE ? ld : CVG_TILE;
    P.pitch = blk + 32;
    P.rt_bytes = (size_t)M * P.pitch * sizeof(float);
    P.tiles_m = (int)cdiv(M, CVG_TILE);
    P.mX = TmapPlan{"X (pass 1, A)", (uint64_t)d, (uint64_t)N, (uint64_t)ld, (uint64_t)N * ld};
    P.mTh = TmapPlan{"theta_hat (pass 1, B)", (uint64_t)d, (uint64_t)M, (uint64_t)Pst, (uint64_t)M * Pst};
    P.mX2 = TmapPlan{"X (pass 2, A: 32 data rows x features, no swizzle)", (uint64_t)d, (uint64_t)N, (uint64_t)ld, (uint64_t)N * ld, (uint32_t)P.a_box_w, (uint32_t)CVG_KC, false};
    for (int64_t b0 = row0; b0 < row1; b0 += blk) {
        CvgBlockPlan B{};
        const int64_t b1 = (b0 + blk < row1) ? b0 + blk : row1;
        const int64_t rows = b1 - b0;
        B.b0 = b0; B.b1 = b1;
        B.mRt = TmapPlan{"R^T (pass 2, B)", (uint64_t)(rows + P.shift), (uint64_t)M, (uint64_t)P.pitch, (uint64_t)M * P.pitch};
        CvgArgs a{};
        a.ld = ld; a.d = d; a.Pst = Pst; a.M = M; a.rt_pitch = P.pitch; a.rt_col0 = P.shift;
        a.row_base = b0; a.row_end = b1; a.tiles_b = P.tiles_m;
        // pass 1: R^T = y - sigmoid(X theta_hat^T + bias)
        a.epi = 0; a.k_chunks = (int)cdiv(d, CVG_KC); a.a_row0 = (int)b0; a.a_k0 = 0; a.b_k0 = 0; a.k_per_split = 0; a.k_end_a = d;
        a.tiles_a = (int)cdiv(rows, CVG_TILE);
        a.n_items = a.tiles_a * P.tiles_m;
        B.grid1 = a.n_items < sms ? a.n_items : sms;
        B.rowsum_warps = (d % CVG_TILE) != 0 ? 0 : (int64_t)M * cdiv(rows + P.shift, 4096);   // 0: pass 2 sums the bias component itself
        B.p1 = a;
        // pass 2: gsum[m][k] += sum_i R^T[m][i] X^T[k][i], K (= rows) split over CTAs
        a.epi = 1; a.a_row0 = 0; a.a_box_w = P.a_box_w;
        a.bias_row = (d % CVG_TILE) != 0 ? d : -1;   // a free row of the last feature tile carries the bias sums; a full tile would cost a whole extra one
        const int tiles_k = (int)cdiv(d, CVG_TILE);
        // ~8 work items per CTA (static round-robin: the tail is at most one item), each at least 2 K segments long
        int64_t splits = cdiv((int64_t)sms * 8, (int64_t)tiles_k * P.tiles_m);
        const int64_t kext = rows + P.shift;
        int64_t kps = cdiv(cdiv(kext, splits), CVG_KC * CVG_SEG) * (CVG_KC * CVG_SEG);
        if (kps < 2 * CVG_KC * CVG_SEG) kps = 2 * CVG_KC * CVG_SEG;
        splits = cdiv(kext, kps);
        a.k_per_split = (int)kps; a.k_chunks = (int)(kps / CVG_KC); a.a_k0 = (int)(b0 - P.shift); a.b_k0 = 0; a.k_end_a = (int)b1;
        a.tiles_a = tiles_k;
        a.n_items = (int)(tiles_k * P.tiles_m * splits);
        B.grid2 = a.n_items < sms ? a.n_items : sms;
        B.p2 = a;
        P.blocks.push_back(B);
    }
    return P;
}

// the rules cuTensorMapEncodeTiled and the TMA unit enforce (a violation is CUDA_ERROR_INVALID_VALUE at best, an illegal
// instruction -- Xid 13 -- at worst), applied to a plan before anything is launched
static void tmap_plan_check(const TmapPlan& m, std::vector<std::string>& bad) {
    char buf[256];
    auto say = [&](const char* why) { snprintf(buf, sizeof buf, "tensor map %s: %s (dims %llu x %llu, pitch %llu floats)", m.what, why,
                                               (unsigned long long)m.inner, (unsigned long long)m.rows, (unsigned long long)m.pitch_floats); bad.push_back(buf); };
    if (m.inner == 0 || m.rows == 0 || m.inner > (1ull << 32) || m.rows > (1ull << 32)) say("globalDim outside [1, 2^32]");
    if ((m.pitch_floats * 4) % 16 != 0) say("globalStrides[0] is not a multiple of 16 bytes");
    if (m.pitch_floats * 4 >= (1ull << 40)) say("globalStrides[0] >= 2^40");
    if (m.pitch_floats < m.inner) say("row pitch shorter than the row");
    if (m.rows > 0 && (m.rows - 1) * m.pitch_floats + m.inner > m.alloc_floats) say("tensor extends past its allocation");
    if (m.box_inner < 1 || m.box_inner > 256 || m.box_rows < 1 || m.box_rows > 256) say("box dimension outside [1, 256]");
    if ((m.box_inner * 4) % 16 != 0) say("inner box extent is not a multiple of 16 bytes");
    if (m.swizzle128 && m.box_inner * 4 > 128) say("inner box extent exceeds the 128-byte swizzle span");
    static_assert(CVG_KC * 4 == 128 && CVG_TILE <= 256, "box = one 128-byte swizzle span x <= 256 rows");
}

static std::vector<std::string> cvg_plan_check(const CvgPlan& P) {
    std::vector<std::string> bad;
    char buf[256];
    tmap_plan_check(P.mX, bad); tmap_plan_check(P.mTh, bad); tmap_plan_check(P.mX2, bad);
    if (CVG_SMEM > 227 * 1024) bad.push_back("cvg_kernel: more than 227 KB of shared memory");
    if ((size_t)P.a_box_w * CVG_KC * 4 > CVG_OP_BYTES) bad.push_back("cvg pass 2: the A tile exceeds its stage");
    for (const CvgBlockPlan& B : P.blocks) {
        tmap_plan_check(B.mRt, bad);
        for (const CvgArgs* a : {&B.p1, &B.p2}) {
            // innermost TMA coordinates: pass 1 a_k0 / b_k0 + chunk * CVG_KC for both operands; pass 2 b_k0 + split * k_per_split + chunk * CVG_KC for
            // R^T (the A operand's K coordinate is a row of X there: any integer) -- every term must be a multiple of 4 floats
            if ((a->epi == 0 && (a->a_k0 & 3)) || (a->b_k0 & 3) || (a->k_per_split & 3)) {
                snprintf(buf, sizeof buf, "cvg pass %d, rows [%lld, %lld): TMA box start not 16-byte aligned (a_k0 %d, b_k0 %d, K per split %d)", a->epi + 1,
                         (long long)B.b0, (long long)B.b1, a->a_k0, a->b_k0, a->k_per_split);
                bad.push_back(buf);
            }
            if (a->n_items < 1 || a->k_chunks < 1 || a->tiles_a < 1 || a->tiles_b < 1) {
                snprintf(buf, sizeof buf, "cvg pass %d, rows [%lld, %lld): empty work decomposition (%d items, %d chunks)", a->epi + 1, (long long)B.b0, (long long)B.b1, a->n_items, a->k_chunks);
                bad.push_back(buf);
            }
            if (a->epi == 1 && (a->k_per_split % (CVG_KC * CVG_SEG)) != 0) bad.push_back("cvg pass 2: K per split is not a whole number of TMEM segments");
        }
        // pass 2 must cover [b0 - shift, b1) exactly once: splits x k_per_split >= extent, a_k0 + extent == b1
        const int64_t splits = B.p2.n_items / (B.p2.tiles_a * B.p2.tiles_b);
        if ((int64_t)B.p2.a_k0 + splits * B.p2.k_per_split < B.b1 || (int64_t)B.p2.a_k0 != B.b0 - P.shift) bad.push_back("cvg pass 2: the K walk does not cover the row block");
        if (B.b1 > (int64_t)INT32_MAX) bad.push_back("cvg: row coordinates beyond int32");
        if (P.shift + (B.b1 - B.b0) > P.pitch) bad.push_back("cvg: R^T pitch shorter than the block");
    }
    if (P.pitch & 3) bad.push_back("cvg: R^T pitch not a multiple of 4 floats");
    return bad;
}

static int launch_logistic_fullgrad_tc(sgm_handle* h, int64_t row0, int64_t row1, int M, const float* theta_hat) {
    const int d = h->d;
    const int64_t rows_all = row1 - row0;
    if (rows_all <= 0) return SGM_OK;
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const CvgPlan P = cvg_plan(h->N, d, h->ld, h->Pst, M, row0, row1, sms);
    {   // a plan that breaks a TMA rule is refused here, on the host: on the device it is an illegal instruction
        const std::vector<std::string> bad = cvg_plan_check(P);
        if (!bad.empty()) return fail(SGM_EINVAL, "logistic full gradient: %s%s", bad[0].c_str(), bad.size() > 1 ? " (and more)" : "");
    }
    CU(cudaFuncSetAttribute(cvg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CVG_SMEM));
    if (P.rt_bytes > h->rt_cap) {
        if (h->Rt) cudaFree(h->Rt);
        h->Rt = nullptr; h->rt_cap = 0;
        CU(cudaMalloc(&h->Rt, P.rt_bytes));
        h->rt_cap = P.rt_bytes;
    }
    CUtensorMap mX, mTh, mX2, mRt;
    if (cvg_encode_2d(&mX, h->X, P.mX.inner, P.mX.rows, P.mX.pitch_floats) || cvg_encode_2d(&mTh, theta_hat, P.mTh.inner, P.mTh.rows, P.mTh.pitch_floats) ||
        cvg_encode_2d(&mX2, h->X, P.mX2.inner, P.mX2.rows, P.mX2.pitch_floats, P.mX2.box_inner, P.mX2.box_rows, P.mX2.swizzle128))
        return fail(SGM_ECUDA, "logistic full gradient: cuTensorMapEncodeTiled failed");
    for (const CvgBlockPlan& B : P.blocks) {
        const int64_t rows = B.b1 - B.b0;
        if (cvg_encode_2d(&mRt, h->Rt, B.mRt.inner, B.mRt.rows, B.mRt.pitch_floats)) return fail(SGM_ECUDA, "logistic full gradient: tensor map (R^T)");
        CvgArgs a = B.p1;
        a.X = h->X; a.theta_hat = theta_hat; a.Rt = h->Rt; a.gsum = h->gsum;
        SGM_LAUNCH(cvg_kernel, B.grid1, CVG_THREADS, CVG_SMEM, h->stream, a, mX, mTh);
        CU(cudaGetLastError());
        if (B.rowsum_warps > 0) {
            SGM_LAUNCH(cvg_rowsum_kernel, (unsigned)cdiv(B.rowsum_warps * 32, 256), 256, 0, h->stream, h->Rt, P.pitch, rows + P.shift, M, h->Pst, d, h->gsum);
            CU(cudaGetLastError());
        }
        a = B.p2;
        a.X = h->X; a.theta_hat = theta_hat; a.Rt = h->Rt; a.gsum = h->gsum;
        SGM_LAUNCH(cvg_kernel, B.grid2, CVG_THREADS, CVG_SMEM, h->stream, a, mX2, mRt);
        CU(cudaGetLastError());
        h->last_launches += B.rowsum_warps > 0 ? 3 : 2;
    }
    h->fullgrad_flops = 4.0 * (double)rows_all * (double)d * (double)M;
    return SGM_OK;
}

// one mode: streaming kernel (TMA bulk tile ring); per-chain modes: row-chunk x mode grid
static int launch_logistic_fullgrad(sgm_handle* h, int64_t row0, int64_t row1, int M, const float* theta_hat) {
    const int Q = h->ld / 4;
    h->fullgrad_flops = 0;
    if (logistic_fullgrad_use_tc(h, M)) return launch_logistic_fullgrad_tc(h, row0, row1, M, theta_hat);
    if (M == 1 && h->opt_fullgrad != 2 && cdiv(row1 - row0, FG_SHORT_ROWS) < (1ll << 31)) {
        if (h->ld <= 128) {
            // lane-per-row needs conflict-free shared rows: odd float4 stride, or 2 (mod 4) with skewed lanes; <= 4 column slots per warp
            if (((Q & 1) || (Q % 4 == 2)) && fg_smem_bytes(h->ld, FG_SHORT_ROWS) <= 113 * 1024)
                return launch_fullgrad_stream_t<false, 1>(h, row0, row1, theta_hat);
        } else if (fg_smem_bytes(h->ld, FG_LONG_ROWS) <= 113 * 1024) {
            const int V = (Q + 31) / 32;
            switch (V) {
                case 2: return launch_fullgrad_stream_t<true, 2>(h, row0, row1, theta_hat);
                case 3: return launch_fullgrad_stream_t<true, 3>(h, row0, row1, theta_hat);
                case 4: return launch_fullgrad_stream_t<true, 4>(h, row0, row1, theta_hat);
                case 5: return launch_fullgrad_stream_t<true, 5>(h, row0, row1, theta_hat);
                case 6: return launch_fullgrad_stream_t<true, 6>(h, row0, row1, theta_hat);
                case 7: return launch_fullgrad_stream_t<true, 7>(h, row0, row1, theta_hat);
                case 8: return launch_fullgrad_stream_t<true, 8>(h, row0, row1, theta_hat);
                case 9: return launch_fullgrad_stream_t<true, 9>(h, row0, row1, theta_hat);
            }
        }
    }
    LOGI_DISPATCH(launch_logistic_fullgrad_gv, h, row0, row1, M, theta_hat);
}

// Gaussian families: dataset small enough to live in shared memory (and a whole number of 16-byte TMA units)?
static bool small_use_smem(const sgm_handle* h, const ChainArgs& a) {
    if (h->opt_small_kernel == 1) return false;
    const size_t bytes = (size_t)a.N * a.ld * sizeof(float);
    // staging the whole dataset pays off once a launch gathers more than it stages: not for one step of one chain (sgmcmcStep)
    const bool enough_work = (int64_t)a.C * (a.iters > 0 ? a.iters : 1) * a.L * a.n * a.ld * 4 >= (int64_t)bytes;
    return a.X == h->X && bytes % 16 == 0 && bytes <= SMALL_SMEM_MAX_BYTES && (enough_work || h->opt_small_kernel == 2);
}

// one launch of the family's chain kernel
static int launch_chain(sgm_handle* h, const ChainArgs& a) {
    h->last_launches += 1;
    switch (h->model) {
        case SGM_MODEL_GAUSS_MEAN:
        case SGM_MODEL_MIXTURE2: {
            int sms = 0;
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
            const bool gauss = h->model == SGM_MODEL_GAUSS_MEAN, narrow = h->d <= 2;   // d <= 2 (the vignettes): register arrays of 2, not 8, entries per tensor
            if (small_use_smem(h, a)) {
                // the dataset staged in shared memory once per CTA, 16 chains per CTA
                const size_t smem = small_smem_bytes(a.N, a.ld);
                int64_t grid = cdiv((int64_t)a.C * 32, SMALL_SMEM_THREADS);
                if (grid > (int64_t)sms * 2) grid = (int64_t)sms * 2;
#define SMALL_SMEM_LAUNCH(M, D)                                                                                                          \
    do {                                                                                                                                 \
        CU(cudaFuncSetAttribute((small_chain_kernel<M, true, D>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        SGM_LAUNCH((small_chain_kernel<M, true, D>), (unsigned)grid, SMALL_SMEM_THREADS, smem, h->stream, a);                           \
    } while (0)
                if (gauss && narrow) SMALL_SMEM_LAUNCH(SMALL_MODEL_GAUSS, 2);
                else if (gauss) SMALL_SMEM_LAUNCH(SMALL_MODEL_GAUSS, SMALL_MAXD);
                else if (narrow) SMALL_SMEM_LAUNCH(SMALL_MODEL_MIX2, 2);
                else SMALL_SMEM_LAUNCH(SMALL_MODEL_MIX2, SMALL_MAXD);
#undef SMALL_SMEM_LAUNCH
                CU(cudaGetLastError());
                return SGM_OK;
            }
            const int threads = 128;
            int64_t grid = cdiv((int64_t)a.C * 32, threads);
            if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
            if (gauss && narrow) SGM_LAUNCH((small_chain_kernel<SMALL_MODEL_GAUSS, false, 2>), (unsigned)grid, threads, 0, h->stream, a);
            else if (gauss) SGM_LAUNCH((small_chain_kernel<SMALL_MODEL_GAUSS, false, SMALL_MAXD>), (unsigned)grid, threads, 0, h->stream, a);
            else if (narrow) SGM_LAUNCH((small_chain_kernel<SMALL_MODEL_MIX2, false, 2>), (unsigned)grid, threads, 0, h->stream, a);
            else SGM_LAUNCH((small_chain_kernel<SMALL_MODEL_MIX2, false, SMALL_MAXD>), (unsigned)grid, threads, 0, h->stream, a);
            CU(cudaGetLastError());
            return SGM_OK;
        }
        case SGM_MODEL_LOGISTIC:
            return launch_logistic(h, a);
        case SGM_MODEL_SOFTMAX_NET: {
            int launches = 0;
            const int rc = softmax_net_launch(h->nn, a, h->stream, &launches);
            h->last_launches += launches - 1;
            if (rc != 0) return fail(SGM_ECUDA, "softmax net launch failed: %s", cudaGetErrorString((cudaError_t)rc));
            return SGM_OK;
        }
    }
    return fail(SGM_ENOMODEL, "unknown model %d", h->model);
}

static int ensure_buf(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return SGM_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CU(cudaMalloc(p, bytes));
    *cap = bytes;
    return SGM_OK;
}

static int check_diverged(sgm_handle* h, const char* msg) {
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (flag) {
        CU(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
        return fail(SGM_EDIVERGED, "%s", msg);
    }
    return SGM_OK;
}

struct Timer {
    sgm_handle* h;
    explicit Timer(sgm_handle* hh) : h(hh) { h->last_launches = 0; cudaEventRecord(h->ev0, h->stream); }
    int stop() {
        CU(cudaEventRecord(h->ev1, h->stream));
        CU(cudaEventSynchronize(h->ev1));
        CU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
        return SGM_OK;
    }
};


// ------------------------------------------------------------------------------------
// dataset upload: the caller's double matrix (column-major as R holds it, or row-major) -> float rows on the device
// ------------------------------------------------------------------------------------
// A cudaMemcpy from pageable memory is staged by the driver on one thread, and it would ship doubles.  Instead host threads convert
// slabs of rows to float32 straight into page-locked staging (half the PCIe bytes; (float)x on the host is the device's
// __double2float_rn), two slabs in flight: slab k + 1 is converted while slab k crosses the bus and is packed into the padded rows.
// The staging is process-level (2 x 32 MB, page-locked once: like the CUDA context it is paid for by the first sampler of a process).
namespace {
constexpr size_t UP_SLAB_BYTES = 32u << 20;
struct UploadStaging {
    std::mutex mu;
    float* host[2] = {nullptr, nullptr};
    bool tried = false;
};
UploadStaging g_upload;

// rows [r0, r1) of the slab that starts at row slab_r0 and holds slab_rows rows: narrowed in the caller's layout (contiguous reads AND
// writes either way: on the development host 48 ms for the c3 matrix at 8 threads, against 85 ms when the host also transposed)
void convert_rows(const double* A, int64_t N, int cols, int row_major, int64_t r0, int64_t r1, int64_t slab_r0, int64_t slab_rows, float* out) {
    if (row_major) {
        const double* s_ = A + r0 * cols;
        float* d_ = out + (r0 - slab_r0) * cols;
        const int64_t cnt = (r1 - r0) * cols;
        for (int64_t i = 0; i < cnt; ++i) d_[i] = (float)s_[i];
        return;
    }
    for (int c = 0; c < cols; ++c) {
        const double* s_ = A + (size_t)c * N + r0;
        float* d_ = out + (size_t)c * slab_rows + (r0 - slab_r0);
        for (int64_t i = 0; i < r1 - r0; ++i) d_[i] = (float)s_[i];
    }
}
}  // namespace

// entry A ([N][cols] doubles on the host) -> dst rows [0, N) at column col0 (row stride ld floats), on `stream`; returns a cudaError_t
static cudaError_t upload_entry(const double* A, int64_t N, int cols, int row_major, float* dst, int ld, int col0, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_upload.mu);
    if (!g_upload.tried) {
        g_upload.tried = true;
        for (int b = 0; b < 2; ++b)
            if (cudaHostAlloc(&g_upload.host[b], UP_SLAB_BYTES, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); g_upload.host[b] = nullptr; }
    }
    float* dev[2] = {nullptr, nullptr};
    cudaEvent_t packed[2] = {nullptr, nullptr};
    const bool staged = g_upload.host[0] && g_upload.host[1];
    cudaError_t e = cudaSuccess;
    const size_t dev_slab = std::min<size_t>(UP_SLAB_BYTES, (size_t)N * cols * sizeof(float));   // a small entry needs a small slab
    if (staged) {
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) { e = cudaMalloc(&dev[b], dev_slab); if (e == cudaSuccess) e = cudaEventCreate(&packed[b]); }
    }
    if (!staged || e != cudaSuccess) {
        // no page-locked staging (or no room for the two device slabs): the whole entry as doubles, converted by a kernel
        cudaGetLastError();
        for (int b = 0; b < 2; ++b) { if (dev[b]) cudaFree(dev[b]); if (packed[b]) cudaEventDestroy(packed[b]); }
        double* stage = nullptr;
        const size_t bytes = (size_t)N * cols * sizeof(double);
        if ((e = cudaMalloc(&stage, bytes)) != cudaSuccess) return e;
        e = cudaMemcpyAsync(stage, A, bytes, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) { SGM_LAUNCH(pack_rows_kernel, (unsigned)cdiv(N * cols, 256), 256, 0, stream, stage, N, cols, dst, ld, col0, row_major); e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(stage);
        return e;
    }
    size_t slab_bytes = UP_SLAB_BYTES;
#ifdef SGM_EMU   // the CPU-emulated test build shrinks the slabs so that toy datasets take several (double buffering, threads, tails)
    if (const char* sb = getenv("SGM_EMU_UPLOAD_SLAB")) { const size_t v = (size_t)atoll(sb); if (v >= 64 && v <= UP_SLAB_BYTES) slab_bytes = v; }
#endif
    const int64_t slab_rows = std::max<int64_t>(1, (int64_t)(slab_bytes / sizeof(float)) / cols);
    unsigned hw = std::thread::hardware_concurrency();
    const int T = (int)(hw ? (hw > 8 ? 8 : hw) : 1);
    int k = 0;
    for (int64_t r0 = 0; r0 < N && e == cudaSuccess; r0 += slab_rows, ++k) {
        const int b = k & 1;
        const int64_t r1 = r0 + slab_rows < N ? r0 + slab_rows : N, rows = r1 - r0;
        if (k >= 2) e = cudaEventSynchronize(packed[b]);          // the slab that used these buffers two rounds ago has left them
        if (e != cudaSuccess) break;
        if ((rows * cols < (1 << 16) && slab_bytes == UP_SLAB_BYTES) || T == 1 || rows < T) convert_rows(A, N, cols, row_major, r0, r1, r0, rows, g_upload.host[b]);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back(convert_rows, A, N, cols, row_major, r0 + rows * t / T, r0 + rows * (t + 1) / T, r0, rows, g_upload.host[b]);
            for (auto& x : th) x.join();
        }
        e = cudaMemcpyAsync(dev[b], g_upload.host[b], (size_t)rows * cols * sizeof(float), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) break;
        SGM_LAUNCH(pack_slab_kernel, (unsigned)cdiv(rows * cols, 256), 256, 0, stream, (const float*)dev[b], rows, cols, row_major, dst, r0, ld, col0);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(packed[b], stream);
    }
    const cudaError_t es = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = es;
    for (int b = 0; b < 2; ++b) { cudaFree(dev[b]); cudaEventDestroy(packed[b]); }
    return e;
}

// ------------------------------------------------------------------------------------
// create
// ------------------------------------------------------------------------------------
static int64_t tensor_size(const sgm_param& p) {
    int64_t s = 1;
    for (int i = 0; i < p.ndim; ++i) s *= p.dims[i];
    return s;
}

// Every chain at the user's start values, thermostats at `a`, momenta zero: what tf$global_variables_initializer() leaves behind
// (initSess, R/update.R:126-131).  Called by sgm_create and again by every sgm_init, so a second initSess restarts the sampler.
static int upload_initial_state(sgm_handle* h) {
    const size_t sbytes = (size_t)h->C * h->Pst * sizeof(float);
    std::vector<float> all((size_t)h->C * h->Pst);
    for (int c = 0; c < h->C; ++c) std::memcpy(all.data() + (size_t)c * h->Pst, h->init_state.data(), h->Pst * sizeof(float));
    CU(cudaMemcpyAsync(h->theta, all.data(), sbytes, cudaMemcpyHostToDevice, h->stream));
    if (h->cv) CU(cudaMemcpyAsync(h->theta_hat, all.data(), sbytes, cudaMemcpyHostToDevice, h->stream));
    std::vector<float> th((size_t)h->C * MAX_TENSORS, 0.f);
    for (int c = 0; c < h->C; ++c) for (int t = 0; t < h->T; ++t) th[(size_t)c * MAX_TENSORS + t] = (float)h->a[t];
    CU(cudaMemcpyAsync(h->thermo, th.data(), th.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->mom, 0, sbytes, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // the host vectors above go out of scope
    return SGM_OK;
}

extern "C" int sgm_create(const sgm_config* cfg, sgm_handle** out) {
    if (!cfg || !out) return fail(SGM_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->abi_version != SGM_ABI_VERSION) return fail(SGM_EINVAL, "ABI version mismatch (%d != %d)", cfg->abi_version, SGM_ABI_VERSION);
    if (cfg->algo < SGM_ALGO_SGLD || cfg->algo > SGM_ALGO_SGNHT) return fail(SGM_EINVAL, "unknown algorithm %d", cfg->algo);
    if (cfg->n_chains < 1) return fail(SGM_EINVAL, "n_chains must be >= 1");
    if (cfg->n_params < 1 || cfg->n_params > MAX_TENSORS || !cfg->params) return fail(SGM_EINVAL, "bad params list");
    if (cfg->model < SGM_MODEL_GAUSS_MEAN || cfg->model > SGM_MODEL_SOFTMAX_NET)
        return fail(SGM_ENOMODEL, "logLik/logPrior must be one of the routed model families "
                                  "(gaussMean, mixture2, logistic, softmaxNet); arbitrary closures cannot run on the device");

    sgm_handle* h = new sgm_handle();
    auto bail = [&](int rc) { sgm_destroy(h); return rc; };
    // device failures inside sgm_create release everything the handle owns (dataset included) before returning
#define CUB(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return bail(fail(SGM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
    h->model = cfg->model; h->algo = cfg->algo; h->cv = cfg->cv ? 1 : 0; h->C = cfg->n_chains;
    h->device = cfg->device; h->flags = cfg->flags; h->chain_offset = cfg->chain_offset; h->use_prior = cfg->use_prior ? 1 : 0;
    h->prior_sd = cfg->prior_sd; h->opt_stepsize = cfg->opt_stepsize; h->n_iters_opt = cfg->n_iters_opt;
    h->L = (cfg->algo == SGM_ALGO_SGHMC) ? cfg->L : 1;
    if (h->L < 1) return bail(fail(SGM_EINVAL, "L must be >= 1"));
    if (cfg->seeded) h->seed = cfg->seed;
    else h->seed = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count() * 0x9E3779B97F4A7C15ull;

    // ---- parameter tensors
    int64_t off = 0;
    for (int t = 0; t < cfg->n_params; ++t) {
        const sgm_param& p = cfg->params[t];
        if (p.ndim < 0 || p.ndim > 4 || !p.init) return bail(fail(SGM_EINVAL, "parameter %d: bad shape or null init", t));
        TensorInfo ti; ti.name = p.name ? p.name : ""; ti.ndim = p.ndim;
        for (int i = 0; i < 4; ++i) ti.dims[i] = i < p.ndim ? p.dims[i] : 1;
        ti.size = tensor_size(p); ti.offset = off; off += ti.size;
        if (!(p.stepsize > 0)) return bail(fail(SGM_EINVAL, "parameter %s: stepsize must be > 0", ti.name.c_str()));
        h->tensors.push_back(ti);
        h->eps.push_back(p.stepsize); h->alpha.push_back(p.alpha); h->a.push_back(p.a);
    }
    h->T = cfg->n_params; h->P = (int)off;

    // ---- dataset shape per family
    const bool synth = cfg->synth_rows > 0;
    auto need_data = [&](int k) -> bool { return cfg->n_data >= k && cfg->data && (synth || [&] { for (int i = 0; i < k; ++i) if (!cfg->data[i].data) return false; return true; }()); };
    int ycols = 0;
    switch (h->model) {
        case SGM_MODEL_GAUSS_MEAN:
        case SGM_MODEL_MIXTURE2: {
            if (synth || !need_data(1)) return bail(fail(SGM_EINVAL, "Gaussian families need dataset entry X"));
            const sgm_array& X = cfg->data[0];
            h->N = X.dims[0]; h->d = X.ndim >= 2 ? (int)X.dims[1] : 1;
            if (h->d < 1 || h->d > SMALL_MAXD) return bail(fail(SGM_EINVAL, "Gaussian families support 1 <= d <= %d (got %d)", SMALL_MAXD, h->d));
            const int want_T = h->model == SGM_MODEL_GAUSS_MEAN ? 1 : 2;
            if (h->T != want_T) return bail(fail(SGM_EINVAL, "model expects %d parameter tensor(s), got %d", want_T, h->T));
            for (auto& ti : h->tensors) if (ti.size != h->d) return bail(fail(SGM_EINVAL, "parameter %s must have %d entries", ti.name.c_str(), h->d));
            h->ld = h->d; h->Pst = h->P;
            break;
        }
        case SGM_MODEL_LOGISTIC: {
            if (synth) { h->N = cfg->synth_rows; h->d = (int)cfg->synth_dim; }
            else {
                if (!need_data(2)) return bail(fail(SGM_EINVAL, "logistic needs dataset entries X, y"));
                h->N = cfg->data[0].dims[0]; h->d = cfg->data[0].ndim >= 2 ? (int)cfg->data[0].dims[1] : 1;
                if (cfg->data[1].dims[0] != h->N) return bail(fail(SGM_EINVAL, "X and y disagree on the number of rows"));
            }
            if (h->T != 2 || h->tensors[0].size != 1 || h->tensors[1].size != h->d)
                return bail(fail(SGM_EINVAL, "logistic expects params (bias: scalar, beta: [%d,1])", h->d));
            h->ld = 4 * (h->d / 4 + 1);
            LogiShape s; if (!logi_shape(h->ld, &s)) return bail(fail(SGM_EINVAL, "logistic: d = %d too large", h->d));
            h->Pst = s.G * s.V * 4;
            break;
        }
        case SGM_MODEL_SOFTMAX_NET: {
            if (synth || !need_data(2)) return bail(fail(SGM_EINVAL, "softmax net needs dataset entries X, y"));
            h->N = cfg->data[0].dims[0]; h->d = (int)cfg->data[0].dims[1];
            ycols = (int)cfg->data[1].dims[1];
            if (cfg->data[1].dims[0] != h->N) return bail(fail(SGM_EINVAL, "X and y disagree on the number of rows"));
            if (h->T != 8) return bail(fail(SGM_EINVAL, "softmax net expects params A, B, a, b, lambdaA, lambdaB, lambdaa, lambdab"));
            const int64_t hid = h->tensors[3].size, outc = h->tensors[2].size;
            if (h->tensors[0].size != hid * outc || h->tensors[1].size != (int64_t)h->d * hid || outc != ycols ||
                h->tensors[4].size != 1 || h->tensors[5].size != 1 || h->tensors[6].size != 1 || h->tensors[7].size != 1)
                return bail(fail(SGM_EINVAL, "softmax net: inconsistent parameter shapes"));
            h->ld = 4 * (int)cdiv(h->d + ycols, 4);
            h->Pst = 4 * (int)cdiv(h->P, 4);
            break;
        }
    }
    if (h->N < 1 || h->N >= (1ll << 31)) return bail(fail(SGM_EINVAL, "dataset rows must be in [1, 2^31)"));

    // ---- convertProp (R/setup.R:128-140)
    if (!(cfg->minibatch_size > 0)) return bail(fail(SGM_EINVAL, "minibatchSize must be > 0"));
    if (cfg->minibatch_size < 1) {
        const double r = std::nearbyint(cfg->minibatch_size * (double)h->N);
        if (r == 0) return bail(fail(SGM_EINVAL, "minibatchSize proportion (%g) too small for dataset size (%lld).", cfg->minibatch_size, (long long)h->N));
        h->n = (int)r;
    } else h->n = (int)cfg->minibatch_size;
    if (h->n > h->N) return bail(fail(SGM_EINVAL, "minibatchSize (%d) exceeds dataset size (%lld)", h->n, (long long)h->N));

    if ((h->model == SGM_MODEL_GAUSS_MEAN || h->model == SGM_MODEL_MIXTURE2) && h->use_prior && !(h->prior_sd > 0))
        return bail(fail(SGM_EINVAL, "prior_sd must be > 0 when use_prior is set"));
    if (h->cv && (!(h->opt_stepsize > 0) || h->n_iters_opt < 0))
        return bail(fail(SGM_EINVAL, "*cv samplers need optStepsize > 0 and nItersOpt >= 0"));

    // ---- everything above is host logic; from here on a CUDA device is mandatory (no CPU fallback)
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return bail(fail(SGM_ECUDA, "no CUDA device available: sgmcmc_b200 has no CPU fallback"));
        }
        if (cfg->device < 0 || cfg->device >= ndev) return bail(fail(SGM_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev));
    }
    if (int rc = set_device(h)) return bail(rc);

    // ---- device allocations
    CUB(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUB(cudaEventCreate(&h->ev0)); CUB(cudaEventCreate(&h->ev1));
    const size_t xbytes = (size_t)h->N * h->ld * sizeof(float);
    if (cudaMalloc(&h->X, xbytes) != cudaSuccess) { cudaGetLastError(); return bail(fail(SGM_ENOMEM, "cannot allocate %zu bytes for the dataset", xbytes)); }
    const uint32_t sk0 = (uint32_t)(cfg->synth_seed & 0xffffffffu), sk1 = (uint32_t)(cfg->synth_seed >> 32);
    if (synth) {
        const int64_t warps = h->N;
        SGM_LAUNCH(synth_logistic_kernel, (unsigned)cdiv(warps * 32, 256), 256, 0, h->stream, h->X, h->N, h->d, h->ld, sk0, sk1);
        CUB(cudaGetLastError());
    } else {
        CUB(cudaMemsetAsync(h->X, 0, xbytes, h->stream));
        int col0 = 0;
        const int entries = (h->model == SGM_MODEL_GAUSS_MEAN || h->model == SGM_MODEL_MIXTURE2) ? 1 : 2;
        for (int k = 0; k < entries; ++k) {
            const sgm_array& A = cfg->data[k];
            const int cols = A.ndim >= 2 ? (int)A.dims[1] : 1;
            const cudaError_t se = upload_entry(A.data, h->N, cols, (int)(A.row_major != 0), h->X, h->ld, col0, h->stream);
            if (se == cudaErrorMemoryAllocation) { cudaGetLastError(); return bail(fail(SGM_ENOMEM, "cannot allocate staging for the dataset upload")); }
            if (se != cudaSuccess) return bail(fail(SGM_ECUDA, "dataset upload failed: %s", cudaGetErrorString(se)));
            col0 += cols;
        }
    }
    const size_t sbytes = (size_t)h->C * h->Pst * sizeof(float);
    CUB(cudaMalloc(&h->theta, sbytes)); CUB(cudaMalloc(&h->mom, sbytes)); CUB(cudaMalloc(&h->scratch, sbytes));
    CUB(cudaMalloc(&h->thermo, (size_t)h->C * MAX_TENSORS * sizeof(float)));
    CUB(cudaMalloc(&h->d_flag, sizeof(int))); CUB(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
    CUB(cudaMalloc(&h->dout, (size_t)h->C * sizeof(double)));
    CUB(cudaMemsetAsync(h->mom, 0, sbytes, h->stream));
    if (h->cv) {
        const size_t hb = (size_t)hat_count(h) * h->Pst * sizeof(float);
        CUB(cudaMalloc(&h->theta_hat, sbytes)); CUB(cudaMalloc(&h->g_hat, sbytes));
        CUB(cudaMalloc(&h->gsum, (size_t)h->C * h->Pst * sizeof(double)));
        CUB(cudaMemsetAsync(h->g_hat, 0, sbytes, h->stream));
        (void)hb;
    }
    if (h->model == SGM_MODEL_SOFTMAX_NET) {
        const int hid = (int)h->tensors[3].size, outc = (int)h->tensors[2].size;
        int rc = softmax_net_create(&h->nn, h->d, hid, outc, h->n, h->C, h->device);
        if (rc != 0) return bail(fail(SGM_ECUDA, "softmax net workspace: %s", cudaGetErrorString((cudaError_t)rc)));
    }
    if (h->model == SGM_MODEL_LOGISTIC && h->ld <= 128) h->tmap_ok = (make_row_tensor_map(h) == 0);
    // initial state: every chain starts at params (internal layout); kept on the host so that initSess can restore it
    {
        std::vector<int> flat_of, tensor_of;
        host_tables(h, flat_of, tensor_of);
        std::vector<float> flat((size_t)h->P);
        for (int t = 0; t < h->T; ++t)
            for (int64_t k = 0; k < h->tensors[t].size; ++k) flat[h->tensors[t].offset + k] = (float)cfg->params[t].init[k];
        h->init_state.assign((size_t)h->Pst, 0.f);
        for (int e = 0; e < h->Pst; ++e) if (flat_of[e] >= 0) h->init_state[e] = flat[flat_of[e]];
        if (int rc = upload_initial_state(h)) return bail(rc);
    }

    // ---- kernel argument template
    ChainArgs& a = h->base;
    std::memset(&a, 0, sizeof a);
    a.X = h->X; a.N = (uint32_t)h->N; a.perm_bits = perm_bits((uint32_t)h->N); a.ld = h->ld; a.d = h->d; a.n = h->n;
    a.corr = (float)((double)h->N / (double)h->n);
    a.C = h->C; a.Pst = h->Pst; a.P = h->P; a.T = h->T;
    a.theta = h->theta; a.mom = h->mom; a.thermo = h->thermo; a.theta_hat = h->theta_hat; a.g_hat = h->g_hat;
    a.hat_stride = h->Pst;
    a.algo = h->algo; a.cv = h->cv; a.L = h->L; a.use_prior = h->use_prior;
    a.prior_inv_var = (h->prior_sd > 0) ? (float)(1.0 / (h->prior_sd * h->prior_sd)) : 0.f;
    for (int i = 0; i < SMALL_MAXD; ++i) a.prec[i * SMALL_MAXD + i] = 1.f;
    if (h->model == SGM_MODEL_GAUSS_MEAN && cfg->precision) {
        // column-major d x d; symmetric, so the transpose is immaterial, but index it properly
        std::vector<double> Pm((size_t)h->d * h->d);
        for (int i = 0; i < h->d; ++i) for (int j = 0; j < h->d; ++j) { a.prec[i * SMALL_MAXD + j] = (float)cfg->precision[i + (size_t)j * h->d]; Pm[i * h->d + j] = cfg->precision[i + (size_t)j * h->d]; }
        // log det by Cholesky (precision must be SPD)
        double logdet = 0;
        for (int i = 0; i < h->d; ++i) {
            for (int j = 0; j <= i; ++j) {
                double s = Pm[i * h->d + j];
                for (int k = 0; k < j; ++k) s -= Pm[i * h->d + k] * Pm[j * h->d + k];
                if (i == j) { if (!(s > 0)) return bail(fail(SGM_EINVAL, "precision matrix is not positive definite")); Pm[i * h->d + i] = std::sqrt(s); logdet += 2 * std::log(Pm[i * h->d + i]); }
                else Pm[i * h->d + j] = s / Pm[j * h->d + j];
            }
        }
        a.logdet_prec = (float)logdet;
    }
    a.opt_lr = (float)h->opt_stepsize;
    for (int t = 0; t < h->T; ++t) {
        const double e = h->eps[t];
        a.tc.half_eps[t] = (float)(0.5 * e); a.tc.eps[t] = (float)e; a.tc.sqrt_2eps[t] = (float)std::sqrt(2 * e);
        a.tc.alpha[t] = (float)h->alpha[t]; a.tc.sqrt_alpha[t] = (float)std::sqrt(h->alpha[t] > 0 ? h->alpha[t] : 0);
        a.tc.sqrt_a[t] = (float)std::sqrt(h->a[t] > 0 ? h->a[t] : 0); a.tc.sqrt_eps[t] = (float)std::sqrt(e);
        a.tc.offset[t] = (int)h->tensors[t].offset;
    }
    a.tc.offset[h->T] = h->P;
    a.k0 = (uint32_t)(h->seed & 0xffffffffu); a.k1 = (uint32_t)(h->seed >> 32);
    a.index_tag = TAG_INDEX;
    a.chain0 = (uint32_t)h->chain_offset;
    a.noise_len = (h->algo == SGM_ALGO_SGHMC) ? (h->L + 1) * h->P : h->P;
    a.diverged = h->d_flag;

    *out = h;
    return SGM_OK;
#undef CUB
}

extern "C" void sgm_destroy(sgm_handle* h) {
    if (!h) return;
    if (h->stream) cudaSetDevice(h->device);
    if (h->comm) {   // the communicator itself lives until the process exits: a unique id must never be used for a second ncclCommInitRank
        auto it = g_comms.find(h->comm_key);
        if (it != g_comms.end()) --it->second.refs;
    }
    if (h->nn) softmax_net_destroy(h->nn);
    for (int r = 0; r < 16; ++r) if (h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
    if (h->d_shard_ptr) cudaFree((void*)h->d_shard_ptr);
    void* ptrs[] = {h->X, h->theta, h->mom, h->thermo, h->theta_hat, h->g_hat, h->scratch, h->gsum, h->dout,
                    h->d_flag, h->d_idx, h->d_noise, h->store, h->Rt, h->d_flat_of, h->d_state_out, h->hat_all, h->gsum_all};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int b = 0; b < 2; ++b) {
        if (h->pipe.d_store[b]) cudaFree(h->pipe.d_store[b]);
        if (h->pipe.d_T[b]) cudaFree(h->pipe.d_T[b]);
        if (h->pipe.d_lp[b]) cudaFree(h->pipe.d_lp[b]);
        if (h->pipe.h_T[b]) cudaFreeHost(h->pipe.h_T[b]);
        if (h->pipe.h_lp[b]) cudaFreeHost(h->pipe.h_lp[b]);
        if (h->pipe.h_flag[b]) cudaFreeHost(h->pipe.h_flag[b]);
        if (h->pipe.ran[b]) cudaEventDestroy(h->pipe.ran[b]);
        if (h->pipe.transposed[b]) cudaEventDestroy(h->pipe.transposed[b]);
        if (h->pipe.copied[b]) cudaEventDestroy(h->pipe.copied[b]);
    }
    if (h->pipe.copy) cudaStreamDestroy(h->pipe.copy);
    if (h->base.prof) cudaFree(h->base.prof);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// ------------------------------------------------------------------------------------
// init (initSess)
// ------------------------------------------------------------------------------------
// Row-sharded pass with one mode per chain (logistic): every rank needs the sums of ITS chains' modes over ALL rows, so the modes
// are all-gathered, each rank contracts its row range against the modes of every rank (the tensor-core pass, M = world x C), and a
// reduce-scatter returns each rank the totals of its own chains.
static bool fullgrad_gathered(const sgm_handle* h) {
    return h->comm && h->world > 1 && h->model == SGM_MODEL_LOGISTIC && !(h->flags & SGM_FLAG_SHARED_MODE);
}
static int fullgrad_gather_modes(sgm_handle* h) {
    const size_t per = (size_t)h->C * h->Pst, all = per * h->world;
    if (!h->hat_all) { CU(cudaMalloc(&h->hat_all, all * sizeof(float))); CU(cudaMalloc(&h->gsum_all, all * sizeof(double))); }
    const int rc = g_nccl.AllGather(h->theta_hat, h->hat_all, per, 7 /* ncclFloat32 */, h->comm, h->stream);
    if (rc != 0) return fail(SGM_ECOMM, "ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return SGM_OK;
}

static int full_gradient(sgm_handle* h) {
    const int M = hat_count(h);
    // replicated dataset: this rank sums its 1/world slice; sharded dataset: its whole local shard
    int64_t row0 = h->n_shards > 1 ? 0 : h->N * h->rank / h->world, row1 = h->n_shards > 1 ? h->N : h->N * (h->rank + 1) / h->world;
    if (h->opt_row1 > 0) {
        if (h->opt_row0 >= h->opt_row1 || h->opt_row1 > h->N) return fail(SGM_EINVAL, "fullgrad_row0 / fullgrad_row1 = [%d, %d) is not a row range of this dataset", h->opt_row0, h->opt_row1);
        row0 = h->opt_row0; row1 = h->opt_row1;
    }
    const bool gathered = fullgrad_gathered(h);
    h->gather_ms = 0.f;
    if (gathered) {
        EventPair evg;
        CU(cudaEventCreate(&evg.e0)); CU(cudaEventCreate(&evg.e1));
        CU(cudaEventRecord(evg.e0, h->stream));
        if (int rc = fullgrad_gather_modes(h)) return rc;
        CU(cudaEventRecord(evg.e1, h->stream));
        CU(cudaEventSynchronize(evg.e1));
        CU(cudaEventElapsedTime(&h->gather_ms, evg.e0, evg.e1));
        CU(cudaMemsetAsync(h->gsum_all, 0, (size_t)M * h->world * h->Pst * sizeof(double), h->stream));
    }
    CU(cudaMemsetAsync(h->gsum, 0, (size_t)M * h->Pst * sizeof(double), h->stream));
    EventPair ev;
    CU(cudaEventCreate(&ev.e0)); CU(cudaEventCreate(&ev.e1));
    cudaEvent_t e0 = ev.e0, e1 = ev.e1;
    CU(cudaEventRecord(e0, h->stream));
    switch (h->model) {
        case SGM_MODEL_GAUSS_MEAN:
        case SGM_MODEL_MIXTURE2: {
            ChainArgs a = h->base;
            int sms = 0; CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
            int64_t gx = cdiv(row1 - row0, 256 * 4); if (gx > sms * 8) gx = sms * 8; if (gx < 1) gx = 1;
            dim3 grid((unsigned)gx, (unsigned)M);
            if (h->model == SGM_MODEL_GAUSS_MEAN) SGM_LAUNCH(small_fullsum_kernel<SMALL_MODEL_GAUSS>, grid, 256, 0, h->stream, a, row0, row1, h->gsum);
            else SGM_LAUNCH(small_fullsum_kernel<SMALL_MODEL_MIX2>, grid, 256, 0, h->stream, a, row0, row1, h->gsum);
            CU(cudaGetLastError());
            break;
        }
        case SGM_MODEL_LOGISTIC:
            if (gathered) {
                double* own = h->gsum;
                h->gsum = h->gsum_all;
                const int rc = launch_logistic_fullgrad(h, row0, row1, M * h->world, h->hat_all);
                h->gsum = own;
                if (rc) return rc;
            } else if (int rc = launch_logistic_fullgrad(h, row0, row1, M, h->theta_hat)) return rc;
            break;
        case SGM_MODEL_SOFTMAX_NET: {
            const int rc = softmax_net_fullgrad(h->nn, h->base, row0, row1, M, h->theta_hat, h->gsum, h->stream);
            if (rc != 0) return fail(SGM_ECUDA, "softmax net full gradient: %s", cudaGetErrorString((cudaError_t)rc));
            break;
        }
    }
    CU(cudaEventRecord(e1, h->stream));
    h->last_launches += 1;
    EventPair evc;
    if (h->comm) { CU(cudaEventCreate(&evc.e0)); CU(cudaEventCreate(&evc.e1)); CU(cudaEventRecord(evc.e0, h->stream)); }
    if (gathered) {
        const int rc = g_nccl.ReduceScatter(h->gsum_all, h->gsum, (size_t)M * h->Pst, 8 /* ncclFloat64 */, 0 /* ncclSum */, h->comm, h->stream);
        if (rc != 0) return fail(SGM_ECOMM, "ncclReduceScatter failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    } else if (h->comm) {
        // ncclDouble = 8, ncclSum = 0
        const int rc = g_nccl.AllReduce(h->gsum, h->gsum, (size_t)M * h->Pst, 8, 0, h->comm, h->stream);
        if (rc != 0) return fail(SGM_ECOMM, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    }
    if (h->comm) CU(cudaEventRecord(evc.e1, h->stream));
    const int total = M * h->Pst;
    switch (h->model) {
        case SGM_MODEL_GAUSS_MEAN:
            SGM_LAUNCH(small_fullgrad_finish<SMALL_MODEL_GAUSS>, (unsigned)cdiv(M, 128), 128, 0, h->stream, h->base, M, h->gsum, h->g_hat, h->theta); break;
        case SGM_MODEL_MIXTURE2:
            SGM_LAUNCH(small_fullgrad_finish<SMALL_MODEL_MIX2>, (unsigned)cdiv(M, 128), 128, 0, h->stream, h->base, M, h->gsum, h->g_hat, h->theta); break;
        case SGM_MODEL_LOGISTIC:
            SGM_LAUNCH(logistic_fullgrad_finish, (unsigned)cdiv(total, 256), 256, 0, h->stream, h->gsum, h->theta_hat, h->g_hat, h->theta, h->Pst, h->d, h->use_prior, total); break;
        case SGM_MODEL_SOFTMAX_NET: {
            const int rc = softmax_net_fullgrad_finish(h->nn, h->base, M, h->gsum, h->theta_hat, h->g_hat, h->theta, h->stream);
            if (rc != 0) return fail(SGM_ECUDA, "softmax net finish: %s", cudaGetErrorString((cudaError_t)rc));
            break;
        }
    }
    CU(cudaGetLastError());
    if (M == 1 && h->C > 1) {  // shared mode: every chain starts at the one mode (cV.R:98-99)
        SGM_LAUNCH(broadcast_rows_kernel, (unsigned)cdiv((int64_t)h->C * h->Pst, 256), 256, 0, h->stream, h->theta, h->theta_hat, h->C, h->Pst);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventElapsedTime(&h->fullgrad_ms, e0, e1));
    h->collective_ms = 0.f;
    if (h->comm) CU(cudaEventElapsedTime(&h->collective_ms, evc.e0, evc.e1));   // the reduction of the sums (all-reduce, or reduce-scatter), on the device
    h->fullgrad_bytes = (row1 - row0) * (int64_t)h->ld * 4 * (h->model == SGM_MODEL_LOGISTIC ? 1 : 1);
    return SGM_OK;
}

// caller-supplied minibatch rows are dataset row numbers (global ones when the dataset is sharded over ranks)
static int check_rows(const sgm_handle* h, const int32_t* idx, size_t count) {
    const int64_t n_rows = (int64_t)h->base.N;
    for (size_t i = 0; i < count; ++i)
        if (idx[i] < 0 || (int64_t)idx[i] >= n_rows) return fail(SGM_EINVAL, "minibatch row %d (entry %zu) is outside [0, %lld)", idx[i], i, (long long)n_rows);
    return SGM_OK;
}

static int logpost_common(sgm_handle* h, const int32_t* idx, const float* theta, double* out, int chains = -1);
static int logpost_launch(sgm_handle* h, const int32_t* idx, const float* theta, int chains);
static inline uint32_t word2_host(uint64_t step, int L, int sub) { return (uint32_t)(step * (uint64_t)(L + 1) + (uint64_t)sub); }   // = word2_of (common.cuh)

// checkOptDivergence (R/controlVariates.R:114-124): estLogPost at the mode estimate(s) on a fresh device-drawn minibatch
static int opt_checkpoint(sgm_handle* h, int M, int64_t iter) {
    std::vector<double> lp((size_t)M);
    h->check_ctr += 1;
    if (int rc = logpost_common(h, nullptr, h->theta_hat, lp.data(), M)) return rc;
    for (double v : lp) if (std::isnan(v)) return fail(SGM_EDIVERGED, "Chain diverged");   // cV.R:118-120
    if (h->progress) h->progress(h->progress_user, 1, iter, lp.data(), M);
    return SGM_OK;
}

// Mode finder sharded by data rows (north star).  Every SGD step each rank sums the log-likelihood gradients of
// ITS part of the minibatch -- n / world rows drawn without replacement from its own row range (its local shard when the dataset
// is sharded), or, in the validation mode, the caller's rows that fall into its range -- with the global N / n factor; rank 0
// adds the prior gradient; one ncclAllReduce of the P-vector; every rank applies theta_hat += optStepsize * g.  The per-rank
// sums come from the family's chain kernel in its gradient-estimate mode on an index list, so no new gradient code exists here.
static int mode_finder_sharded(sgm_handle* h, const ChainArgs& tmpl, bool det, const int32_t* opt_idx) {
    if (!h->comm || h->world < 2) return fail(SGM_ESTATE, "the row-sharded mode finder needs sgm_comm_init with world > 1");
    if (!(h->flags & SGM_FLAG_SHARED_MODE)) return fail(SGM_EINVAL, "the row-sharded mode finder finds ONE mode for the whole job: set SGM_FLAG_SHARED_MODE");
    const int W = h->world, r = h->rank;
    const bool sharded_data = h->n_shards > 1;
    const int64_t n_glob = sharded_data ? h->N * W : h->N;                    // rows of the whole dataset
    const int64_t row0 = sharded_data ? 0 : h->N * r / W, row1 = sharded_data ? h->N : h->N * (r + 1) / W;   // this rank's rows (local numbering)
    const int64_t glob0 = sharded_data ? h->N * r : row0;                     // global number of local row `row0`
    const int64_t rows = row1 - row0;
    const int n_r = h->n / W + (r < h->n % W ? 1 : 0);
    if (!det && (n_r < 1 || n_r > rows)) return fail(SGM_EINVAL, "row-sharded mode finder: minibatch %d over %d ranks leaves rank %d with %d of its %lld rows", h->n, W, r, n_r, (long long)rows);
    // validation mode: the caller's GLOBAL rows of every iteration, filtered to this rank's range (local numbering)
    std::vector<int32_t> mine; std::vector<int64_t> off((size_t)h->n_iters_opt + 1, 0);
    if (det) {
        for (int64_t it = 0; it < h->n_iters_opt; ++it) {
            for (int j = 0; j < h->n; ++j) {
                const int64_t g = opt_idx[it * h->n + j];
                if (g >= glob0 && g < glob0 + rows) mine.push_back((int32_t)(g - glob0 + row0));
            }
            off[it + 1] = (int64_t)mine.size();
        }
    }
    const size_t idx_bytes = std::max<size_t>(det ? mine.size() : (size_t)n_r, 1) * sizeof(int32_t);
    if (int rc = ensure_buf((void**)&h->d_idx, &h->d_idx_cap, idx_bytes)) return rc;
    if (det && !mine.empty()) CU(cudaMemcpyAsync(h->d_idx, mine.data(), mine.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    DevBuf<float> gbuf;
    DevBuf<int32_t> d_zero;
    CU(gbuf.alloc((size_t)h->Pst)); CU(d_zero.alloc(1));
    float* const g_sum = gbuf.p;
    const int32_t* const zero_row = d_zero.p;
    CU(cudaMemsetAsync(d_zero.p, 0, sizeof(int32_t), h->stream));
    CU(cudaMemsetAsync(g_sum, 0, (size_t)h->Pst * sizeof(float), h->stream));   // the kernels write the live elements only: the padding must be zero, not garbage
    ChainArgs a = tmpl;
    a.C = 1; a.theta = h->theta_hat; a.algo = ALGO_GRAD; a.cv = 0; a.L = 1; a.iters = 1; a.det = 1; a.noise = nullptr; a.noise_len = 0; a.store = nullptr;
    a.N = (uint32_t)h->N; a.n_shards = 0; a.shard_ptr = nullptr;              // every row named below is a LOCAL row
    const float corr = (float)((double)n_glob / (double)h->n);
    const uint32_t bits = perm_bits((uint32_t)rows);
    for (int64_t it = 0; it < h->n_iters_opt; ++it) {
        int cnt = n_r;
        const int32_t* d_list = h->d_idx;
        if (det) { cnt = (int)(off[it + 1] - off[it]); d_list = h->d_idx + off[it]; }
        else {
            SGM_LAUNCH(opt_indices_kernel, (unsigned)cdiv(n_r, 128), 128, 0, h->stream, h->d_idx, n_r, (uint32_t)row0, (uint32_t)rows, bits, (uint32_t)r,
                       word2_host((uint64_t)it, 1, 1), h->base.k0, h->base.k1);
            CU(cudaGetLastError());
        }
        if (cnt > 0) {
            a.idx = d_list; a.n = cnt; a.corr = corr; a.use_prior = (r == 0) ? h->use_prior : 0; a.mom = g_sum; a.step0 = (uint64_t)it;
            if (int rc = launch_chain(h, a)) return rc;
        } else if (r == 0 && h->use_prior) {
            // rank 0 holds none of this minibatch's rows (possible in the validation mode only) but still owes the prior gradient:
            // the estimate on one row with a zero N / n factor is exactly the prior term
            a.idx = zero_row; a.n = 1; a.corr = 0.f; a.use_prior = 1; a.mom = g_sum; a.step0 = (uint64_t)it;
            if (int rc = launch_chain(h, a)) return rc;
        } else {
            CU(cudaMemsetAsync(g_sum, 0, (size_t)h->Pst * sizeof(float), h->stream));
        }
        const int rc = g_nccl.AllReduce(g_sum, g_sum, (size_t)h->Pst, 7 /* ncclFloat32 */, 0 /* ncclSum */, h->comm, h->stream);
        if (rc != 0) return fail(SGM_ECOMM, "ncclAllReduce (mode finder) failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        SGM_LAUNCH(opt_apply_kernel, (unsigned)cdiv(h->Pst, 256), 256, 0, h->stream, h->theta_hat, (const float*)g_sum, (const float*)nullptr, a.opt_lr, h->Pst, h->d_flag);
        CU(cudaGetLastError());
        h->last_launches += 2;
        if ((it + 1) % 100 == 0) {
            if (int rc2 = check_diverged(h, "Chain diverged")) return rc2;
            if (!det) if (int rc2 = opt_checkpoint(h, 1, it + 1)) return rc2;
        }
    }
    return check_diverged(h, "Chain diverged");
}

// What follows a known mode (found by initSess, or restored by sgm_load_state): the stride of the mode arrays and, for the logistic
// family with ONE mode for every chain, the per-datum residual at the mode cached next to y in the rows (logistic_staged.cuh)
static int after_mode_setup(sgm_handle* h, int M) {
    h->base.hat_stride = (M == 1 && h->C > 1) ? 0 : h->Pst;
    h->rhat_cached = false;
    // Peers that map this rank's rows (sgm_attach_peer_shards) read the cached residual too, so with a sharded dataset the
    // cache is only valid when every rank holds the SAME mode (SGM_FLAG_SHARED_MODE); one chain per rank without the flag
    // means one mode per rank, and the kernels then recompute x . beta_hat per gathered row.
    const bool same_mode_everywhere = (h->flags & SGM_FLAG_SHARED_MODE) || h->n_shards <= 1;
    if (h->model == SGM_MODEL_LOGISTIC && M == 1 && same_mode_everywhere && (h->d & 3) != 3) {
        SGM_LAUNCH(logistic_cache_rhat_kernel, (unsigned)cdiv(h->N * 32, 256), 256, 0, h->stream, h->X, h->N, h->ld, h->d, h->theta_hat);
        CU(cudaGetLastError());
        h->rhat_cached = true;
        h->last_launches += 1;
    }
    return SGM_OK;
}

static int init_common(sgm_handle* h, const float* momentum_noise, const int32_t* opt_idx, bool det) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    if (int rc = set_device(h)) return rc;
    Timer tm(h);
    // variable initialisation (R/update.R:127-129): params, optimiser params, thermostats and momenta back to their start values
    if (int rc = upload_initial_state(h)) return rc;
    h->rhat_cached = false;
    // momenta (sghmc's is overwritten by the first refresh; sgnht's persists) -- R/dynamics.R:47,84
    std::vector<int> flat_of, tensor_of;
    host_tables(h, flat_of, tensor_of);
    DevBuf<int> flat_buf, tens_buf;
    CU(flat_buf.alloc(h->Pst)); CU(tens_buf.alloc(h->Pst));
    int *d_flat = flat_buf.p, *d_tens = tens_buf.p;
    CU(cudaMemcpyAsync(d_flat, flat_of.data(), h->Pst * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_tens, tensor_of.data(), h->Pst * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (h->algo != SGM_ALGO_SGLD) {
        const float* dn = nullptr;
        if (det && momentum_noise) {
            const size_t bytes = (size_t)h->C * h->P * sizeof(float);
            if (int rc = ensure_buf((void**)&h->d_noise, &h->d_noise_cap, bytes)) return rc;
            CU(cudaMemcpyAsync(h->d_noise, momentum_noise, bytes, cudaMemcpyHostToDevice, h->stream));
            dn = h->d_noise;
        }
        if (det && !momentum_noise) CU(cudaMemsetAsync(h->mom, 0, (size_t)h->C * h->Pst * sizeof(float), h->stream));
        else {
            SGM_LAUNCH(momentum_init_kernel, (unsigned)cdiv((int64_t)h->C * h->Pst, 256), 256, 0, h->stream, h->mom, h->C, h->Pst, h->P, d_flat, d_tens, h->base.tc, dn, h->base.k0, h->base.k1, h->base.chain0);
            CU(cudaGetLastError());
        }
    }
    CU(cudaStreamSynchronize(h->stream));

    if (h->cv) {
        // getMode (R/controlVariates.R:37-52): n_iters_opt plain SGD ascent steps on theta_hat
        const int M = hat_count(h);
        ChainArgs a = h->base;
        a.C = M; a.theta = h->theta_hat; a.mom = h->scratch; a.algo = ALGO_OPT; a.cv = 0; a.L = 1;
        if (M == 1 && (h->flags & SGM_FLAG_SHARED_MODE)) a.chain0 = 0;  // one mode for the whole job, identical on every rank
        a.index_tag = TAG_OPT_INDEX; a.step0 = 0; a.iters = (int)h->n_iters_opt; a.store = nullptr;
        a.noise_len = 0;
        if (det) {
            if (h->n_iters_opt > 0 && !opt_idx) return fail(SGM_EINVAL, "opt_idx required for a deterministic *cv init");
            if (opt_idx) if (int rc = check_rows(h, opt_idx, (size_t)M * h->n_iters_opt * h->n)) return rc;
            const size_t bytes = (size_t)M * h->n_iters_opt * h->n * sizeof(int32_t);
            if (bytes) {
                if (int rc = ensure_buf((void**)&h->d_idx, &h->d_idx_cap, bytes)) return rc;
                CU(cudaMemcpyAsync(h->d_idx, opt_idx, bytes, cudaMemcpyHostToDevice, h->stream));
            }
            a.det = 1; a.idx = h->d_idx;
        }
        const bool sharded_finder = h->opt_sharded_finder && h->n_iters_opt > 0;
        if (sharded_finder) {
            if (int rc = mode_finder_sharded(h, a, det, opt_idx)) return rc;
        } else if (h->n_iters_opt > 0) {
            if (det) {   // validation mode: one launch over the caller's index table
                if (int rc = launch_chain(h, a)) return rc;
                if (int rc = check_diverged(h, "Chain diverged")) return rc;   // cV.R:119
            } else {
                // getMode's loop (cV.R:42-47) in launches of <= 100 iterations: after each one checkOptDivergence -- the log posterior
                // estimate at theta_hat on a FRESH minibatch; NaN stops with "Chain diverged"; the value goes to the progress callback
                for (int64_t done = 0; done < h->n_iters_opt;) {
                    const int64_t chunk = std::min<int64_t>(100 - done % 100, h->n_iters_opt - done);
                    a.step0 = (uint64_t)done; a.iters = (int)chunk;
                    if (int rc = launch_chain(h, a)) return rc;
                    done += chunk;
                    if (int rc = check_diverged(h, "Chain diverged")) return rc;
                    if (done % 100 == 0) if (int rc = opt_checkpoint(h, M, done)) return rc;
                }
            }
        }
        if (int rc = full_gradient(h)) return rc;
        if (int rc = after_mode_setup(h, M)) return rc;
        if (h->n_shards > 1) {
            // peers read the residuals cached in OUR rows: nobody may sample before every rank has written its shard.  A one-double
            // all-reduce enqueued after the cache kernel completes only when all ranks have reached it (stream-ordered rendezvous).
            if (!h->comm) return fail(SGM_ESTATE, "a sharded dataset with control variates needs sgm_comm_init (reduction of the shard sums)");
            const int rc = g_nccl.AllReduce(h->gsum, h->gsum, 1, 8, 0, h->comm, h->stream);
            if (rc != 0) return fail(SGM_ECOMM, "ncclAllReduce (rendezvous) failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
            CU(cudaStreamSynchronize(h->stream));
        }
    }
    h->inited = true;
    h->step = 0;
    return tm.stop();
}

extern "C" int sgm_init(sgm_handle* h) { return init_common(h, nullptr, nullptr, false); }
extern "C" int sgm_init_deterministic(sgm_handle* h, const float* momentum_noise, const int32_t* opt_idx) {
    return init_common(h, momentum_noise, opt_idx, true);
}

// ------------------------------------------------------------------------------------
// stepping
// ------------------------------------------------------------------------------------
extern "C" int sgm_step_deterministic(sgm_handle* h, const int32_t* idx, const float* noise) {
    if (!h || !idx || !noise) return fail(SGM_EINVAL, "null argument");
    if (!h->inited) return fail(SGM_ESTATE, "sgm_init has not been called");
    if (int rc = check_rows(h, idx, (size_t)h->C * h->L * h->n)) return rc;
    if (int rc = set_device(h)) return rc;
    const size_t ib = (size_t)h->C * h->L * h->n * sizeof(int32_t);
    const size_t nb = (size_t)h->C * h->base.noise_len * sizeof(float);
    if (int rc = ensure_buf((void**)&h->d_idx, &h->d_idx_cap, ib)) return rc;
    if (int rc = ensure_buf((void**)&h->d_noise, &h->d_noise_cap, nb)) return rc;
    CU(cudaMemcpyAsync(h->d_idx, idx, ib, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_noise, noise, nb, cudaMemcpyHostToDevice, h->stream));
    Timer tm(h);
    ChainArgs a = h->base;
    a.det = 1; a.idx = h->d_idx; a.noise = h->d_noise; a.iters = 1; a.step0 = h->step; a.store = nullptr;
    if (int rc = launch_chain(h, a)) return rc;
    h->step += 1;
    if (int rc = tm.stop()) return rc;
    return SGM_OK;
}

static int ensure_store(sgm_handle* h, int64_t extra) {
    const int64_t need = h->store_len + extra;
    if (need <= h->store_cap) return SGM_OK;
    int64_t cap = h->store_cap * 2 > need ? h->store_cap * 2 : need;
    float* nb = nullptr;
    const size_t row = (size_t)h->C * h->P * sizeof(float);
    if (cudaMalloc(&nb, cap * row) != cudaSuccess) {
        cudaGetLastError();
        cap = need;
        if (cudaMalloc(&nb, cap * row) != cudaSuccess) { cudaGetLastError(); return fail(SGM_ENOMEM, "cannot allocate %zu bytes of chain storage", cap * row); }
    }
    if (h->store) {
        CU(cudaMemcpyAsync(nb, h->store, h->store_len * row, cudaMemcpyDeviceToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        cudaFree(h->store);
    }
    h->store = nb; h->store_cap = cap;
    return SGM_OK;
}

extern "C" int sgm_reserve_chain(sgm_handle* h, int64_t n_iters) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    if (int rc = set_device(h)) return rc;
    return ensure_store(h, n_iters);
}

extern "C" int sgm_run(sgm_handle* h, int64_t n_iters, int store) {
    if (!h || n_iters < 0) return fail(SGM_EINVAL, "bad argument");
    if (n_iters > (int64_t)INT32_MAX / (h->L + 1)) return fail(SGM_EINVAL, "n_iters = %lld is more than one launch runs; call sgm_run in chunks", (long long)n_iters);
    if (!h->inited) return fail(SGM_ESTATE, "sgm_init has not been called");
    if (n_iters == 0) return SGM_OK;
    if (int rc = set_device(h)) return rc;
    if (store) if (int rc = ensure_store(h, n_iters)) return rc;
    Timer tm(h);
    ChainArgs a = h->base;
    a.det = 0; a.iters = (int)n_iters; a.step0 = h->step;
    a.store = store ? h->store : nullptr; a.store_pos = h->store_len;
    if (int rc = launch_chain(h, a)) return rc;
    h->step += (uint64_t)n_iters;
    if (store) h->store_len += n_iters;
    if (int rc = tm.stop()) return rc;
    return check_diverged(h, "Chain diverged, try decreasing stepsize");  // R/update.R:202
}

extern "C" int sgm_step(sgm_handle* h) { return sgm_run(h, 1, 0); }

// ---- runSGMCMC's loop (R/update.R:145-163) in ONE call, the chain streamed to the host while later iterations run ---------------
static int pipe_setup(sgm_handle* h, int64_t blk) {
    sgm_handle::Pipe& p = h->pipe;
    if (p.blk == blk && p.copy) return SGM_OK;
    for (int b = 0; b < 2; ++b) {
        if (p.d_store[b]) { cudaFree(p.d_store[b]); p.d_store[b] = nullptr; }
        if (p.d_T[b]) { cudaFree(p.d_T[b]); p.d_T[b] = nullptr; }
        if (p.h_T[b]) { cudaFreeHost(p.h_T[b]); p.h_T[b] = nullptr; }
    }
    const size_t bytes = (size_t)blk * h->C * h->P * sizeof(float);
    if (!p.copy) CU(cudaStreamCreateWithFlags(&p.copy, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        CU(cudaMalloc(&p.d_store[b], bytes)); CU(cudaMalloc(&p.d_T[b], bytes));
        CU(cudaHostAlloc(&p.h_T[b], bytes, cudaHostAllocDefault));
        if (!p.d_lp[b]) {
            CU(cudaMalloc(&p.d_lp[b], (size_t)h->C * sizeof(double)));
            CU(cudaHostAlloc(&p.h_lp[b], (size_t)h->C * sizeof(double), cudaHostAllocDefault));
            CU(cudaHostAlloc(&p.h_flag[b], sizeof(int), cudaHostAllocDefault));
            CU(cudaEventCreate(&p.ran[b])); CU(cudaEventCreate(&p.transposed[b])); CU(cudaEventCreate(&p.copied[b]));
        }
    }
    p.blk = blk;
    return SGM_OK;
}

// float [C][P][cnt] (one block) -> double out[(c * P + f) * n_iters + it0 + i], rows of `cnt` contiguous values; split over host threads.
// The destination is written once and not read again here: streaming (non-temporal) stores skip the read-for-ownership of every
// destination line, 1.85x the rate of plain stores at 8 threads on the development host (12.4 vs 22.9 ms for a c3 step of 286 MB).
static void widen_rows(const float* src, double* out, int64_t r0, int64_t r1, int64_t cnt, int64_t n_iters, int64_t it0) {
    for (int64_t r = r0; r < r1; ++r) {
        const float* s_ = src + r * cnt;
        double* d_ = out + r * n_iters + it0;
        int64_t i = 0;
#if defined(__SSE2__)
        if (((uintptr_t)d_ & 15u) && cnt > 0) { d_[0] = (double)s_[0]; i = 1; }   // doubles are 8-byte aligned: one step reaches 16
        for (; i + 4 <= cnt; i += 4) {
            const __m128 v = _mm_loadu_ps(s_ + i);
            _mm_stream_pd(d_ + i, _mm_cvtps_pd(v));
            _mm_stream_pd(d_ + i + 2, _mm_cvtps_pd(_mm_movehl_ps(v, v)));
        }
#endif
        for (; i < cnt; ++i) d_[i] = (double)s_[i];
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
}
// `ranks_on_host`: the processes of the job that share this host's cores (sgm_comm_init's world on a single node): each takes its share
static void widen_block(const float* src, double* out, int C, int P, int64_t cnt, int64_t n_iters, int64_t it0, int ranks_on_host) {
    const int64_t rows = (int64_t)C * P;
    unsigned hw = std::thread::hardware_concurrency();
    if (ranks_on_host > 1) hw = hw / (unsigned)ranks_on_host > 0 ? hw / (unsigned)ranks_on_host : 1;
    int T = (int)(hw ? (hw > 16 ? 16 : hw) : 1);
    if (rows * cnt < (1 << 16)) T = 1;
    if (T == 1) { widen_rows(src, out, 0, rows, cnt, n_iters, it0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(widen_rows, src, out, rows * t / T, rows * (t + 1) / T, cnt, n_iters, it0);
    for (auto& x : th) x.join();
}

extern "C" int sgm_sample(sgm_handle* h, int64_t n_iters, int32_t check_every, double* out) {
    if (!h || !out || n_iters < 0 || check_every < 0) return fail(SGM_EINVAL, "bad argument");
    if (!h->inited) return fail(SGM_ESTATE, "sgm_init has not been called");
    if (n_iters == 0) return SGM_OK;
    if (int rc = set_device(h)) return rc;
    // block = the divergence-check cadence (100 in runSGMCMC), shortened so that one block of the chain stays <= 128 MiB
    int64_t blk = check_every > 0 ? check_every : 100;
    const size_t per_iter = (size_t)h->C * h->P * sizeof(float);
    while (blk > 1 && (size_t)blk * per_iter > (128ull << 20)) blk = (blk + 1) / 2;
    if (check_every > 0) while (check_every % blk) --blk;     // checks fall on block boundaries
    if (int rc = pipe_setup(h, blk)) return rc;
    sgm_handle::Pipe& p = h->pipe;
    // A page-locked caller array (sgm_host_alloc, cudaHostAlloc, cudaHostRegister) is written by the GPU directly, as doubles, block by
    // block on the copy stream: no staging buffer and no host-side widening.  A pageable array (what R allocates) takes the staged path:
    // float32 over PCIe into page-locked staging, widened into the caller's array by host threads.
    double* out_dev = nullptr;
    if (h->opt_sample_out != 1) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, out) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) out_dev = static_cast<double*>(pa.devicePointer);
        else cudaGetLastError();
    }
    Timer tm(h);
    const int64_t n_blocks = cdiv(n_iters, blk);
    int status = SGM_OK;
    auto retire = [&](int64_t k) -> int {   // block k is on its way to the host: wait for it, check it, widen it into `out`
        const int b = (int)(k & 1);
        const int64_t it0 = k * blk, cnt = std::min(blk, n_iters - it0);
        CU(cudaEventSynchronize(p.copied[b]));
        if (*p.h_flag[b]) return fail(SGM_EDIVERGED, "Chain diverged, try decreasing stepsize");   // R/update.R:202
        if (check_every > 0 && (it0 + cnt) % check_every == 0) {
            for (int c = 0; c < h->C; ++c) if (std::isnan(p.h_lp[b][c])) return fail(SGM_EDIVERGED, "Chain diverged, try decreasing stepsize");
            if (h->progress) h->progress(h->progress_user, 2, it0 + cnt, p.h_lp[b], h->C);
        }
        if (!out_dev) widen_block(p.h_T[b], out, h->C, h->P, cnt, n_iters, it0, h->world);
        return SGM_OK;
    };
    for (int64_t k = 0; k < n_blocks && status == SGM_OK; ++k) {
        const int b = (int)(k & 1);
        const int64_t it0 = k * blk, cnt = std::min(blk, n_iters - it0);
        // compute stream: block k's kernel may overwrite d_store[b] once block k - 2 has been transposed out of it
        if (k >= 2) CU(cudaStreamWaitEvent(h->stream, p.transposed[b], 0));
        ChainArgs a = h->base;
        a.det = 0; a.iters = (int)cnt; a.step0 = h->step; a.store = p.d_store[b]; a.store_pos = 0;
        if (int rc = launch_chain(h, a)) { status = rc; break; }
        h->step += (uint64_t)cnt;
        const bool check = check_every > 0 && (it0 + cnt) % check_every == 0;
        if (check) {   // checkDivergence (R/update.R:197-207): estLogPost on a fresh minibatch, on the device, result copied with the block
            h->check_ctr += 1;
            if (int rc = logpost_launch(h, nullptr, h->theta, h->C)) { status = rc; break; }
            CU(cudaMemcpyAsync(p.d_lp[b], h->dout, (size_t)h->C * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        }
        CU(cudaEventRecord(p.ran[b], h->stream));
        // copy stream: transpose block k, ship it (float), the divergence flag and the check values
        if (k >= 2) { /* h_T[b] was consumed by retire(k - 2) below before this point */ }
        CU(cudaStreamWaitEvent(p.copy, p.ran[b], 0));
        {
            dim3 grid((unsigned)cdiv(cnt, 32), (unsigned)cdiv(h->P, 32), (unsigned)1), block(32, 8);
            for (int c0 = 0; c0 < h->C; c0 += 65535) {
                const int cn = std::min(65535, h->C - c0);
                grid.z = (unsigned)cn;
                if (out_dev) SGM_LAUNCH(chain_block_to_host_kernel, grid, block, 0, p.copy, (const float*)p.d_store[b], cnt, h->C, h->P, c0, n_iters, it0, out_dev);
                else SGM_LAUNCH(chain_transpose_kernel<float>, grid, block, 0, p.copy, (const float*)p.d_store[b], cnt, h->C, h->P, c0, cn, p.d_T[b] + (size_t)c0 * h->P * cnt);
            }
            CU(cudaGetLastError());
            h->last_launches += (h->C + 65534) / 65535;
        }
        CU(cudaEventRecord(p.transposed[b], p.copy));
        if (!out_dev) CU(cudaMemcpyAsync(p.h_T[b], p.d_T[b], (size_t)cnt * per_iter, cudaMemcpyDeviceToHost, p.copy));
        CU(cudaMemcpyAsync(p.h_flag[b], h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, p.copy));
        if (check) CU(cudaMemcpyAsync(p.h_lp[b], p.d_lp[b], (size_t)h->C * sizeof(double), cudaMemcpyDeviceToHost, p.copy));
        CU(cudaEventRecord(p.copied[b], p.copy));
        // host: while block k runs, finish block k - 1 (its buffers are the OTHER pair)
        if (k >= 1) status = retire(k - 1);
    }
    if (status == SGM_OK) status = retire(n_blocks - 1);
    cudaStreamSynchronize(p.copy);
    cudaStreamSynchronize(h->stream);
    if (status == SGM_EDIVERGED) cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream);
    if (status != SGM_OK) return status;
    return tm.stop();
}

// debug: per-phase clock64() totals of thread 0 of every CTA (see LS_TICK in logistic_staged.cuh); n slots of uint64
extern "C" int sgm_debug_phase_cycles(sgm_handle* h, int enable, unsigned long long* out, int n) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    if (int rc = set_device(h)) return rc;
    if (enable && !h->base.prof) { CU(cudaMalloc(&h->base.prof, 64 * sizeof(unsigned long long))); }
    if (h->base.prof && out) { CU(cudaMemcpy(out, h->base.prof, (n > 64 ? 64 : n) * sizeof(unsigned long long), cudaMemcpyDeviceToHost)); }
    if (h->base.prof) CU(cudaMemset(h->base.prof, 0, 64 * sizeof(unsigned long long)));
    if (!enable && h->base.prof) { cudaFree(h->base.prof); h->base.prof = nullptr; }
    return SGM_OK;
}


// ------------------------------------------------------------------------------------
// state dump / restore (the reference leaves persistence to the user of the step-by-step API, NEWS.md:19)
// ------------------------------------------------------------------------------------
namespace {
struct StateHeader {
    uint64_t magic;          // "SGMSTAT1"
    int32_t model, algo, cv, C, P, Pst, T, M;
    int64_t N; int32_t n, d;
    uint64_t seed, step, check_ctr;
    uint64_t payload_floats;
};
constexpr uint64_t STATE_MAGIC = 0x31544154534d4753ull;
}  // namespace

static size_t state_floats(const sgm_handle* h, int M) {
    return (size_t)2 * h->C * h->Pst + (size_t)h->C * MAX_TENSORS + (h->cv ? (size_t)2 * M * h->Pst : 0);
}

extern "C" int sgm_save_state(sgm_handle* h, void* buf, size_t cap, size_t* need) {
    if (!h || !need) return fail(SGM_EINVAL, "null argument");
    if (!h->inited) return fail(SGM_ESTATE, "sgm_init has not been called");
    const int M = h->cv ? hat_count(h) : 0;
    const size_t nf = state_floats(h, M), bytes = sizeof(StateHeader) + nf * sizeof(float);
    *need = bytes;
    if (!buf) return SGM_OK;                       // size query
    if (cap < bytes) return fail(SGM_EINVAL, "state buffer of %zu bytes, %zu needed", cap, bytes);
    if (int rc = set_device(h)) return rc;
    StateHeader hd{};
    hd.magic = STATE_MAGIC; hd.model = h->model; hd.algo = h->algo; hd.cv = h->cv; hd.C = h->C; hd.P = h->P; hd.Pst = h->Pst; hd.T = h->T; hd.M = M;
    hd.N = h->N; hd.n = h->n; hd.d = h->d; hd.seed = h->seed; hd.step = h->step; hd.check_ctr = h->check_ctr; hd.payload_floats = nf;
    std::memcpy(buf, &hd, sizeof hd);
    float* out = reinterpret_cast<float*>(static_cast<char*>(buf) + sizeof hd);
    const size_t sb = (size_t)h->C * h->Pst, tb = (size_t)h->C * MAX_TENSORS, mb = (size_t)M * h->Pst;
    CU(cudaMemcpyAsync(out, h->theta, sb * sizeof(float), cudaMemcpyDeviceToHost, h->stream)); out += sb;
    CU(cudaMemcpyAsync(out, h->mom, sb * sizeof(float), cudaMemcpyDeviceToHost, h->stream)); out += sb;
    CU(cudaMemcpyAsync(out, h->thermo, tb * sizeof(float), cudaMemcpyDeviceToHost, h->stream)); out += tb;
    if (h->cv) {
        CU(cudaMemcpyAsync(out, h->theta_hat, mb * sizeof(float), cudaMemcpyDeviceToHost, h->stream)); out += mb;
        CU(cudaMemcpyAsync(out, h->g_hat, mb * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return SGM_OK;
}

extern "C" int sgm_load_state(sgm_handle* h, const void* buf, size_t bytes) {
    if (!h || !buf) return fail(SGM_EINVAL, "null argument");
    if (bytes < sizeof(StateHeader)) return fail(SGM_EINVAL, "state blob too short");
    StateHeader hd;
    std::memcpy(&hd, buf, sizeof hd);
    const int M = h->cv ? hat_count(h) : 0;
    if (hd.magic != STATE_MAGIC) return fail(SGM_EINVAL, "not a sampler state written by sgm_save_state");
    if (hd.model != h->model || hd.algo != h->algo || hd.cv != h->cv || hd.C != h->C || hd.P != h->P || hd.Pst != h->Pst || hd.T != h->T || hd.M != M ||
        hd.N != h->N || hd.n != h->n || hd.d != h->d)
        return fail(SGM_EINVAL, "the state was saved by a sampler of another shape (model / algorithm / chains / parameters / dataset size / minibatch differ)");
    if (hd.payload_floats != state_floats(h, M) || bytes < sizeof hd + hd.payload_floats * sizeof(float)) return fail(SGM_EINVAL, "state blob truncated");
    if (int rc = set_device(h)) return rc;
    const float* in = reinterpret_cast<const float*>(static_cast<const char*>(buf) + sizeof hd);
    const size_t sb = (size_t)h->C * h->Pst, tb = (size_t)h->C * MAX_TENSORS, mb = (size_t)M * h->Pst;
    CU(cudaMemcpyAsync(h->theta, in, sb * sizeof(float), cudaMemcpyHostToDevice, h->stream)); in += sb;
    CU(cudaMemcpyAsync(h->mom, in, sb * sizeof(float), cudaMemcpyHostToDevice, h->stream)); in += sb;
    CU(cudaMemcpyAsync(h->thermo, in, tb * sizeof(float), cudaMemcpyHostToDevice, h->stream)); in += tb;
    if (h->cv) {
        CU(cudaMemcpyAsync(h->theta_hat, in, mb * sizeof(float), cudaMemcpyHostToDevice, h->stream)); in += mb;
        CU(cudaMemcpyAsync(h->g_hat, in, mb * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    // the random streams are keyed by (seed, chain, iteration): restoring the seed and the counters continues them exactly
    h->seed = hd.seed; h->base.k0 = (uint32_t)(hd.seed & 0xffffffffu); h->base.k1 = (uint32_t)(hd.seed >> 32);
    h->step = hd.step; h->check_ctr = hd.check_ctr;
    if (h->cv) { if (int rc = after_mode_setup(h, M)) return rc; CU(cudaStreamSynchronize(h->stream)); }
    h->inited = true;
    return SGM_OK;
}

// ------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------
// getParams is called once per iteration by users of the step-by-step API (vignettes/nn.Rmd:99-102): the index table and the
// conversion buffer live with the handle (a cudaMalloc / cudaFree pair per call costs more than the sgmcmcStep it follows)
static int state_out(sgm_handle* h, const float* src, int count, double* out) {
    if (int rc = set_device(h)) return rc;
    if (!h->d_flat_of) {
        std::vector<int> flat_of, tensor_of;
        host_tables(h, flat_of, tensor_of);
        CU(cudaMalloc(&h->d_flat_of, (size_t)h->Pst * sizeof(int)));
        CU(cudaMemcpyAsync(h->d_flat_of, flat_of.data(), h->Pst * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));   // the host vector goes out of scope
    }
    if (int rc = ensure_buf((void**)&h->d_state_out, &h->d_state_out_cap, (size_t)count * h->P * sizeof(double))) return rc;
    int* d_flat = h->d_flat_of; double* d_out = h->d_state_out;
    SGM_LAUNCH(state_to_double_kernel, (unsigned)cdiv((int64_t)count * h->Pst, 256), 256, 0, h->stream, src, count, h->Pst, h->P, d_flat, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, (size_t)count * h->P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SGM_OK;
}

extern "C" int sgm_get_params(sgm_handle* h, double* out) {
    if (!h || !out) return fail(SGM_EINVAL, "null argument");
    return state_out(h, h->theta, h->C, out);
}

extern "C" int sgm_get_state(sgm_handle* h, double* momentum, double* thermostat) {
    if (!h) return fail(SGM_EINVAL, "null argument");
    if (momentum) if (int rc = state_out(h, h->mom, h->C, momentum)) return rc;
    if (thermostat) {
        if (int rc = set_device(h)) return rc;
        std::vector<float> th((size_t)h->C * MAX_TENSORS);
        CU(cudaMemcpyAsync(th.data(), h->thermo, th.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (int c = 0; c < h->C; ++c) for (int t = 0; t < h->T; ++t) thermostat[(size_t)c * h->T + t] = th[(size_t)c * MAX_TENSORS + t];
    }
    return SGM_OK;
}

extern "C" int sgm_get_mode(sgm_handle* h, double* theta_hat, double* grad_hat) {
    if (!h) return fail(SGM_EINVAL, "null argument");
    if (!h->cv) return fail(SGM_ESTATE, "not a control-variate sampler");
    const int M = hat_count(h);
    std::vector<double> tmp((size_t)M * h->P);
    for (int which = 0; which < 2; ++which) {
        double* dst = which ? grad_hat : theta_hat;
        if (!dst) continue;
        if (int rc = state_out(h, which ? h->g_hat : h->theta_hat, M, tmp.data())) return rc;
        for (int c = 0; c < h->C; ++c) std::memcpy(dst + (size_t)c * h->P, tmp.data() + (size_t)(M == 1 ? 0 : c) * h->P, h->P * sizeof(double));
    }
    return SGM_OK;
}

extern "C" int64_t sgm_chain_length(const sgm_handle* h) { return h ? h->store_len : 0; }
extern "C" int sgm_reset_chain(sgm_handle* h) { if (!h) return fail(SGM_EINVAL, "null handle"); h->store_len = 0; return SGM_OK; }

extern "C" int sgm_copy_chain(sgm_handle* h, double* out) {
    if (!h || !out) return fail(SGM_EINVAL, "null argument");
    if (h->store_len == 0) return SGM_OK;
    if (int rc = set_device(h)) return rc;
    const int64_t iters = h->store_len;
    // chains per slab so that the double staging buffer stays <= 256 MiB
    const size_t per_chain = (size_t)iters * h->P * sizeof(double);
    int slab = (int)((256ull << 20) / (per_chain ? per_chain : 1));
    if (slab < 1) slab = 1;
    if (slab > h->C) slab = h->C;
    if (slab > 65535) slab = 65535;
    DevBuf<double> stage_buf;
    if (cudaMalloc(&stage_buf.p, (size_t)slab * per_chain) != cudaSuccess) { cudaGetLastError(); return fail(SGM_ENOMEM, "cannot allocate chain staging buffer"); }
    double* stage = stage_buf.p;
    for (int c0 = 0; c0 < h->C; c0 += slab) {
        const int cn = (h->C - c0 < slab) ? h->C - c0 : slab;
        dim3 grid((unsigned)cdiv(iters, 32), (unsigned)cdiv(h->P, 32), (unsigned)cn), block(32, 8);
        SGM_LAUNCH(chain_transpose_kernel<double>, grid, block, 0, h->stream, h->store, iters, h->C, h->P, c0, cn, stage);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out + (size_t)c0 * iters * h->P, stage, (size_t)cn * per_chain, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return SGM_OK;
}

// launches the family's estLogPost kernel into h->dout (no copy, no synchronisation)
static int logpost_launch(sgm_handle* h, const int32_t* idx, const float* theta, int chains) {
    if (int rc = set_device(h)) return rc;
    ChainArgs a = h->base;
    a.theta = const_cast<float*>(theta);
    if (chains > 0) a.C = chains;   // the mode estimates: hat_count(h) states
    const int Cn = a.C;
    if (idx) {
        const size_t ib = (size_t)h->C * h->n * sizeof(int32_t);
        if (int rc = ensure_buf((void**)&h->d_idx, &h->d_idx_cap, ib)) return rc;
        CU(cudaMemcpyAsync(h->d_idx, idx, ib, cudaMemcpyHostToDevice, h->stream));
        a.det = 1; a.idx = h->d_idx;
    }
    const uint32_t word2 = (uint32_t)h->check_ctr;
    switch (h->model) {
        case SGM_MODEL_GAUSS_MEAN:
            SGM_LAUNCH(small_logpost_kernel<SMALL_MODEL_GAUSS>, (unsigned)cdiv((int64_t)Cn * 32, 128), 128, 0, h->stream, a, word2, h->dout); break;
        case SGM_MODEL_MIXTURE2:
            SGM_LAUNCH(small_logpost_kernel<SMALL_MODEL_MIX2>, (unsigned)cdiv((int64_t)Cn * 32, 128), 128, 0, h->stream, a, word2, h->dout); break;
        case SGM_MODEL_LOGISTIC:
            SGM_LAUNCH(logistic_logpost_kernel, Cn, 256, 0, h->stream, a, word2, theta, h->dout); break;
        case SGM_MODEL_SOFTMAX_NET: {
            const int rc = softmax_net_logpost(h->nn, a, word2, theta, h->dout, h->stream);
            if (rc != 0) return fail(SGM_ECUDA, "softmax net logpost: %s", cudaGetErrorString((cudaError_t)rc));
            break;
        }
    }
    CU(cudaGetLastError());
    (void)Cn;
    return SGM_OK;
}

static int logpost_common(sgm_handle* h, const int32_t* idx, const float* theta, double* out, int chains) {
    if (int rc = logpost_launch(h, idx, theta, chains)) return rc;
    const int Cn = chains > 0 ? chains : h->C;
    CU(cudaMemcpyAsync(out, h->dout, (size_t)Cn * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SGM_OK;
}

// what sgm_init reports every 100 iterations of the mode finder (include/sgmcmc_b200.h)
extern "C" int sgm_set_progress(sgm_handle* h, sgm_progress_fn fn, void* user) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    h->progress = fn; h->progress_user = user;
    return SGM_OK;
}

extern "C" int sgm_est_logpost(sgm_handle* h, double* out) {
    if (!h || !out) return fail(SGM_EINVAL, "null argument");
    h->check_ctr += 1;
    return logpost_common(h, nullptr, h->theta, out);
}

extern "C" int sgm_est_logpost_deterministic(sgm_handle* h, const int32_t* idx, double* out) {
    if (!h || !out || !idx) return fail(SGM_EINVAL, "null argument");
    if (int rc = check_rows(h, idx, (size_t)h->C * h->n)) return rc;
    return logpost_common(h, idx, h->theta, out);
}

extern "C" int sgm_grad_estimate(sgm_handle* h, const int32_t* idx, double* out) {
    if (!h || !idx || !out) return fail(SGM_EINVAL, "null argument");
    if (!h->inited) return fail(SGM_ESTATE, "sgm_init has not been called");
    if (int rc = check_rows(h, idx, (size_t)h->C * h->n)) return rc;
    if (int rc = set_device(h)) return rc;
    const size_t ib = (size_t)h->C * h->n * sizeof(int32_t);
    if (int rc = ensure_buf((void**)&h->d_idx, &h->d_idx_cap, ib)) return rc;
    CU(cudaMemcpyAsync(h->d_idx, idx, ib, cudaMemcpyHostToDevice, h->stream));
    // plain-SGD kernel with unit step on a scratch copy: g = theta' - theta would lose bits, so the
    // kernels' ALGO_GRAD path writes the estimate into the momentum slot of a scratch state instead
    CU(cudaMemcpyAsync(h->scratch, h->theta, (size_t)h->C * h->Pst * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    DevBuf<float> g_buf;
    CU(g_buf.alloc((size_t)h->C * h->Pst));
    float* gbuf = g_buf.p;
    CU(cudaMemsetAsync(gbuf, 0, (size_t)h->C * h->Pst * sizeof(float), h->stream));
    ChainArgs a = h->base;
    a.theta = h->scratch; a.mom = gbuf; a.algo = ALGO_GRAD; a.L = 1; a.det = 1; a.idx = h->d_idx; a.noise = nullptr;
    a.iters = 1; a.store = nullptr; a.noise_len = 0;
    int rc = launch_chain(h, a);
    if (rc == SGM_OK) rc = state_out(h, gbuf, h->C, out);
    return rc;
}

// log-likelihood of the current state of every chain on caller data (no prior, no N/n factor); replaces the
// user-built TensorFlow tensors of vignettes/nn.Rmd:113-123 and the R loop of logisticRegression.Rmd:89-115
extern "C" int sgm_heldout_loglik(sgm_handle* h, const sgm_array* data, int32_t n_data, double* out) {
    if (!h || !data || !out) return fail(SGM_EINVAL, "null argument");
    const int entries = (h->model == SGM_MODEL_GAUSS_MEAN || h->model == SGM_MODEL_MIXTURE2) ? 1 : 2;
    if (n_data < entries) return fail(SGM_EINVAL, "held-out data needs %d entries", entries);
    const int64_t Nh = data[0].dims[0];
    if (Nh < 1 || Nh >= (1ll << 31)) return fail(SGM_EINVAL, "held-out rows must be in [1, 2^31)");
    if (int rc = set_device(h)) return rc;
    DevBuf<float> xh_buf;
    const size_t xbytes = (size_t)Nh * h->ld * sizeof(float);
    if (cudaMalloc(&xh_buf.p, xbytes) != cudaSuccess) { cudaGetLastError(); return fail(SGM_ENOMEM, "cannot allocate %zu bytes for the held-out set", xbytes); }
    float* Xh = xh_buf.p;
    CU(cudaMemsetAsync(Xh, 0, xbytes, h->stream));
    int col0 = 0;
    for (int k = 0; k < entries; ++k) {
        const sgm_array& A = data[k];
        const int cols = A.ndim >= 2 ? (int)A.dims[1] : 1;
        const int want = (k == 0) ? h->d : (h->ld >= h->d + cols ? cols : -1);
        if (!A.data || A.dims[0] != Nh || (k == 0 && cols != want) || want < 0) return fail(SGM_EINVAL, "held-out entry %d has the wrong shape", k);
        CU(upload_entry(A.data, Nh, cols, (int)(A.row_major != 0), Xh, h->ld, col0, h->stream));
        col0 += cols;
    }
    ChainArgs a = h->base;
    a.X = Xh; a.N = (uint32_t)Nh; a.n = (int)Nh; a.corr = 1.0f; a.use_prior = 0; a.det = 2; a.idx = nullptr;
    int rc = SGM_OK;
    switch (h->model) {
        case SGM_MODEL_GAUSS_MEAN:
            SGM_LAUNCH(small_logpost_kernel<SMALL_MODEL_GAUSS>, (unsigned)cdiv((int64_t)h->C * 32, 128), 128, 0, h->stream, a, 0u, h->dout); break;
        case SGM_MODEL_MIXTURE2:
            SGM_LAUNCH(small_logpost_kernel<SMALL_MODEL_MIX2>, (unsigned)cdiv((int64_t)h->C * 32, 128), 128, 0, h->stream, a, 0u, h->dout); break;
        case SGM_MODEL_LOGISTIC:
            SGM_LAUNCH(logistic_logpost_kernel, h->C, 256, 0, h->stream, a, 0u, h->theta, h->dout); break;
        case SGM_MODEL_SOFTMAX_NET: {
            const int e = softmax_net_logpost(h->nn, a, 0u, h->theta, h->dout, h->stream);
            if (e != 0) rc = fail(SGM_ECUDA, "softmax net held-out: %s", cudaGetErrorString((cudaError_t)e));
            break;
        }
    }
    if (rc == SGM_OK && cudaGetLastError() != cudaSuccess) rc = fail(SGM_ECUDA, "held-out kernel launch failed");
    if (rc == SGM_OK) {
        cudaMemcpyAsync(out, h->dout, (size_t)h->C * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(SGM_ECUDA, "held-out evaluation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return rc;
}

// Pins one kernel path of a handle (tests and measurements; 0 restores the default dispatch).  Replaces the environment
// switches of the first round: nothing on the dispatch path reads the environment.
extern "C" int sgm_set_option(sgm_handle* h, const char* key, int32_t value) {
    if (!h || !key) return fail(SGM_EINVAL, "null argument");
    struct Opt { const char* key; int* slot; int max; };
    const Opt opts[] = {{"logistic_kernel", &h->opt_logistic_kernel, 1}, {"gather", &h->opt_gather, 4}, {"fullgrad", &h->opt_fullgrad, 2},
                        {"fullgrad_row0", &h->opt_row0, INT32_MAX}, {"fullgrad_row1", &h->opt_row1, INT32_MAX},
                        {"sharded_mode_finder", &h->opt_sharded_finder, 1}, {"small_kernel", &h->opt_small_kernel, 2},
                        {"mode_finder_split", &h->opt_mode_finder_split, LSP_MAX_SPLIT}, {"chain_split", &h->opt_chain_split, LSP_MAX_SPLIT}, {"sample_out", &h->opt_sample_out, 1}};
    for (const Opt& o : opts)
        if (!strcmp(key, o.key)) {
            if (value < 0 || value > o.max) return fail(SGM_EINVAL, "option %s: value %d outside [0, %d]", key, value, o.max);
            *o.slot = value;
            return SGM_OK;
        }
    if (h->nn && !strncmp(key, "nn_", 3)) {
        if (softmax_net_set_option(h->nn, key, value) == 0) return SGM_OK;
        return fail(SGM_EINVAL, "option %s = %d is not one the softmax-net kernels know", key, value);
    }
    return fail(SGM_EINVAL, "unknown option '%s'", key);
}

// ------------------------------------------------------------------------------------
// launch plan (host only)
// ------------------------------------------------------------------------------------
namespace {
struct PlanOut {
    std::string js;
    std::vector<std::string> bad;
    bool first_launch = true, first_map = true;
    std::string launches, maps;
    void launch(const char* kernel, int64_t gx, int64_t gy, int64_t gz, int threads, size_t smem, int tmem_cols, const char* note = "") {
        char b[512];
        snprintf(b, sizeof b, "%s{\"kernel\":\"%s\",\"grid\":[%lld,%lld,%lld],\"block\":%d,\"smem\":%zu,\"tmem_cols\":%d,\"note\":\"%s\"}", first_launch ? "" : ",", kernel,
                 (long long)gx, (long long)gy, (long long)gz, threads, smem, tmem_cols, note);
        launches += b; first_launch = false;
        auto say = [&](const char* why) { snprintf(b, sizeof b, "%s: %s (grid %lld x %lld x %lld, %d threads, %zu B shared)", kernel, why, (long long)gx, (long long)gy, (long long)gz, threads, smem); bad.push_back(b); };
        if (gx < 1 || gy < 1 || gz < 1) say("empty grid");
        if (gx > 2147483647ll || gy > 65535 || gz > 65535) say("grid dimension over the limit (2^31-1, 65535, 65535)");
        if (threads < 1 || threads > 1024) say("block size outside [1, 1024]");
        if (smem > 227 * 1024) say("more than 227 KB of dynamic shared memory");
        if (tmem_cols > 512) say("more than 512 TMEM columns");
    }
    void map(const TmapPlan& m, int box0, int box1, const char* swizzle) {
        char b[512];
        snprintf(b, sizeof b, "%s{\"what\":\"%s\",\"dims\":[%llu,%llu],\"pitch_bytes\":%llu,\"box\":[%d,%d],\"swizzle\":\"%s\"}", first_map ? "" : ",", m.what,
                 (unsigned long long)m.inner, (unsigned long long)m.rows, (unsigned long long)m.pitch_floats * 4, box0, box1, swizzle);
        maps += b; first_map = false;
        tmap_plan_check(m, bad);
        if (box0 < 1 || box0 > 256 || box1 < 1 || box1 > 256) bad.push_back(std::string("tensor map ") + m.what + ": boxDim outside [1, 256]");
        if ((box0 * 4) % 16) bad.push_back(std::string("tensor map ") + m.what + ": inner box is not a multiple of 16 bytes");
    }
};
}  // namespace

extern "C" int sgm_debug_plan(const sgm_plan_query* q, char* out, size_t cap) {
    if (!q || !out || cap < 2) return fail(SGM_EINVAL, "null argument");
    if (q->N < 1 || q->N >= (1ll << 31) || q->d < 1 || q->n < 1 || q->n > q->N * (q->sharded_dataset ? (q->world > 0 ? q->world : 1) : 1) || q->n_chains < 1 || q->sms < 1)
        return fail(SGM_EINVAL, "plan query: bad sizes");
    const int world = q->world > 0 ? q->world : 1, rank = q->rank;
    if (rank < 0 || rank >= world) return fail(SGM_EINVAL, "plan query: rank outside the world");
    PlanOut po;
    // a handle shaped like the query, without a device: the dispatch predicates below are the ones the real launches use
    sgm_handle hh;
    sgm_handle* h = &hh;
    h->model = q->model; h->cv = q->cv ? 1 : 0; h->N = q->N; h->d = q->d; h->n = q->n; h->C = q->n_chains; h->flags = q->flags;
    h->world = world; h->rank = rank; h->n_shards = q->sharded_dataset ? world : 1;
    const int M = h->cv ? hat_count(h) : 0;
    int64_t row0 = h->n_shards > 1 ? 0 : h->N * rank / world, row1 = h->n_shards > 1 ? h->N : h->N * (rank + 1) / world;
    if (q->row0 >= 0) row0 = q->row0;
    if (q->row1 >= 0) row1 = q->row1;
    if (row0 < 0 || row1 > h->N || row0 >= row1) return fail(SGM_EINVAL, "plan query: bad row range [%lld, %lld)", (long long)row0, (long long)row1);
    const int sms = q->sms;
    SoftmaxNetPlan np{};
    switch (h->model) {
        case SGM_MODEL_GAUSS_MEAN:
        case SGM_MODEL_MIXTURE2: {
            if (h->d > SMALL_MAXD) return fail(SGM_EINVAL, "Gaussian families support d <= %d", SMALL_MAXD);
            h->ld = h->d; h->P = h->Pst = (h->model == SGM_MODEL_MIXTURE2 ? 2 : 1) * h->d;
            const size_t dbytes = (size_t)h->N * h->ld * sizeof(float);
            if (dbytes % 16 == 0 && dbytes <= SMALL_SMEM_MAX_BYTES) {
                int64_t grid = cdiv((int64_t)h->C * 32, SMALL_SMEM_THREADS);
                if (grid > (int64_t)sms * 2) grid = (int64_t)sms * 2;
                po.launch("small_chain_kernel<smem>", grid, 1, 1, SMALL_SMEM_THREADS, small_smem_bytes(h->N, h->ld), 0, "one warp per chain; dataset staged in shared memory once per CTA (TMA bulk copies)");
            } else {
                int64_t grid = cdiv((int64_t)h->C * 32, 128);
                if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
                po.launch("small_chain_kernel", grid, 1, 1, 128, 0, 0, "one warp per chain; rows from global memory");
            }
            if (h->cv) {
                int64_t gx = cdiv(row1 - row0, 256 * 4); if (gx > sms * 8) gx = sms * 8; if (gx < 1) gx = 1;
                po.launch("small_chain_kernel (mode finder)", cdiv((int64_t)M * 32, 128), 1, 1, 128, 0, 0);
                po.launch("small_fullsum_kernel", gx, M, 1, 256, 0, 0);
            }
            break;
        }
        case SGM_MODEL_LOGISTIC: {
            h->ld = 4 * (h->d / 4 + 1);
            LogiShape s;
            if (!logi_shape(h->ld, &s)) return fail(SGM_EINVAL, "logistic: d = %d too large", h->d);
            h->Pst = s.G * s.V * 4; h->P = h->d + 1;
            h->tmap_ok = h->ld <= 128;
            ChainArgs a{}; a.n = h->n; a.C = h->C; a.cv = h->cv;
            auto chain = [&](int chains, const char* what) {
                a.C = chains;
                char note[160];
                if (logistic_use_staged(h, a)) {
                    const size_t smem = ls_smem_bytes(h->ld, h->n);
                    int per_sm = (int)((227 * 1024) / (smem + 1024)); if (per_sm > 2) per_sm = 2; if (per_sm < 1) per_sm = 1;
                    const bool g4 = logistic_can_gather4(h);
                    snprintf(note, sizeof note, "%s; rows per lane R = %d; gather = %s", what, (h->n + LS_THREADS - 1) / LS_THREADS, g4 ? "TMA tile::gather4" : "TMA bulk copy per row");
                    po.launch("logistic_staged_kernel", chains < sms * per_sm ? chains : sms * per_sm, 1, 1, LS_THREADS, smem, 0, note);
                    if (g4) po.map(TmapPlan{"dataset rows (gather4)", (uint64_t)h->ld, (uint64_t)h->N, (uint64_t)h->ld, (uint64_t)h->N * h->ld}, h->ld, 1, "none");
                } else {
                    const int threads = logi_threads(h, s);
                    const size_t smem = logi_smem_bytes(threads, s, h->n);
                    snprintf(note, sizeof note, "%s; <G, V> = <%d, %d>%s", what, s.G, s.V, h->n_shards > 1 ? "; rows resolved through the peer-shard table" : "");
                    po.launch("logistic_chain_kernel", chains < sms ? chains : sms, 1, 1, threads, smem, 0, note);
                }
            };
            chain(h->C, "sampling");
            if (h->cv) {
                const int split = split_of(h->opt_mode_finder_split, M, h->n, sms, h->ld);
                if (split > 1 && lsp_smem_bytes(s.G, s.V, h->ld, h->n, split) <= 227 * 1024) {
                    char note[160];
                    int clusters = sms / split; if (clusters > M) clusters = M;
                    snprintf(note, sizeof note, "mode finder; thread-block clusters of %d CTAs per mode (DSMEM reduction), %d minibatch rows per CTA; <G, V> = <%d, %d>", split, (h->n + split - 1) / split, s.G, s.V);
                    po.launch("logistic_split_kernel", (int64_t)clusters * split, 1, 1, LSP_THREADS, lsp_smem_bytes(s.G, s.V, h->ld, h->n, split), 0, note);
                } else {
                    chain(M, "mode finder");
                }
                const bool gathered = world > 1 && !(h->flags & SGM_FLAG_SHARED_MODE);
                const int Mg = gathered ? M * world : M;
                const int Q = h->ld / 4;
                if (logistic_fullgrad_use_tc(h, Mg)) {
                    const CvgPlan P = cvg_plan(h->N, h->d, h->ld, h->Pst, Mg, row0, row1, sms);
                    for (const std::string& b : cvg_plan_check(P)) po.bad.push_back(b);
                    po.map(P.mX, CVG_KC, CVG_TILE, "128B"); po.map(P.mTh, CVG_KC, CVG_TILE, "128B"); po.map(P.mX2, (int)P.mX2.box_inner, (int)P.mX2.box_rows, "none");
                    char note[320];
                    size_t shown = 0;
                    for (const CvgBlockPlan& B : P.blocks) {
                        if (shown++ >= 2 && &B != &P.blocks.back()) continue;   // first two and the last block are printed; all are checked
                        po.map(B.mRt, CVG_KC, CVG_TILE, "128B");
                        snprintf(note, sizeof note, "pass 1, rows [%lld, %lld): %d items x %d K stages", (long long)B.b0, (long long)B.b1, B.p1.n_items, B.p1.k_chunks);
                        po.launch("cvg_kernel", B.grid1, 1, 1, CVG_THREADS, CVG_SMEM, 512, note);
                        if (B.rowsum_warps > 0) po.launch("cvg_rowsum_kernel", cdiv(B.rowsum_warps * 32, 256), 1, 1, 256, 0, 0);
                        snprintf(note, sizeof note, "pass 2: %d items, K per split %d (from global row %d); A = %d features x 32 rows of X as stored (no transposed copy); %zu block(s), R^T %zu B",
                                 B.p2.n_items, B.p2.k_per_split, B.p2.a_k0, P.a_box_w, P.blocks.size(), P.rt_bytes);
                        po.launch("cvg_kernel", B.grid2, 1, 1, CVG_THREADS, CVG_SMEM, 512, note);
                    }
                } else if (Mg == 1 && h->ld <= 128 && ((Q & 1) || (Q % 4 == 2)) && fg_smem_bytes(h->ld, FG_SHORT_ROWS) <= 113 * 1024) {
                    const int64_t tiles = cdiv(row1 - row0, FG_SHORT_ROWS);
                    po.launch("logistic_fullgrad_stream_kernel<short>", tiles < (int64_t)sms * 2 ? tiles : (int64_t)sms * 2, 1, 1, FG_THREADS, fg_smem_bytes(h->ld, FG_SHORT_ROWS), 0, "TMA bulk tile ring");
                } else if (Mg == 1 && h->ld > 128 && (Q + 31) / 32 <= 9 && fg_smem_bytes(h->ld, FG_LONG_ROWS) <= 113 * 1024) {
                    const int64_t tiles = cdiv(row1 - row0, FG_LONG_ROWS);
                    po.launch("logistic_fullgrad_stream_kernel<long>", tiles < (int64_t)sms * 2 ? tiles : (int64_t)sms * 2, 1, 1, FG_THREADS, fg_smem_bytes(h->ld, FG_LONG_ROWS), 0, "TMA bulk tile ring");
                } else {
                    const int threads = (s.V >= 5) ? 512 : 256;
                    int64_t chunks = (int64_t)sms * 8, rpc = cdiv(row1 - row0, chunks);
                    const int64_t min_rows = 4 * (threads / s.G);
                    if (rpc < min_rows) rpc = min_rows;
                    chunks = cdiv(row1 - row0, rpc);
                    po.launch("logistic_fullgrad_kernel", chunks, Mg, 1, threads, sizeof(float) * (size_t)(1 + threads / 32) * (s.V * s.G * 4), 0, "row chunks x modes");
                }
                if (M == 1 && ((h->flags & SGM_FLAG_SHARED_MODE) || h->n_shards <= 1) && (h->d & 3) != 3)
                    po.launch("logistic_cache_rhat_kernel", cdiv(h->N * 32, 256), 1, 1, 256, 0, 0, "residual at the shared mode cached in the rows");
            }
            break;
        }
        case SGM_MODEL_SOFTMAX_NET: {
            if (softmax_net_plan(h->d, q->d_hidden, q->d_out, &np) != 0) return fail(SGM_EINVAL, "softmax net: unsupported widths");
            h->ld = 4 * (int)cdiv(h->d + q->d_out, 4); h->Pst = np.Pst;
            char note[160];
            snprintf(note, sizeof note, "forward GEMM %s; S1 ring %zu of %zu B", np.tma_fed ? "TMA-fed (gather4 + 3-D box)" : "register-staged", np.s1_ring_bytes, np.ring_capacity_bytes);
            po.launch("softmax_net_chain_kernel", h->C < sms ? h->C : sms, 1, 1, np.threads, np.smem_bytes, np.tmem_cols, note);
            if (np.tma_fed) {
                po.map(TmapPlan{"dataset inputs (gather4)", (uint64_t)h->d, (uint64_t)h->N, (uint64_t)h->ld, (uint64_t)h->N * h->ld}, np.s1_box_inner, 1, "128B");
                // 3-D parameter map: (input, hidden unit, chain); checked as its 2-D (input, hidden x chain) footprint
                po.map(TmapPlan{"B block of the chain states", (uint64_t)h->d, (uint64_t)q->d_hidden, (uint64_t)h->d, (uint64_t)h->d * q->d_hidden}, np.s1_box_inner, q->d_hidden, "128B");
                if ((np.offB & 3) || (np.Pst & 3)) po.bad.push_back("softmax net: parameter map base / chain stride not 16-byte aligned");
            }
            if (h->cv) {
                int gx = (int)((sms + M - 1) / M); if (gx < 1) gx = 1;
                po.launch("softmax_net_fullgrad_kernel", gx, M, 1, np.threads, np.smem_bytes, np.tmem_cols);
            }
            break;
        }
        default: return fail(SGM_ENOMODEL, "unknown model %d", h->model);
    }
    std::string js = "{\"ld\":" + std::to_string(h->ld) + ",\"Pst\":" + std::to_string(h->Pst) + ",\"rows\":[" + std::to_string(row0) + "," + std::to_string(row1) + "],\"launches\":[" +
                     po.launches + "],\"tensor_maps\":[" + po.maps + "],\"violations\":[";
    for (size_t i = 0; i < po.bad.size(); ++i) {
        std::string v = po.bad[i];
        for (char& c : v) if (c == '"' || c == '\\') c = '\'';
        js += (i ? ",\"" : "\"") + v + "\"";
    }
    js += "]}";
    snprintf(out, cap, "%s", js.c_str());
    return (int)po.bad.size();
}

extern "C" int64_t sgm_num_rows(const sgm_handle* h) { return h ? h->N : 0; }
extern "C" int32_t sgm_minibatch(const sgm_handle* h) { return h ? h->n : 0; }
extern "C" int32_t sgm_num_scalars(const sgm_handle* h) { return h ? h->P : 0; }
extern "C" int32_t sgm_noise_len(const sgm_handle* h) { return h ? h->base.noise_len : 0; }

extern "C" int sgm_last_timing(const sgm_handle* h, float* kernel_ms, int32_t* launches) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    if (kernel_ms) *kernel_ms = h->last_ms;
    if (launches) *launches = h->last_launches;
    return SGM_OK;
}

extern "C" int sgm_last_fullgrad_timing(const sgm_handle* h, float* kernel_ms, int64_t* bytes) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    if (kernel_ms) *kernel_ms = h->fullgrad_ms;
    if (bytes) *bytes = h->fullgrad_bytes;
    return SGM_OK;
}

extern "C" int sgm_last_collective_timing(const sgm_handle* h, float* reduce_ms, float* gather_ms) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    if (reduce_ms) *reduce_ms = h->collective_ms;
    if (gather_ms) *gather_ms = h->gather_ms;
    return SGM_OK;
}

extern "C" int sgm_copy_dataset(sgm_handle* h, int32_t entry, int64_t row0, int64_t rows, double* out) {
    if (!h || !out || row0 < 0 || rows < 0 || row0 + rows > h->N) return fail(SGM_EINVAL, "bad argument");
    if (int rc = set_device(h)) return rc;
    int col0 = 0, cols = h->d;
    if (entry == 1) {
        if (h->model == SGM_MODEL_LOGISTIC) { col0 = h->d; cols = 1; }
        else if (h->model == SGM_MODEL_SOFTMAX_NET) { col0 = h->d; cols = (int)h->tensors[2].size; }
        else return fail(SGM_EINVAL, "model has no dataset entry %d", entry);
    } else if (entry != 0) return fail(SGM_EINVAL, "model has no dataset entry %d", entry);
    DevBuf<double> stage_buf;
    CU(stage_buf.alloc((size_t)rows * cols));
    double* stage = stage_buf.p;
    SGM_LAUNCH(unpack_rows_kernel, (unsigned)cdiv(rows * cols, 256), 256, 0, h->stream, h->X, row0, rows, cols, h->ld, col0, stage);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, stage, (size_t)rows * cols * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SGM_OK;
}

extern "C" int sgm_last_fullgrad_flops(const sgm_handle* h, double* flops) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    if (flops) *flops = h->fullgrad_flops;
    return SGM_OK;
}

// measurement: the full-data gradient kernel of a *cv handle, `reps` back-to-back launches between two CUDA events
// (a single launch inside sgm_init is bracketed by launch gaps that matter for a 20-70 us kernel); state is untouched
extern "C" int sgm_bench_full_gradient(sgm_handle* h, int reps, float* ms_per_rep, int64_t* bytes) {
    if (!h || reps < 1) return fail(SGM_EINVAL, "bad argument");
    if (!h->cv || !h->inited) return fail(SGM_ESTATE, "needs an initialised *cv sampler");
    if (h->model != SGM_MODEL_LOGISTIC) return fail(SGM_EINVAL, "only the logistic family is instrumented");
    if (int rc = set_device(h)) return rc;
    int M = hat_count(h);
    const int64_t row0 = h->n_shards > 1 ? 0 : h->N * h->rank / h->world, row1 = h->n_shards > 1 ? h->N : h->N * (h->rank + 1) / h->world;
    const float* hat = h->theta_hat;
    double* own = h->gsum;
    if (fullgrad_gathered(h)) {   // kernels only: the modes of all ranks were gathered by sgm_init
        if (!h->hat_all) return fail(SGM_ESTATE, "modes not gathered yet");
        M *= h->world; hat = h->hat_all; h->gsum = h->gsum_all;
    }
    int rc = SGM_OK;
    for (int warm = 0; warm < 2 && !rc; ++warm) rc = launch_logistic_fullgrad(h, row0, row1, M, hat);
    if (!rc) CU(cudaEventRecord(h->ev0, h->stream));
    for (int r = 0; r < reps && !rc; ++r) rc = launch_logistic_fullgrad(h, row0, row1, M, hat);
    h->gsum = own;
    if (rc) return rc;
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (ms_per_rep) *ms_per_rep = ms / reps;
    if (bytes) *bytes = (row1 - row0) * (int64_t)h->ld * 4;
    return SGM_OK;
}

// pinned host buffers for result copy-out (optional: any caller-allocated buffer works, pinned ones copy faster)
extern "C" void* sgm_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void sgm_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------
// multi-GPU
// ------------------------------------------------------------------------------------
extern "C" int sgm_comm_unique_id(void* id_out, int32_t id_bytes) {
    if (!id_out || id_bytes < SGM_NCCL_ID_BYTES) return fail(SGM_EINVAL, "id buffer must hold %d bytes", SGM_NCCL_ID_BYTES);
    if (int rc = load_nccl()) return rc;
    const int rc = g_nccl.GetUniqueId(id_out);
    if (rc != 0) return fail(SGM_ECOMM, "ncclGetUniqueId failed (%d)", rc);
    return SGM_OK;
}

// ---- dataset sharded over ranks, peers mapped through CUDA IPC ------------------------------------------------------------
extern "C" int sgm_dataset_ipc_handle(sgm_handle* h, void* out, int32_t out_bytes) {
    if (!h || !out || out_bytes < (int32_t)sizeof(cudaIpcMemHandle_t)) return fail(SGM_EINVAL, "bad argument");
    if (int rc = set_device(h)) return rc;
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, h->X));
    memcpy(out, &mh, sizeof mh);
    return SGM_OK;
}

extern "C" int sgm_attach_peer_shards(sgm_handle* h, const void* handles, int32_t world, int32_t rank) {
    if (!h || !handles || world < 1 || world > 16 || rank < 0 || rank >= world) return fail(SGM_EINVAL, "bad argument");
    if (h->model != SGM_MODEL_LOGISTIC) return fail(SGM_EINVAL, "a dataset sharded over ranks is implemented for the logistic family");
    if (h->inited) return fail(SGM_ESTATE, "attach the peer shards before sgm_init");
    if ((uint64_t)h->N * (uint64_t)world > 0xFFFFFFFFull) return fail(SGM_EINVAL, "more than 2^32 rows in total");
    if (h->n_shards > 1 || h->d_shard_ptr) return fail(SGM_ESTATE, "peer shards are already attached (sgm_detach_peer_shards first)");
    if (int rc = set_device(h)) return rc;
    std::vector<const float*> table((size_t)world, nullptr);
    for (int r = 0; r < world; ++r) {
        if (r == rank) { table[r] = h->X; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, (const char*)handles + (size_t)r * sizeof mh, sizeof mh);
        void* p = nullptr;
        const cudaError_t oe = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
        if (oe != cudaSuccess) {   // unmap what was mapped so far: the handle stays usable on its local rows
            for (int q = 0; q < 16; ++q) if (h->peer_base[q]) { cudaIpcCloseMemHandle(h->peer_base[q]); h->peer_base[q] = nullptr; }
            return fail(SGM_ECUDA, "cudaIpcOpenMemHandle (rank %d's shard) failed: %s", r, cudaGetErrorString(oe));
        }
        h->peer_base[r] = p;
        table[r] = (const float*)p;
    }
    CU(cudaMalloc((void**)&h->d_shard_ptr, sizeof(const float*) * (size_t)world));
    CU(cudaMemcpy((void*)h->d_shard_ptr, table.data(), sizeof(const float*) * (size_t)world, cudaMemcpyHostToDevice));
    h->n_shards = world;
    // every rank holds h->N rows (equal shards); the chains sample from all of them, N / n is the global ratio (R/setup.R:117)
    const uint64_t n_global = (uint64_t)h->N * (uint64_t)world;
    h->base.N = (uint32_t)n_global; h->base.perm_bits = perm_bits((uint32_t)n_global);
    h->base.corr = (float)((double)n_global / (double)h->n);
    h->base.shard_ptr = h->d_shard_ptr; h->base.shard_rows = (uint32_t)h->N; h->base.n_shards = world;
    return SGM_OK;
}

// unmap the peers' shards (every rank calls this, THEN a barrier, THEN the owners may free their datasets)
extern "C" int sgm_detach_peer_shards(sgm_handle* h) {
    if (!h) return fail(SGM_EINVAL, "null handle");
    if (int rc = set_device(h)) return rc;
    CU(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < 16; ++r) if (h->peer_base[r]) { cudaIpcCloseMemHandle(h->peer_base[r]); h->peer_base[r] = nullptr; }
    if (h->d_shard_ptr) { cudaFree((void*)h->d_shard_ptr); h->d_shard_ptr = nullptr; }
    h->n_shards = 1; h->base.n_shards = 0; h->base.shard_ptr = nullptr; h->base.shard_rows = 0;
    // back to the local row space: N, the permutation width and N / n (R/setup.R:116-118) were the global ones
    h->base.N = (uint32_t)h->N; h->base.perm_bits = perm_bits((uint32_t)h->N);
    h->base.corr = (float)((double)h->N / (double)h->n);
    h->rhat_cached = false;
    h->inited = false;   // the sampler state referred to the global row space
    return SGM_OK;
}

extern "C" int sgm_comm_init(sgm_handle* h, const void* id, int32_t id_bytes, int32_t rank, int32_t world) {
    if (!h || !id || id_bytes < SGM_NCCL_ID_BYTES || world < 1 || rank < 0 || rank >= world) return fail(SGM_EINVAL, "bad argument");
    // without the flag every rank finds its own mode(s) (chain_offset differs), whatever the number of chains per rank: summing
    // partial gradients taken at different modes is meaningless
    if (h->cv && world > 1 && !(h->flags & SGM_FLAG_SHARED_MODE) && h->model != SGM_MODEL_LOGISTIC)
        return fail(SGM_EINVAL, "a row-sharded control-variate pass needs SGM_FLAG_SHARED_MODE (every rank must hold the same mode); "
                                "one mode per chain across ranks is implemented for the logistic family");
    if (int rc = load_nccl()) return rc;
    if (int rc = set_device(h)) return rc;
    if (h->comm) return fail(SGM_ESTATE, "the handle already has a communicator");
    const std::string key((const char*)id, SGM_NCCL_ID_BYTES);
    auto it = g_comms.find(key);
    if (it != g_comms.end()) {
        if (it->second.rank != rank || it->second.world != world) return fail(SGM_EINVAL, "this unique id already names a communicator with another rank / world size");
        h->comm = it->second.comm; ++it->second.refs;
    } else {
        NcclId nid;
        std::memcpy(nid.internal, id, SGM_NCCL_ID_BYTES);
        const int rc = g_nccl.CommInitRank(&h->comm, world, nid, rank);
        if (rc != 0) return fail(SGM_ECOMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        g_comms[key] = SharedComm{h->comm, 1, rank, world};
    }
    h->comm_key = key;
    h->rank = rank; h->world = world;
    return SGM_OK;
}
