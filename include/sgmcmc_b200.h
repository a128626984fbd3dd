/*
 * sgmcmc_b200 -- C ABI of the B200-native sgmcmc sampler hot path.
 *
 * This is the drop-in boundary.  The reference (STOR-i/sgmcmc,
 * an R package) has no FFI of its own: below its exported R API every numeric
 * operation is a TensorFlow op reached through reticulate (R/update.R:44,127-129).
 * The functions declared here are what the replacement R files bind with .Call
 * (see INTEGRATION.md and r-package/src/rshim.c); each one names the reference
 * interface it replaces.
 *
 * Conventions
 *   - plain C, no exceptions / longjmp across the boundary; every function returns
 *     SGM_OK (0) or a negative sgm_status, and sgm_last_error() holds the message
 *     (thread-local);
 *   - host arrays are R-native: column-major doubles.  The library copies inputs in
 *     sgm_create (the caller may free them at once) and writes outputs into
 *     caller-allocated buffers;
 *   - a handle owns all of its device memory; distinct handles are independent, one
 *     handle is not thread-safe;
 *   - "flat parameter order" = tensors in names(params) order, each in column-major
 *     element order; P = total number of scalars;
 *   - n_chains > 1 runs that many independent replicas of the reference's single
 *     chain (extension; n_chains = 1 reproduces the reference's behaviour).
 */
#ifndef SGMCMC_B200_H
#define SGMCMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGM_ABI_VERSION 1

typedef enum sgm_status {
    SGM_OK = 0,
    SGM_EINVAL = -1,    /* bad argument / unsupported shape                         */
    SGM_ECUDA = -2,     /* CUDA runtime failure, or no CUDA device (no CPU fallback) */
    SGM_ENOMODEL = -3,  /* logLik/logPrior is not one of the routed model families    */
    SGM_EDIVERGED = -4, /* "Chain diverged, try decreasing stepsize" (R/update.R:202) */
    SGM_ENOMEM = -5,
    SGM_ESTATE = -6,    /* call order violated (e.g. step before init)                */
    SGM_ECOMM = -7      /* NCCL failure                                               */
} sgm_status;

/* Model families routed to hand-written gradient kernels. */
typedef enum sgm_model {
    SGM_MODEL_GAUSS_MEAN = 1,  /* vignettes/mvGauss.Rmd:37-54, tests/testthat/test{1d,3d}Gauss.R */
    SGM_MODEL_MIXTURE2 = 2,    /* vignettes/gaussMixture.Rmd:48-76                               */
    SGM_MODEL_LOGISTIC = 3,    /* vignettes/logisticRegression.Rmd:45-57                         */
    SGM_MODEL_SOFTMAX_NET = 4  /* vignettes/nn.Rmd:58-79                                         */
} sgm_model;

/* Dynamics (R/dynamics.R:7-31, 35-69, 73-120). */
typedef enum sgm_algo { SGM_ALGO_SGLD = 1, SGM_ALGO_SGHMC = 2, SGM_ALGO_SGNHT = 3 } sgm_algo;

/* One entry of the reference's `dataset` list (R/setup.R:105-113): sliced on dim 0. */
typedef struct sgm_array {
    const char* name;
    const double* data; /* may be NULL iff sgm_config.synth_rows > 0 */
    int32_t ndim;       /* 1 or 2 */
    int32_t row_major;  /* 0: column-major (R's layout, what the reference passes); 1: row-major (C order, numpy's default).
                         * Either is packed into the device row format by a kernel; no host-side transposition is ever needed.
                         * (Occupies what was padding: size and offsets of the struct are unchanged.) */
    int64_t dims[4];
} sgm_array;

/* One entry of the reference's `params` list plus its tuning constants
 * (convertList, R/setup.R:160-170). */
typedef struct sgm_param {
    const char* name;
    const double* init; /* column-major start value, replicated to every chain */
    int32_t ndim;       /* 0 = R scalar, 1 = vector, 2 = matrix */
    int64_t dims[4];
    double stepsize;    /* epsilon (sgld) / eta (sghmc, sgnht) */
    double alpha;       /* sghmc friction; ignored otherwise   */
    double a;           /* sgnht diffusion; ignored otherwise  */
} sgm_param;

#define SGM_FLAG_SHARED_MODE 1u /* *cv with n_chains > 1: find ONE mode (chain 0's stream) and share it */

/* Everything createSGMCMC / createSGMCMCCV / *Setup take (R/setup.R:2-28,
 * R/controlVariates.R:2-25, R/sgld.R:114-123 ...). */
typedef struct sgm_config {
    int32_t abi_version; /* SGM_ABI_VERSION */
    int32_t model;       /* sgm_model */
    int32_t algo;        /* sgm_algo  */
    int32_t cv;          /* 0: sgld/sghmc/sgnht, 1: *cv */
    int32_t n_data;
    const sgm_array* data; /* model's dataset entries, in the order the family documents.  Row limits (row numbers are 32-bit on the device):
                            * < 2^31 rows per handle; with the dataset sharded over ranks (sgm_attach_peer_shards) <= 2^32 rows in the whole job
                            * (the scale-out config's 1e8 is far below both); sgm_create / sgm_attach_peer_shards return SGM_EINVAL beyond them */
    int32_t n_params;
    const sgm_param* params; /* names(params) order = update order (Gauss-Seidel, R/update.R:43-45) */
    double minibatch_size;   /* < 1: proportion of N (convertProp, R/setup.R:128-140) */
    int32_t L;               /* sghmc sub-steps (R/sghmc.R:40-41)          */
    double opt_stepsize;     /* *cv: SGD step of the mode finder (cV.R:87) */
    int64_t n_iters_opt;     /* *cv: mode-finder iterations (cV.R:42)      */
    int32_t seeded;          /* 0 = seed=NULL (library draws one from the clock) */
    uint64_t seed;
    int32_t n_chains;        /* >= 1 */
    int64_t chain_offset;    /* global index of this handle's first chain (multi-GPU chain groups); 0 otherwise */
    int32_t device;          /* CUDA device ordinal */
    uint32_t flags;
    /* model constants (what the user's closures hard-code) */
    int32_t use_prior;       /* 0 = logPrior = NULL (flat, R/setup.R:119-123)              */
    double prior_sd;         /* Gaussian families: Normal(0, prior_sd) on every mean entry   */
    const double* precision; /* GAUSS_MEAN: d x d Sigma^{-1} (column-major), NULL = identity */
    /* synthetic dataset generated on the device (too large to ship from the host) */
    int64_t synth_rows;      /* > 0: ignore data[].data, generate N = synth_rows rows        */
    int64_t synth_dim;       /* LOGISTIC: d                                                   */
    uint64_t synth_seed;
} sgm_config;

typedef struct sgm_handle sgm_handle;

/* ---- lifecycle ------------------------------------------------------------------- */

/* *Setup (R/sgld.R:114-123 etc.): validates, uploads the dataset (column-major double
 * -> row-major float32, 16-byte-aligned row stride), allocates chain state. */
int sgm_create(const sgm_config* cfg, sgm_handle** out);

/* Progress of the *cv mode finder.  The reference's getMode calls checkOptDivergence every 100 optimiser iterations
 * (R/controlVariates.R:45-47,114-124): the log posterior estimate at the current mode estimate on a FRESH minibatch; NaN stops with
 * "Chain diverged", otherwise -- when verbose -- "Iteration: i  Log posterior estimate: v" is printed.  sgm_init does the same on the
 * device in launches of <= 100 iterations and hands the value(s) to fn (phase = 1, iter = optimiser iterations done, logpost[n] =
 * one estimate per mode); the R shim / Python mirror prints the reference's line from it.  fn runs on the calling thread. */
typedef void (*sgm_progress_fn)(void* user, int32_t phase, int64_t iter, const double* logpost, int32_t n);
int sgm_set_progress(sgm_handle* h, sgm_progress_fn fn, void* user);

/* initSess (R/update.R:126-140): draws the initial momenta (R/dynamics.R:47,84); for
 * *cv runs getMode + calcFullGrads and sets theta := theta_hat (cV.R:37-52,105-111).
 * Returns SGM_EDIVERGED with "Chain diverged" if the optimiser produced NaN (cV.R:119). */
int sgm_init(sgm_handle* h);

/* Deterministic validation mode of initSess: caller supplies standard normals for the
 * initial momenta ([n_chains][P], may be NULL = zeros) and, for *cv, the 0-based
 * minibatch rows of every optimiser iteration ([n_chains][n_iters_opt][n]). */
int sgm_init_deterministic(sgm_handle* h, const float* momentum_noise, const int32_t* opt_idx);

void sgm_destroy(sgm_handle* h);

/* ---- stepping -------------------------------------------------------------------- */

/* sgmcmcStep (R/update.R:40-81) with device-side minibatch sampling and noise. */
int sgm_step(sgm_handle* h);

/* sgmcmcStep in the deterministic validation mode: idx = 0-based rows
 * [n_chains][n_sub][n] (n_sub = L for sghmc, else 1); noise = standard normals
 * [n_chains][noise_len] in flat parameter order, noise_len = P (sgld, sgnht) or
 * (L+1)*P (sghmc: refresh block, then one block per sub-step).  Scale factors
 * (sqrt(2 eps) etc., R/dynamics.R) are applied on the device. */
int sgm_step_deterministic(sgm_handle* h, const int32_t* idx, const float* noise);

/* runSGMCMC's loop (R/update.R:154-161) on the device: n_iters steps in one launch;
 * store != 0 appends every post-update state to the device-resident chain
 * (storeState, R/storage.R:54-60).  Returns SGM_EDIVERGED if a state went non-finite.
 * Callers loop in chunks (<= 100) to print progress at the reference's cadence. */
int sgm_run(sgm_handle* h, int64_t n_iters, int store);

/* runSGMCMC (R/update.R:145-163) in ONE call: n_iters sgmcmcSteps, every post-update state stored (storeState, R/storage.R:54-60) and
 * the whole chain written to out in the layout of sgm_copy_chain.  The chain leaves the device WHILE later iterations run: device
 * storage is two blocks of <= check_every iterations; block k is transposed and copied (as float32: half the PCIe bytes) on a second
 * stream and widened to double by host threads while block k + 1 is being sampled, so device memory stays bounded and the copy-out
 * hides behind the kernel.  Every check_every iterations (100 in the reference; 0 = never) checkDivergence (R/update.R:197-207) runs on
 * the device: estLogPost on a fresh minibatch; NaN returns SGM_EDIVERGED ("Chain diverged, try decreasing stepsize"), otherwise the
 * values go to the progress callback (sgm_set_progress, phase = 2, one estimate per chain).  The device-resident chain of sgm_run
 * (sgm_chain_length / sgm_copy_chain) is not touched.
 * When `out` is page-locked host memory (sgm_host_alloc, cudaHostAlloc, cudaHostRegister) the GPU writes the doubles into it directly,
 * block by block, and no host thread touches the chain; a pageable array (what R allocates) takes the staged path described above. */
int sgm_sample(sgm_handle* h, int64_t n_iters, int32_t check_every, double* out);

/* ---- reading results ------------------------------------------------------------- */

/* getParams (R/storage.R:37-39): [n_chains][P] doubles, flat parameter order. */
int sgm_get_params(sgm_handle* h, double* out);
/* Momenta [n_chains][P] and thermostats [n_chains][n_params] (validation). */
int sgm_get_state(sgm_handle* h, double* momentum, double* thermostat);
/* *cv: mode estimate and full-data gradient at it, each [n_chains][P]. */
int sgm_get_mode(sgm_handle* h, double* theta_hat, double* grad_hat);

/* State dump / restore (the reference leaves persistence to the user of the step-by-step API -- NEWS.md:19, vignettes/nn.Rmd:11 --
 * who can only save getParams output; a device-resident sampler also carries momenta, thermostats, the mode and its gradient, and
 * its position in the random streams).  sgm_save_state writes an opaque blob: call it with buf = NULL to learn the size (*need).
 * sgm_load_state restores it into a handle created with the same configuration and dataset (checked: model, algorithm, chains,
 * parameter count, rows, minibatch) -- without sgm_init, i.e. without repeating the mode search -- after which sgm_step / sgm_run /
 * sgm_sample continue the run bit for bit (the streams are keyed by seed, chain and iteration; the seed travels in the blob). */
int sgm_save_state(sgm_handle* h, void* buf, size_t cap, size_t* need);
int sgm_load_state(sgm_handle* h, const void* buf, size_t bytes);
int64_t sgm_chain_length(const sgm_handle* h);
/* One copy-out of the stored chain (replaces initStorage/updateStorage,
 * R/storage.R:42-80).  Layout: for chain c, for parameter p: a column-major
 * (iters, dims_p...) double array; arrays are concatenated p-major inside c. */
int sgm_copy_chain(sgm_handle* h, double* out);
int sgm_reset_chain(sgm_handle* h);
/* Optional page-locked host buffers for sgm_copy_chain / sgm_get_params outputs (faster device-to-host copies). */
void* sgm_host_alloc(size_t bytes);
void sgm_host_free(void* p);
/* Pre-allocates device storage for n_iters more stored states (initStorage, R/storage.R:42-51). */
int sgm_reserve_chain(sgm_handle* h, int64_t n_iters);

/* estLogPost on a fresh device-drawn minibatch, one value per chain
 * (checkDivergence, R/update.R:197-207). */
int sgm_est_logpost(sgm_handle* h, double* out);
int sgm_est_logpost_deterministic(sgm_handle* h, const int32_t* idx, double* out);

/* Validation: the gradient estimate getGradients.{sgmcmc,sgmcmccv} would produce at the
 * current state on the given rows (R/setup.R:36-62): [n_chains][P]. */
int sgm_grad_estimate(sgm_handle* h, const int32_t* idx, double* out);

/* Held-out log-likelihood of the current state on caller data (replaces the user-built
 * TF tensors of vignettes/nn.Rmd:113-123): data as in sgm_config.data, out[n_chains]. */
int sgm_heldout_loglik(sgm_handle* h, const sgm_array* data, int32_t n_data, double* out);

/* ---- introspection / measurement -------------------------------------------------- */

int64_t sgm_num_rows(const sgm_handle* h);      /* N */
int32_t sgm_minibatch(const sgm_handle* h);     /* n after convertProp */
int32_t sgm_num_scalars(const sgm_handle* h);   /* P */
int32_t sgm_noise_len(const sgm_handle* h);
/* Device time (CUDA events on the launch stream) and launch count of the last
 * sgm_init / sgm_step* / sgm_run call's kernels. */
int sgm_last_timing(const sgm_handle* h, float* kernel_ms, int32_t* launches);
/* Device time of the last control-variate full-data gradient pass and the bytes it read. */
int sgm_last_fullgrad_timing(const sgm_handle* h, float* kernel_ms, int64_t* bytes);
/* Re-runs the full-data gradient kernel of an initialised logistic *cv handle `reps` times between two CUDA events
 * (measurement only; sampler state is untouched). */
int sgm_bench_full_gradient(sgm_handle* h, int reps, float* ms_per_rep, int64_t* bytes);
/* 4 N d M flops of the last full-data gradient when it ran as tensor-core contractions (logistic family, one mode per
 * chain: calcFullGrads of R/controlVariates.R:105-111 for M independent chains at once); 0 when the streaming path ran. */
int sgm_last_fullgrad_flops(const sgm_handle* h, double* flops);
/* Pins a kernel path of this handle (tests, A/B measurements); value 0 restores the default dispatch.  Keys:
 *   "logistic_kernel"  1 = generic warp-group kernel also for short rows
 *   "gather"           short-row kernel: 1 = 16-byte cp.async, 2 = one TMA bulk copy per row, 3 = mixed, 4 = TMA tile::gather4
 *   "fullgrad"         logistic full-data gradient: 1 = tensor-core contractions, 2 = SIMT row-chunk x mode grid
 *   "fullgrad_row0", "fullgrad_row1"   row1 > 0: the full-data pass sums rows [row0, row1) only (what one rank of a row-sharded
 *                      job computes), without a communicator: single-GPU pre-flight of arbitrary shard starts
 *   "small_kernel"     Gaussian families: 1 = gather rows from global memory even when the dataset fits in shared memory,
 *                      2 = stage the dataset in shared memory even for launches too short to amortise it
 *   "sharded_mode_finder"  1 = the *cv mode finder sharded by data rows: every SGD step each rank sums the gradients of its part of the
 *                      minibatch (drawn from its own rows) and the P-vector is all-reduced (needs sgm_comm_init and
 *                      SGM_FLAG_SHARED_MODE; default: every rank runs the same finder on the rows it can reach, no communication)
 *   "mode_finder_split"  logistic *cv mode finder: 1 = one CTA per mode always; 2, 4, 8 = a thread-block cluster of that many CTAs per
 *                      mode, each summing its part of every minibatch, partial sums exchanged through distributed shared memory
 *                      (default: the largest of these that fits modes x CTAs on the device and leaves every CTA >= 16 rows)
 *   "chain_split"      logistic sampling kernels, same values: with few chains each chain gets a thread-block cluster (its CTAs split every
 *                      minibatch, exchange partial sums through distributed shared memory and keep identical replicas of the state), which
 *                      shortens the latency of a step -- the regime of the reference's own usage, ONE chain; with many chains every SM has
 *                      a chain of its own and the one-CTA kernels run
 *   "sample_out"       sgm_sample: 1 = stage the chain as float32 and widen it on the host even when the caller's array is page-locked
 *                      (default: a page-locked array is written by the GPU directly, as doubles)
 *   "nn_update"        softmax net: 1 = register-staged update of every tensor (default: streamed through shared memory)
 *   "nn_stage1"        softmax net: 1 = register-staged forward GEMM operands (default: TMA-fed)
 * The library never reads environment variables on the dispatch path. */
int sgm_set_option(sgm_handle* h, const char* key, int32_t value);
/* Launch plan of a configuration WITHOUT touching a device (works in a process that has no GPU): every kernel of initSess /
 * sgmcmcStep with its grid, block, shared memory and tensor memory, every TMA tensor map with its extents, and the violations of
 * the rules the driver / the TMA unit enforce (16-byte aligned box starts and strides, box <= extent of the allocation, <= 227 KB
 * of shared memory, <= 512 TMEM columns, grid limits).  The library runs the same check before it launches the tensor-core
 * full-gradient pass; this entry point lets a test sweep full-size configurations, world sizes and shard starts on a CPU.
 * Writes a JSON object into out (NUL-terminated, truncated at cap); returns the number of violations (>= 0) or a negative
 * sgm_status. */
typedef struct sgm_plan_query {
    int32_t model;           /* sgm_model */
    int32_t cv;
    int64_t N;               /* rows held by this process */
    int32_t d;               /* features / inputs */
    int32_t d_hidden, d_out; /* SOFTMAX_NET only */
    int32_t n;               /* minibatch rows */
    int32_t n_chains;
    uint32_t flags;          /* SGM_FLAG_SHARED_MODE */
    int32_t world, rank;     /* row-sharded control-variate pass (sgm_comm_init); world = 1: one process */
    int32_t sharded_dataset; /* 1: sgm_attach_peer_shards (every rank holds N rows) */
    int64_t row0, row1;      /* full-gradient row range override (simulated shard starts); -1 = this rank's default range */
    int32_t sms;             /* streaming multiprocessors (148 on B200) */
} sgm_plan_query;
int sgm_debug_plan(const sgm_plan_query* q, char* out, size_t cap);
/* Debug: per-phase clock64() totals of the logistic short-row kernel (thread 0 of every CTA). */
int sgm_debug_phase_cycles(sgm_handle* h, int enable, unsigned long long* out, int n);
/* Copies rows [row0, row0+rows) of the device dataset back as column-major doubles
 * (entry k of sgm_config.data); lets tests feed device-generated data to the oracle. */
int sgm_copy_dataset(sgm_handle* h, int32_t entry, int64_t row0, int64_t rows, double* out);

/* ---- multi-GPU (row-sharded control-variate pass) -------------------- */

#define SGM_NCCL_ID_BYTES 128
/* Rank 0 fills id (an ncclUniqueId); every rank then calls sgm_comm_init with the same id. */
int sgm_comm_unique_id(void* id_out, int32_t id_bytes);
/* Communicators are process-level: handles given the SAME unique id share one ncclComm (created by the first call; it lives until
 * the process exits), so a job pays ncclCommInitRank once, not once per sampler.
 * After this, sgm_init computes the full-data gradient on rows [rank*N/world, (rank+1)*N/world)
 * and all-reduces (sum, float32) the [n_chains][P] result with NCCL. */
int sgm_comm_init(sgm_handle* h, const void* id, int32_t id_bytes, int32_t rank, int32_t world);
/* Device time of the collectives of the last control-variate full-data pass: the reduction of the sums (all-reduce, or
 * reduce-scatter with one mode per chain) and the all-gather of the modes (0 when there was none). */
int sgm_last_collective_timing(const sgm_handle* h, float* reduce_ms, float* gather_ms);
/* Dataset sharded over ranks (one process per GPU, equal shards of N rows each; new: the reference is single-process).
 * sgm_dataset_ipc_handle writes the CUDA IPC handle (64 bytes) of this rank's device dataset; after an all-gather of the
 * handles, sgm_attach_peer_shards maps the peers' shards: the chains then sample minibatches from all world x N rows
 * (remote rows are loaded over NVLink / NVSwitch), N / n is the global ratio, and the control-variate pass sums this rank's
 * shard (needs sgm_comm_init for the reduction).  Logistic family, before sgm_init. */
#define SGM_IPC_HANDLE_BYTES 64
int sgm_dataset_ipc_handle(sgm_handle* h, void* out, int32_t out_bytes);
int sgm_attach_peer_shards(sgm_handle* h, const void* handles, int32_t world, int32_t rank);
/* Unmaps the peers' shards: every rank calls it, then a barrier, then the handles may be destroyed in any order. */
int sgm_detach_peer_shards(sgm_handle* h);

const char* sgm_last_error(void);
const char* sgm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SGMCMC_B200_H */
