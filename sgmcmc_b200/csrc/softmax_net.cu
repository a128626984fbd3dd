// 784-100-10 softmax-softmax network family (vignettes/nn.Rmd:58-79): one CTA per chain,
// one persistent launch for `iters` sgmcmcSteps.
//
// Replaces, for this family, dataFeed/dataSelect (R/update.R:166-194), tf$gradients of
// estLogPost (R/setup.R:36-62,116-125), declareDynamics.* (R/dynamics.R:7-120) and the
// per-tensor Session.run loops of sgmcmcStep.* (R/update.R:40-81).
//
// Forward / backward, per minibatch row block (<= 128 rows):
//   S1  Z1 = X B + b,  H = softmax_row(Z1)                       [n x 784] x [784 x 100]
//   S2  Z2 = H A + a,  Pr = softmax_row(Z2),  dZ2 = Y - Pr * rowsum(Y)
//   S3  grad A = H^T dZ2,  grad a = colsum(dZ2)
//   S4  dH = dZ2 A^T,  dZ1 = H * (dH - rowsum(dH * H)),  grad b = colsum(dZ1)
//   S5  grad B = X^T dZ1                                          [784 x n] x [n x 100]
// The reference updates the 8 parameter tensors one Session.run at a time (Gauss-Seidel,
// R/update.R:43-45,60-64): A, B, a, b, lambdaA, lambdaB, lambdaa, lambdab.  Because Z1/H
// depend only on (B, b), the exact sequence needs S1 twice and S5 once per sub-step:
//   S1 S2 S3 -> A;  S2 S4 S5 -> B;  S1 S2 S3 -> a;  S2 S4 -> b;  lambdas (prior only).
//
// Internal parameter layout per chain (float32, stride Pst): A[k][l] row-major, B[j][k]
// row-major, a, b, 4 lambdas; the flat (ABI) order keeps R's column-major element order
// inside A and B.  Dynamics noise is indexed by INTERNAL element (Philox block = 4
// consecutive internal elements) when drawn on the device, by flat index when supplied.
// Split over softmax_net_common.cuh (sizes, shared-memory layout), softmax_net_stages.cuh (S1..S5, gradient evaluation) and
// softmax_net_update.cuh (dynamics); this file holds the kernels and the host side.
#include "softmax_net_update.cuh"

namespace sgm {

__global__ void __launch_bounds__(NN_THREADS, 1) softmax_net_chain_kernel(const ChainArgs a, const NNDims dm, float* Gall, float* Ghatall,
                                                                          const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmTh,
                                                                          const __grid_constant__ CUtensorMap tmHat, int use_tma) {
    SGM_DYN_SMEM(float, smem_f, 16);
    NNSmem sm;
    nn_carve(dm, smem_f, sm);
    const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
    TcCtx tc;
    nn_tc_init(sm, tc);
    if (threadIdx.x == 0) nn_t_prev[blockIdx.x & 1023] = clock64();
    for (int c = blockIdx.x; c < a.C; c += gridDim.x) {
        NNChainState st;
        st.th = a.theta + (size_t)c * a.Pst; st.mo = a.mom + (size_t)c * a.Pst;
        st.hat = a.cv ? a.theta_hat + (size_t)c * a.hat_stride : nullptr;
        st.ghat = a.cv ? a.g_hat + (size_t)c * a.hat_stride : nullptr;
        st.G = Gall + (size_t)c * a.Pst; st.Ghat = Ghatall ? Ghatall + (size_t)c * a.Pst : nullptr;
        for (int t = 0; t < MAX_TENSORS; ++t) st.zeta[t] = a.thermo[c * MAX_TENSORS + t];
        st.ssq_valid = 0u; st.ssq_hat_valid = 0u;
        st.fast = seidel && !a.cv && nn_stream_updates;
        if (st.fast) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int si = (int)threadIdx.x + NN_THREADS * k;
                st.mo_reg[k] = si < nn_small_count(dm) ? st.mo[nn_small_elem(dm, si)] : 0.f;
            }
            for (int w4 = 0; w4 < 4; ++w4) { st.lam_th[w4] = st.th[dm.offl + w4]; st.lam_mo[w4] = st.mo[dm.offl + w4]; }
        }
        bool small_valid = false;   // shared copies of A / a / b belong to this chain's theta
        const NNMaps mpT{use_tma ? &tmX : nullptr, &tmTh, c}, mpH{use_tma ? &tmX : nullptr, &tmHat, a.hat_stride ? c : 0};
        for (int it = 0; it < a.iters; ++it) {
            const uint64_t step = a.step0 + (uint64_t)it;
            if (a.algo == ALGO_SGHMC) {  // refresh (R/dynamics.R:62)
                for (int e = threadIdx.x; e < a.P; e += NN_THREADS) {
                    const int t = nn_tensor_of(dm, e);
                    const float z = a.det ? a.noise[(size_t)c * a.noise_len + nn_flat_of_internal(dm, e)]
                                          : normal_at((uint32_t)e, a.chain0 + (uint32_t)c, word2_of(step, a.L, 0), TAG_NOISE, a.k0, a.k1);
                    st.mo[e] = 0.5f * (a.tc.sqrt_2eps[t] * z);
                }
                if (st.fast) {   // the register copies of the small momenta take the same refreshed values
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int si = (int)threadIdx.x + NN_THREADS * k;
                        if (si < nn_small_count(dm)) {
                            const int e = nn_small_elem(dm, si), t = nn_tensor_of(dm, e);
                            const float z = a.det ? a.noise[(size_t)c * a.noise_len + nn_flat_of_internal(dm, e)]
                                                  : normal_at((uint32_t)e, a.chain0 + (uint32_t)c, word2_of(step, a.L, 0), TAG_NOISE, a.k0, a.k1);
                            st.mo_reg[k] = 0.5f * (a.tc.sqrt_2eps[t] * z);
                        }
                    }
                    for (int w4 = 0; w4 < 4; ++w4) {
                        const int e = dm.offl + w4;
                        const float z = a.det ? a.noise[(size_t)c * a.noise_len + nn_flat_of_internal(dm, e)]
                                              : normal_at((uint32_t)e, a.chain0 + (uint32_t)c, word2_of(step, a.L, 0), TAG_NOISE, a.k0, a.k1);
                        st.lam_mo[w4] = 0.5f * (a.tc.sqrt_2eps[4 + w4] * z);
                    }
                }
                __syncthreads();
            }
            for (int l = 0; l < a.L; ++l) {
                const int sub = l + 1;
                const Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), a.index_tag, a.k0, a.k1);
                const int32_t* idx = a.det ? a.idx + (((size_t)c * a.iters + it) * a.L + l) * a.n : nullptr;
                const Perm* pp = a.det ? nullptr : &perm;
                if (a.cv) nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.hat, NEED_ALL, false, false, st.Ghat, tc, mpH);
                if (seidel && st.fast) {
                    // the small tensors are updated in shared memory / registers (written through): nothing to reload between evaluations
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_A, false, false, st.G, tc, mpT, small_valid ? 0 : LOAD_ALL);
                    small_valid = true;
                    nn_update_small(a, dm, sm, st, 0, c, step, sub);
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_B, true, false, st.G, tc, mpT, 0);
                    nn_update_tensor(a, dm, sm, st, 1, c, step, sub, nullptr);
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_a, false, false, st.G, tc, mpT, 0);
                    nn_update_small(a, dm, sm, st, 2, c, step, sub);
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_b, true, false, st.G, tc, mpT, 0);
                    nn_update_small(a, dm, sm, st, 3, c, step, sub);
                    for (int t = 4; t < 8; ++t) nn_update_lambda(a, dm, sm, st, t, c, step, sub);
                } else if (seidel) {
                    // A   (shared A / a / b are current from the previous sub-step except b, unless the mode's were loaded)
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_A, false, false, st.G, tc, mpT, (a.cv || !small_valid) ? LOAD_ALL : LOAD_b);
                    small_valid = true;
                    nn_update_tensor(a, dm, sm, st, 0, c, step, sub, nullptr);
                    // B: H still valid (depends on B, b only)
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_B, true, false, st.G, tc, mpT, LOAD_A);
                    nn_update_tensor(a, dm, sm, st, 1, c, step, sub, nullptr);
                    // a: B changed -> recompute H
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_a, false, false, st.G, tc, mpT, 0);
                    nn_update_tensor(a, dm, sm, st, 2, c, step, sub, nullptr);
                    // b
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_b, true, false, st.G, tc, mpT, LOAD_a);
                    nn_update_tensor(a, dm, sm, st, 3, c, step, sub, nullptr);
                    for (int t = 4; t < 8; ++t) nn_update_tensor(a, dm, sm, st, t, c, step, sub, nullptr);
                } else {
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_ALL, false, false, st.G, tc, mpT);
                    float uu[MAX_TENSORS];
                    // All tensors move from ONE evaluation.  The mode finder writes theta in place, so the
                    // precisions (whose gradient reads ||W||^2) go first and the weights see the OLD precisions.
                    float lam_old[4], lam_new[4];
                    for (int t = 0; t < 4; ++t) lam_old[t] = st.th[dm.offl + t];
                    for (int t = 7; t >= 4; --t) nn_update_tensor(a, dm, sm, st, t, c, step, sub, &uu[t]);
                    if (a.algo == ALGO_OPT) {
                        for (int t = 0; t < 4; ++t) lam_new[t] = st.th[dm.offl + t];
                        __syncthreads();
                        if (threadIdx.x == 0) for (int t = 0; t < 4; ++t) st.th[dm.offl + t] = lam_old[t];
                        __syncthreads();
                    }
                    for (int t = 3; t >= 0; --t) nn_update_tensor(a, dm, sm, st, t, c, step, sub, &uu[t]);
                    if (a.algo == ALGO_OPT) {
                        if (threadIdx.x == 0) for (int t = 0; t < 4; ++t) st.th[dm.offl + t] = lam_new[t];
                        __syncthreads();
                    }
                    if (a.algo == ALGO_SGNHT) {  // theta += u, then thermostats (R/update.R:75-80)
                        for (int e = threadIdx.x; e < a.P; e += NN_THREADS) st.th[e] += st.mo[e];
                        for (int t = 0; t < 8; ++t) {
                            int lo, hi;
                            nn_tensor_range(dm, t, lo, hi);
                            st.zeta[t] += uu[t] / (float)(hi - lo) - a.tc.eps[t];
                        }
                        __syncthreads();
                    }
                }
            }
            // the last updates of the step (the precisions: thread 0 writes them through to global memory) must be visible to the
            // threads that store the state below -- without this barrier the stored chain carried the PREVIOUS step's lambdaB / lambdaa /
            // lambdab while getParams was right (a race found by the emulated build: tests/test_emu_parity.py::test_streamed_sampling)
            __syncthreads();
            bool bad = false;
            for (int e = threadIdx.x; e < a.P; e += NN_THREADS) {
                const float v = st.th[e];
                bad |= !isfinite(v);
                if (a.store) a.store[((size_t)(a.store_pos + it) * a.C + c) * a.P + nn_flat_of_internal(dm, e)] = v;
            }
            if (bad) *a.diverged = 1;
            __syncthreads();
            NN_TICK(7)
        }
        if (threadIdx.x == 0)
            for (int t = 0; t < MAX_TENSORS; ++t) a.thermo[c * MAX_TENSORS + t] = st.zeta[t];
    }
    nn_tc_fini(tc);
}

// full-data gradient partials: CTA (x, m) sums rows [row0 + x*rows_per, ...) at theta_hat[m] into part[(m*gx + x)][Pst]
__global__ void __launch_bounds__(NN_THREADS, 1) softmax_net_fullgrad_kernel(const ChainArgs a, const NNDims dm, int64_t row0, int64_t row1,
                                                                             int rows_per, const float* __restrict__ theta_hat, float* __restrict__ part,
                                                                             const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmHat, int use_tma) {
    SGM_DYN_SMEM(float, smem_f, 16);
    NNSmem sm;
    nn_carve(dm, smem_f, sm);
    const int m = blockIdx.y;
    const int64_t lo = row0 + (int64_t)blockIdx.x * rows_per;
    int64_t hi = lo + rows_per; if (hi > row1) hi = row1;
    float* G = part + ((size_t)m * gridDim.x + blockIdx.x) * a.Pst;
    if (lo >= hi) { for (int e = threadIdx.x; e < a.Pst; e += NN_THREADS) G[e] = 0.f; return; }
    TcCtx tc;
    nn_tc_init(sm, tc);
    nn_eval(dm, sm, a.X, a.ld, (int)(hi - lo), nullptr, nullptr, lo, theta_hat + (size_t)m * a.Pst, NEED_ALL, false, false, G, tc,
            NNMaps{use_tma ? &tmX : nullptr, &tmHat, m});
    for (int e = dm.offl + threadIdx.x; e < a.Pst; e += NN_THREADS) G[e] = 0.f;
    nn_tc_fini(tc);
}

__global__ void softmax_net_part_reduce(const float* __restrict__ part, int gx, int Pst, double* __restrict__ gsum, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int m = i / Pst, e = i % Pst;
    double s = 0.0;
    for (int x = 0; x < gx; ++x) s += (double)part[((size_t)m * gx + x) * Pst + e];
    gsum[i] = s;
}

__global__ void __launch_bounds__(NN_THREADS) softmax_net_finish_kernel(const ChainArgs a, const NNDims dm, const double* __restrict__ gsum,
                                                                        const float* __restrict__ theta_hat, float* __restrict__ g_hat,
                                                                        float* __restrict__ theta) {
    __shared__ float red[64];
    const int m = blockIdx.x;
    const float* th = theta_hat + (size_t)m * a.Pst;
    float ssq[4];
    for (int t = 0; t < 4; ++t) {
        int lo, hi;
        nn_tensor_range(dm, t, lo, hi);
        ssq[t] = nn_sumsq(th + lo, hi - lo, red);
    }
    for (int e = threadIdx.x; e < a.P; e += NN_THREADS) {
        const int t = nn_tensor_of(dm, e);
        float g = (t < 4) ? (float)gsum[(size_t)m * a.Pst + e] : 0.f;
        if (a.use_prior) g += nn_prior_grad(dm, th, e, t, t >= 4 ? ssq[t - 4] : 0.f);
        g_hat[(size_t)m * a.Pst + e] = g;
        theta[(size_t)m * a.Pst + e] = th[e];
    }
}

__global__ void __launch_bounds__(NN_THREADS, 1) softmax_net_logpost_kernel(const ChainArgs a, const NNDims dm, uint32_t word2,
                                                                            const float* __restrict__ theta, float* Gscratch,
                                                                            double* __restrict__ out, const __grid_constant__ CUtensorMap tmX,
                                                                            const __grid_constant__ CUtensorMap tmTh, int use_tma) {
    SGM_DYN_SMEM(float, smem_f, 16);
    NNSmem sm;
    nn_carve(dm, smem_f, sm);
    const int c = blockIdx.x;
    const float* th = theta + (size_t)c * a.Pst;
    const Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2, TAG_CHECK_INDEX, a.k0, a.k1);
    const int32_t* idx = (a.det == 1) ? a.idx + (size_t)c * a.n : nullptr;   // det == 2: contiguous rows 0..n-1 (held-out set)
    TcCtx tc;
    nn_tc_init(sm, tc);
    float ll = nn_eval(dm, sm, a.X, a.ld, a.n, a.det ? nullptr : &perm, idx, 0, th, 0, false, true, Gscratch + (size_t)c * a.Pst, tc,
                       NNMaps{use_tma ? &tmX : nullptr, &tmTh, c});
    ll = nn_block_sum(ll, sm.red);
    float lp = 0.f;
    if (a.use_prior) {
        const float LOG_2PI = 1.8378770664093453f;
        for (int t = 0; t < 4; ++t) {
            int lo, hi;
            nn_tensor_range(dm, t, lo, hi);
            const float ssq = nn_sumsq(th + lo, hi - lo, sm.red);
            const float lam = th[dm.offl + t];
            lp += (float)(hi - lo) * (0.5f * logf(lam) - 0.5f * LOG_2PI) - 0.5f * lam * ssq - lam;
        }
    }
    if (threadIdx.x == 0) out[c] = (double)(lp + a.corr * ll);
    nn_tc_fini(tc);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*NNEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static NNEncodeFn nn_encode_fn() {
    static NNEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { cudaGetLastError(); return nullptr; }
        fn = (NNEncodeFn)p;
    }
    return fn;
}

// dataset map: 2-D (input column, row), box = 32 columns x 1 row (tile::gather4 fetches four such rows), 128-byte swizzle;
// the column extent is din, so the K tail past the last input (labels, padding) reads as zeros
static const CUtensorMap* nn_x_map(SoftmaxNetWork* w, const float* X, long long N, int ld) {
    for (int i = 0; i < w->n_maps; ++i)
        if (w->maps[i].base == X && w->maps[i].rows == N && w->maps[i].stride == ld) return &w->maps[i].map;
    NNEncodeFn fn = nn_encode_fn();
    if (!fn || !X) return nullptr;
    if (w->n_maps == 8) w->n_maps = 0;
    SoftmaxNetWork::MapEntry& e = w->maps[w->n_maps];
    const cuuint64_t gdim[2] = {(cuuint64_t)w->dm.din, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)S1_KC, 1};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return nullptr;
    e.base = X; e.rows = N; e.stride = ld;
    ++w->n_maps;
    return &e.map;
}

// parameter map: 3-D (input j, hidden unit k, chain) over the B block of a [count][stride] state array, box = 32 x dh x 1
static const CUtensorMap* nn_b_map(SoftmaxNetWork* w, const float* state, long long count, long long stride) {
    if (stride == 0 || count < 1) { count = 1; stride = 4 * ((w->dm.P + 3) / 4); }
    for (int i = 0; i < w->n_maps; ++i)
        if (w->maps[i].base == state && w->maps[i].rows == -count && w->maps[i].stride == stride) return &w->maps[i].map;
    NNEncodeFn fn = nn_encode_fn();
    if (!fn || !state) return nullptr;
    if (w->n_maps == 8) w->n_maps = 0;
    SoftmaxNetWork::MapEntry& e = w->maps[w->n_maps];
    const cuuint64_t gdim[3] = {(cuuint64_t)w->dm.din, (cuuint64_t)w->dm.dh, (cuuint64_t)count};
    const cuuint64_t gstride[2] = {(cuuint64_t)w->dm.din * 4, (cuuint64_t)stride * 4};
    const cuuint32_t box[3] = {(cuuint32_t)S1_KC, (cuuint32_t)w->dm.dh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(state) + w->dm.offB, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return nullptr;
    e.base = state; e.rows = -count; e.stride = stride;
    ++w->n_maps;
    return &e.map;
}

static bool nn_tma_enabled(const SoftmaxNetWork* w) {
    if (w->opt_stage1 == 1) return false;
    return (w->dm.offB % 4) == 0 && (size_t)nn_s1_ring_bytes(w->dm) <= (size_t)(2 * NN_RB * w->dm.dh + nn_tile_floats_all(w->dm)) * 4;
}

static bool nn_dims(NNDims& dm, int d_in, int d_hidden, int d_out) {
    if (d_in % 4 || d_hidden % 4 || d_hidden > NN_MAXH || d_out > NN_MAXO || d_out < 1 || (d_hidden * d_out) % 4) return false;
    dm.din = d_in; dm.dh = d_hidden; dm.dout = d_out; dm.dop = 4 * ((d_out + 3) / 4);
    dm.offA = 0; dm.offB = d_hidden * d_out; dm.offa = dm.offB + d_in * d_hidden; dm.offb = dm.offa + d_out; dm.offl = dm.offb + d_hidden;
    dm.P = dm.offl + 4;
    return true;
}

// host-only description of the launches of this family (sgm_debug_plan): no device is touched
int softmax_net_plan(int d_in, int d_hidden, int d_out, SoftmaxNetPlan* out) {
    NNDims dm;
    if (!nn_dims(dm, d_in, d_hidden, d_out)) return -1;
    out->threads = NN_THREADS;
    out->smem_bytes = nn_smem_bytes(dm);
    out->tmem_cols = 512;
    out->Pst = 4 * ((dm.P + 3) / 4);
    out->offB = dm.offB;
    SoftmaxNetWork w; w.dm = dm;
    out->tma_fed = nn_tma_enabled(&w) ? 1 : 0;
    out->s1_box_inner = S1_KC;
    out->s1_ring_bytes = (size_t)nn_s1_ring_bytes(dm);
    out->ring_capacity_bytes = (size_t)(2 * NN_RB * dm.dh + nn_tile_floats_all(dm)) * 4;
    return 0;
}

int softmax_net_create(SoftmaxNetWork** out, int d_in, int d_hidden, int d_out, int n, int chains, int device) {
    SoftmaxNetWork* w = new SoftmaxNetWork();
    NNDims& dm = w->dm;
    if (!nn_dims(dm, d_in, d_hidden, d_out)) { delete w; return (int)cudaErrorInvalidValue; }
    w->n = n; w->chains = chains; w->device = device;
    w->smem = nn_smem_bytes(dm);
    const int Pst = 4 * ((dm.P + 3) / 4);
    cudaError_t e = cudaMalloc(&w->G, (size_t)chains * Pst * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&w->Ghat, (size_t)chains * Pst * sizeof(float));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(softmax_net_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(softmax_net_fullgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(softmax_net_logpost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->smem);
    if (e != cudaSuccess) { softmax_net_destroy(w); return (int)e; }
    *out = w;
    return 0;
}

int softmax_net_set_option(SoftmaxNetWork* w, const char* key, int value) {
    if (!w || !key || value < 0 || value > 1) return -1;
    if (!strcmp(key, "nn_update")) { w->opt_update = value; return 0; }
    if (!strcmp(key, "nn_stage1")) { w->opt_stage1 = value; return 0; }
#ifdef NN_ENABLE_S5_TMEM
    if (!strcmp(key, "nn_stage5")) { w->opt_stage5 = value; return 0; }
#endif
    return -1;
}

void softmax_net_destroy(SoftmaxNetWork* w) {
    if (!w) return;
    if (w->G) cudaFree(w->G);
    if (w->Ghat) cudaFree(w->Ghat);
    if (w->part) cudaFree(w->part);
    delete w;
}

void softmax_net_tables(const SoftmaxNetWork* w, int* flat_of, int* tensor_of, int Pst) {
    const NNDims& dm = w->dm;
    for (int e = 0; e < Pst; ++e) {
        if (e >= dm.P) { flat_of[e] = -1; tensor_of[e] = 0; continue; }
        flat_of[e] = e;
        tensor_of[e] = e < dm.offB ? 0 : (e < dm.offa ? 1 : (e < dm.offb ? 2 : (e < dm.offl ? 3 : 4 + (e - dm.offl))));
    }
}

int softmax_net_launch(SoftmaxNetWork* w, const ChainArgs& a, cudaStream_t stream, int* launches) {
    cudaMemcpyToSymbolAsync(nn_prof, &a.prof, sizeof(a.prof), 0, cudaMemcpyHostToDevice, stream);
    {
        const int on = w->opt_update == 1 ? 0 : 1;
        cudaMemcpyToSymbolAsync(nn_stream_updates, &on, sizeof(on), 0, cudaMemcpyHostToDevice, stream);
        const int s5on = w->opt_stage5 == 1 ? 1 : 0;   // opt-in until it has run on a GPU
        cudaMemcpyToSymbolAsync(nn_s5_tmem, &s5on, sizeof(s5on), 0, cudaMemcpyHostToDevice, stream);
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
    const int grid = a.C < sms ? a.C : sms;
    static const CUtensorMap zero_map{};
    const CUtensorMap* mx = nn_tma_enabled(w) ? nn_x_map(w, a.X, a.N, a.ld) : nullptr;
    const CUtensorMap* mt = mx ? nn_b_map(w, a.theta, a.C, a.Pst) : nullptr;
    const CUtensorMap* mh = (mx && a.cv) ? nn_b_map(w, a.theta_hat, a.hat_stride ? a.C : 1, a.hat_stride) : mt;
    const int use_tma = (mx && mt && mh) ? 1 : 0;
    SGM_LAUNCH(softmax_net_chain_kernel, grid, NN_THREADS, w->smem, stream, a, w->dm, w->G, a.cv ? w->Ghat : nullptr, use_tma ? *mx : zero_map,
               use_tma ? *mt : zero_map, use_tma ? *mh : zero_map, use_tma);
    *launches = 1;
    return (int)cudaGetLastError();
}

int softmax_net_fullgrad(SoftmaxNetWork* w, const ChainArgs& a, int64_t row0, int64_t row1, int M, const float* theta_hat,
                         double* gsum, cudaStream_t stream) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
    const int64_t rows = row1 - row0;
    int gx = (int)((sms + M - 1) / M);
    if (gx < 1) gx = 1;
    int64_t rows_per = (rows + gx - 1) / gx;
    if (rows_per < NN_RB) rows_per = NN_RB;
    rows_per = ((rows_per + NN_RB - 1) / NN_RB) * NN_RB;
    gx = (int)((rows + rows_per - 1) / rows_per);
    if (gx < 1) gx = 1;
    const size_t need = (size_t)M * gx * a.Pst * sizeof(float);
    if (need > w->part_cap) {
        if (w->part) cudaFree(w->part);
        w->part = nullptr; w->part_cap = 0;
        cudaError_t e = cudaMalloc(&w->part, need);
        if (e != cudaSuccess) return (int)e;
        w->part_cap = need;
    }
    dim3 grid((unsigned)gx, (unsigned)M);
    static const CUtensorMap zero_map{};
    const CUtensorMap* mx = nn_tma_enabled(w) ? nn_x_map(w, a.X, a.N, a.ld) : nullptr;
    const CUtensorMap* mh = mx ? nn_b_map(w, theta_hat, M, a.Pst) : nullptr;
    const int use_tma = (mx && mh) ? 1 : 0;
    SGM_LAUNCH(softmax_net_fullgrad_kernel, grid, NN_THREADS, w->smem, stream, a, w->dm, row0, row1, (int)rows_per, theta_hat, w->part,
               use_tma ? *mx : zero_map, use_tma ? *mh : zero_map, use_tma);
    const int total = M * a.Pst;
    SGM_LAUNCH(softmax_net_part_reduce, (total + 255) / 256, 256, 0, stream, w->part, gx, a.Pst, gsum, total);
    return (int)cudaGetLastError();
}

int softmax_net_fullgrad_finish(SoftmaxNetWork* w, const ChainArgs& a, int M, const double* gsum, const float* theta_hat,
                                float* g_hat, float* theta, cudaStream_t stream) {
    SGM_LAUNCH(softmax_net_finish_kernel, M, NN_THREADS, 0, stream, a, w->dm, gsum, theta_hat, g_hat, theta);
    return (int)cudaGetLastError();
}

int softmax_net_logpost(SoftmaxNetWork* w, const ChainArgs& a, uint32_t word2, const float* theta, double* out, cudaStream_t stream) {
    static const CUtensorMap zero_map{};
    const CUtensorMap* mx = nn_tma_enabled(w) ? nn_x_map(w, a.X, a.N, a.ld) : nullptr;
    const CUtensorMap* mt = mx ? nn_b_map(w, theta, a.C, a.Pst) : nullptr;
    const int use_tma = (mx && mt) ? 1 : 0;
    SGM_LAUNCH(softmax_net_logpost_kernel, a.C, NN_THREADS, w->smem, stream, a, w->dm, word2, theta, w->G, out, use_tma ? *mx : zero_map,
               use_tma ? *mt : zero_map, use_tma);
    return (int)cudaGetLastError();
}

}  // namespace sgm
