// Logistic regression, short rows (ld <= 128 floats): the whole minibatch is staged in
// shared memory asynchronously, one CTA per chain, lane-per-row arithmetic, 2 CTAs per SM.
//
//   gather   16-byte cp.async (LDGSTS) straight into shared memory; lane pairs cover one 32-byte
//            sector, so every fetched sector is used once.  (A first version issued one TMA bulk
//            copy per row: cp.async.bulk is a per-warp instruction, so 32 row copies per warp were
//            serialised into a 9-instruction x 32 loop = 18% of all issued instructions for 224-byte
//            rows -- profiles/r01_c3_staged_tma_per_row.txt.  TMA bulk copies are used where a row
//            is long enough to amortise the issue: logistic_ring.cuh.)
//   phase 1  lane-per-row dot products x.beta; each lane walks its R = ceil(n/256) rows together so
//            that one broadcast load of beta serves R rows.
//   phase 2  column mapping: a warp sums c_i * x_i over all rows for one (or two) float4 columns,
//            lanes stride the rows, one shuffle tree per warp.  No per-thread accumulator arrays.
//   update   fused SGLD / SGHMC / SGNHT / mode-finder step with Philox noise; thread e owns element e
//            (momentum, mode, full gradient stay in its registers for the whole launch).
// The reference's per-tensor Gauss-Seidel order (bias, then beta; R/update.R:43-45,60-64) costs only a
// second sigmoid per row here: the dots stay in registers, the rows stay in shared memory.
//
// Control variates with ONE shared mode: theta_hat never changes after initSess, so the per-datum
// residual y_i - s(x_i.beta_hat + bias_hat) is computed once for all N rows (logistic_cache_rhat_kernel)
// and kept in the spare slot next to y in the row's tail float4; it rides along with the gather for
// free and removes the second dot product from every step.
//
// Shared-memory row stride S (in float4): rows are stored UNPADDED when ld/4 = 2 (mod 4) -- e.g. the
// covertype shape ld = 56 -- so that a 500-row minibatch (112 KB) fits twice per SM; bank conflicts
// are removed by skewing the lanes instead (phase 1: odd row-quads start one float4 later; phase 2:
// even/odd lanes take adjacent columns).  Otherwise S = ld/4 | 1 (odd stride, conflict free as is).
//
// Replaces the same reference code as logistic.cuh (dataFeed/dataSelect R/update.R:166-194,
// tf$gradients R/setup.R:36-62, declareDynamics.* R/dynamics.R:7-120, sgmcmcStep.* R/update.R:40-81).
#pragma once
#include "common.cuh"
#include "logistic.cuh"

namespace sgm {

constexpr int LS_THREADS = 256;
constexpr int LS_WARPS = LS_THREADS / 32;
constexpr int LS_MAX_R = 4;   // rows per lane (n <= 1024)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// sigmoid with a full-precision exponential and a single-instruction reciprocal
__device__ __forceinline__ float sigmoid_fast(float z) { return __fdividef(1.0f, 1.0f + expf(-z)); }
__device__ __forceinline__ float fdot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}

__host__ __device__ inline int ls_row_stride4(int ld) { const int Q = ld >> 2; return (Q % 4 == 2) ? Q : (Q | 1); }
static inline size_t ls_smem_bytes(int ld, int n) {
    return sizeof(float) * ((size_t)n * 4 * ls_row_stride4(ld) + 3 * (size_t)ld + (size_t)n + 64);
}

// scratch slots
constexpr int LS_BIAS = 40, LS_BIAS_HAT = 41, LS_UBIAS = 42, LS_PERM = 48;   // [48..54] = Perm words

struct LogiStagedSmem {
    float* rows;     // [n][4*S]
    float* th;       // [ld] beta, with the bias slot (index d) and the padding held at ZERO: full-row dots need no masking
    float* hat;      // [ld] same for the mode
    float* grad;     // [ld]
    float* coef;     // [n]
    float* scratch;  // [64]
};

__device__ __forceinline__ float ls_block_sum(float v, float* scratch) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LS_WARPS; ++w) t += scratch[w];
    return t;   // callers must __syncthreads() before the next ls_block_sum
}

// r_hat_i = y_i - s(bias_hat + x_i . beta_hat) for every row, stored at column d + 1 (spare slot of the tail float4)
__global__ void logistic_cache_rhat_kernel(float* __restrict__ X, int64_t N, int ld, int d, const float* __restrict__ theta_hat) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    const float* xr = X + row * ld;
    float z = 0.f;
    for (int k = lane; k < d; k += 32) z = fmaf(xr[k], theta_hat[k], z);
    z = warp_sum(z);
    if (lane == 0) X[row * ld + d + 1] = xr[d] - sigmoidf(theta_hat[d] + z);
}

// HAT_DOT: control variates whose residual at the mode is NOT cached in the rows (per-chain modes)
template <bool HAT_DOT>
__global__ void __launch_bounds__(LS_THREADS, 2) logistic_staged_kernel(const ChainArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = a.d, ld = a.ld, n = a.n;
    const int Q = ld >> 2, S = ls_row_stride4(ld);
    const bool paired = (S & 1) == 0;      // unpadded rows: skewed lanes
    const int qt = d >> 2, ct = d & 3;     // float4 / slot holding y (and, for the parameters, the bias)
    const int R = (n + LS_THREADS - 1) / LS_THREADS;
    const float SQRT_HALF = 0.70710678118654752440f;
    LogiStagedSmem sm;
    sm.rows = reinterpret_cast<float*>(smem_raw);
    sm.th = sm.rows + (size_t)n * 4 * S; sm.hat = sm.th + ld; sm.grad = sm.hat + ld; sm.coef = sm.grad + ld; sm.scratch = sm.coef + n;
    const float4* rows4 = reinterpret_cast<const float4*>(sm.rows);
    const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
    const bool cached = a.cv && !HAT_DOT;   // r_hat rides in the row
    const int e = tid;                      // the parameter element this thread owns in the update (e <= d)
    const bool owner = e <= d;
    const int t_own = (e == d) ? 0 : 1, f_own = (e == d) ? 0 : e + 1;

    for (int c = blockIdx.x; c < a.C; c += gridDim.x) {
        __syncthreads();
        float th_e = 0.f, mo_e = 0.f, hat_e = 0.f, ghat_e = 0.f;
        if (owner) {
            th_e = a.theta[(size_t)c * a.Pst + e]; mo_e = a.mom[(size_t)c * a.Pst + e];
            if (a.cv) { hat_e = a.theta_hat[(size_t)c * a.hat_stride + e]; ghat_e = a.g_hat[(size_t)c * a.hat_stride + e]; }
        }
        if (tid < ld) {
            sm.th[tid] = (tid < d) ? th_e : 0.f;
            sm.hat[tid] = (tid < d) ? hat_e : 0.f;
        }
        if (e == d) { sm.scratch[LS_BIAS] = th_e; sm.scratch[LS_BIAS_HAT] = hat_e; }
        float zeta_bias = a.thermo[c * MAX_TENSORS + 0], zeta_beta = a.thermo[c * MAX_TENSORS + 1];
        bool bad = false;

        for (int it = 0; it < a.iters; ++it) {
            const uint64_t step = a.step0 + (uint64_t)it;
            auto noise_own = [&](int sub) -> float {
                if (a.det) return a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + f_own];
                return normal_at((uint32_t)f_own, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1);
            };
            if (a.algo == ALGO_SGHMC && owner) mo_e = 0.5f * (a.tc.sqrt_2eps[t_own] * noise_own(0));   // refresh, R/dynamics.R:62
            for (int l = 0; l < a.L; ++l) {
                const int sub = l + 1;
                const int32_t* idx = a.det ? a.idx + (((size_t)c * a.iters + it) * a.L + l) * n : nullptr;
                // the permutation keys cost one Philox call: warp 0 computes, everybody reads
                if (!a.det && warp == 0) {
                    const Perm p = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), a.index_tag, a.k0, a.k1);
                    if (lane < 4) reinterpret_cast<uint32_t*>(sm.scratch + LS_PERM)[lane] = p.key[lane];
                }
                __syncthreads();   // previous minibatch fully consumed; th/hat/bias/keys published
                Perm perm;
                perm.N = a.N; perm.mask = (a.perm_bits >= 32) ? 0xFFFFFFFFu : ((1u << a.perm_bits) - 1u);
                perm.shift = a.perm_bits / 2 > 0 ? a.perm_bits / 2 : 1;
#pragma unroll
                for (int k = 0; k < 4; ++k) perm.key[k] = reinterpret_cast<const uint32_t*>(sm.scratch + LS_PERM)[k];
                // ---- gather: each warp fetches the rows its lanes own (R x 32), everything in flight at once
                const int sub16 = lane & 1, rl = lane >> 1;
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
                    const int j = r * LS_THREADS + tid;
                    const int j0 = r * LS_THREADS + warp * 32;
                    uint32_t my_row = 0;
                    if (j < n) my_row = idx ? (uint32_t)idx[j] : perm_at(perm, (uint32_t)j);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int jr = h * 16 + rl;
                        const uint32_t row = __shfl_sync(0xffffffffu, my_row, jr);
                        if (j0 + jr < n) {
                            const float* src = a.X + ((size_t)row * ld + 4 * sub16);
                            uint32_t dst = smem_u32(sm.rows) + (uint32_t)((j0 + jr) * S + sub16) * 16u;
#pragma unroll 2
                            for (int q = sub16; q < Q; q += 2) {
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                                src += 8; dst += 32u;
                            }
                        }
                    }
                }
                cp_async_commit();
                cp_async_wait_all();
                __syncwarp();
                // ---- phase 1: lane-per-row dots, R rows per lane walked together (one beta load serves R rows)
                const float4* t4 = reinterpret_cast<const float4*>(sm.th);
                const float4* h4 = reinterpret_cast<const float4*>(sm.hat);
                // skew (unpadded rows only): row quads 4..7 (mod 8) start one float4 later, so the 8 lanes of a
                // shared-memory phase hit 8 different 16-byte bank groups although the row stride is even
                const int skew = paired ? ((tid >> 2) & 1) : 0;
                float dz[LS_MAX_R], dh[LS_MAX_R];
                const float4* xr[LS_MAX_R];
#pragma unroll
                for (int r = 0; r < LS_MAX_R; ++r) {
                    dz[r] = 0.f; dh[r] = 0.f;
                    const int j = r * LS_THREADS + tid;
                    xr[r] = rows4 + (size_t)((r < R && j < n) ? j : 0) * S;   // out-of-range lanes recompute row 0 (discarded)
                }
                for (int k = 0; k < Q; ++k) {
                    int q = k + skew; q = (q >= Q) ? q - Q : q;
                    const float4 tv = t4[q];
                    float4 hv;
                    if (HAT_DOT) hv = h4[q];
#pragma unroll
                    for (int r = 0; r < LS_MAX_R; ++r) {
                        if (r < R) {
                            const float4 x = xr[r][q];
                            dz[r] = fdot4(x, tv, dz[r]);
                            if (HAT_DOT) dh[r] = fdot4(x, hv, dh[r]);
                        }
                    }
                }
                float yv[LS_MAX_R], rhat[LS_MAX_R];
                const float bias_hat = sm.scratch[LS_BIAS_HAT];
#pragma unroll
                for (int r = 0; r < LS_MAX_R; ++r) {
                    yv[r] = 0.f; rhat[r] = 0.f;
                    if (r < R) {
                        const float4 tail = xr[r][qt];
                        yv[r] = get4(tail, ct);
                        if (cached) rhat[r] = get4(tail, ct + 1);
                        else if (HAT_DOT) rhat[r] = yv[r] - sigmoid_fast(bias_hat + dh[r]);
                        if (cached) {
                            // the cached residual sits in a slot where beta is zero-padded, so it never enters dz
                        }
                    }
                }
                const int passes = seidel ? 2 : 1;
                for (int pass = 0; pass < passes; ++pass) {
                    const float bias = sm.scratch[LS_BIAS];
                    float csum = 0.f;
#pragma unroll
                    for (int r = 0; r < LS_MAX_R; ++r) {
                        const int j = r * LS_THREADS + tid;
                        if (r < R && j < n) {
                            const float rr = yv[r] - sigmoid_fast(bias + dz[r]);
                            const float cf = a.cv ? rr - rhat[r] : rr;
                            csum += cf;
                            sm.coef[j] = cf;
                        }
                    }
                    const float cs_tot = ls_block_sum(csum, sm.scratch);   // = gradient sum of the bias; also publishes coef
                    if (!(seidel && pass == 0)) {
                        // ---- phase 2: column sums of c_i * x_i
                        if (paired) {
                            const int cpar = lane & 1;
                            for (int q0 = 2 * warp; q0 < Q; q0 += 2 * LS_WARPS) {
                                const float4* col = rows4 + q0 + cpar;
                                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                                for (int j = rl; j < n; j += 16) {
                                    const float4 x = col[(size_t)j * S];
                                    const float cf = sm.coef[j];
                                    acc.x = fmaf(cf, x.x, acc.x); acc.y = fmaf(cf, x.y, acc.y); acc.z = fmaf(cf, x.z, acc.z); acc.w = fmaf(cf, x.w, acc.w);
                                }
#pragma unroll
                                for (int o = 2; o < 32; o <<= 1) {
                                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                                }
                                if (lane < 2) *reinterpret_cast<float4*>(sm.grad + 4 * (q0 + cpar)) = acc;
                            }
                        } else {
                            for (int q = warp; q < Q; q += LS_WARPS) {
                                const float4* col = rows4 + q;
                                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                                for (int j = lane; j < n; j += 32) {
                                    const float4 x = col[(size_t)j * S];
                                    const float cf = sm.coef[j];
                                    acc.x = fmaf(cf, x.x, acc.x); acc.y = fmaf(cf, x.y, acc.y); acc.z = fmaf(cf, x.z, acc.z); acc.w = fmaf(cf, x.w, acc.w);
                                }
                                acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
                                if (lane == 0) *reinterpret_cast<float4*>(sm.grad + 4 * q) = acc;
                            }
                        }
                        __syncthreads();
                    }
                    // ---- fused update: thread e owns element e; the y slot of grad is garbage, the bias uses cs_tot
                    float uu = 0.f;
                    if (owner && !(seidel && t_own != pass)) {
                        float g = a.corr * ((e == d) ? cs_tot : sm.grad[e]);
                        if (a.use_prior) g += -signf(th_e);
                        if (a.cv) {
                            if (a.use_prior) g += signf(hat_e);
                            g += ghat_e;
                        }
                        const int t = t_own;
                        if (a.algo == ALGO_SGLD) {
                            th_e = th_e + (a.tc.half_eps[t] * g + SQRT_HALF * (a.tc.sqrt_2eps[t] * noise_own(sub)));
                        } else if (a.algo == ALGO_SGHMC) {
                            mo_e += a.tc.eps[t] * g - a.tc.alpha[t] * mo_e + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * noise_own(sub));
                            th_e = th_e + mo_e;
                        } else if (a.algo == ALGO_SGNHT) {
                            mo_e += a.tc.eps[t] * g - mo_e * (t == 0 ? zeta_bias : zeta_beta) + a.tc.sqrt_a[t] * (a.tc.sqrt_2eps[t] * noise_own(sub));
                            th_e = th_e + mo_e;
                            if (t == 1) uu = mo_e * mo_e;
                        } else if (a.algo == ALGO_OPT) {
                            th_e = th_e + a.opt_lr * g;
                        } else {
                            mo_e = g;
                        }
                        if (e == d) { sm.scratch[LS_BIAS] = th_e; if (a.algo == ALGO_SGNHT) sm.scratch[LS_UBIAS] = mo_e; }
                        else sm.th[e] = th_e;
                    }
                    if (a.algo == ALGO_SGNHT) {  // thermostats, R/dynamics.R:108-116
                        __syncthreads();
                        const float tot = ls_block_sum(uu, sm.scratch);
                        const float ub = sm.scratch[LS_UBIAS];
                        zeta_beta += tot / (float)d - a.tc.eps[1];
                        zeta_bias += ub * ub - a.tc.eps[0];
                    }
                    __syncthreads();
                }
            }
            if (owner) {
                bad |= !isfinite(th_e);
                if (a.store) a.store[((size_t)(a.store_pos + it) * a.C + c) * a.P + f_own] = th_e;
            }
        }
        if (owner) {
            a.theta[(size_t)c * a.Pst + e] = th_e;
            a.mom[(size_t)c * a.Pst + e] = mo_e;
        }
        if (tid == 0) { a.thermo[c * MAX_TENSORS + 0] = zeta_bias; a.thermo[c * MAX_TENSORS + 1] = zeta_beta; }
        if (bad) *a.diverged = 1;
    }
}

}  // namespace sgm
