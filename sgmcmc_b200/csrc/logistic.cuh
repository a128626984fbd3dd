// Logistic regression family: one CTA per chain, G lanes per gathered row.
//
// One launch runs `iters` sgmcmcSteps for every chain: minibatch index sampling
// (keyed permutation = without replacement), row gather with 16-byte vector loads,
// per-datum gradient, shuffle + shared-memory reduction, x N/n, and the fused
// SGLD / SGHMC / SGNHT (or mode-finder SGD) update with Philox noise.
//
// Replaces dataFeed/dataSelect (R/update.R:166-194), tf$gradients of estLogPost for
// the closure of vignettes/logisticRegression.Rmd:45-57 (R/setup.R:36-62),
// declareDynamics.* (R/dynamics.R:7-120), sgmcmcStep.* (R/update.R:40-81) and
// optUpdate (R/controlVariates.R:30-33).
//
// Row layout (float32): d/4 full float4 of x, then one tail float4 holding the last
// d%4 features, y, and zero padding; ld = 4*(d/4 + 1).  Internal parameter layout:
// beta[0..d-1], bias at index d (the slot y occupies in the row, which the kernel
// overwrites with 1.0 after extracting y), zero padding.  Flat (ABI) order is the
// vignette's names(params) order: bias, beta.
#pragma once
#include "common.cuh"

namespace sgm {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float get4(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
__device__ __forceinline__ void set4(float4& v, int k, float s) {
    v.x = (k == 0) ? s : v.x; v.y = (k == 1) ? s : v.y; v.z = (k == 2) ? s : v.z; v.w = (k == 3) ? s : v.w;
}
__device__ __forceinline__ float signf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float block_sum(float v, float* scratch /* [32] */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? scratch[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) { t = warp_sum(t); if (threadIdx.x == 0) scratch[0] = t; }
    __syncthreads();
    return scratch[0];
}

enum LogiPass : int { PASS_FULL = 0, PASS_BIAS = 1, PASS_BETA = 2 };

struct LogiSmem {
    float* th;    // [ldp] current state (beta, bias)
    float* hat;   // [ldp] mode
    float* ghat;  // [ldp] full-data gradient at the mode
    float* mo;    // [ldp] momentum
    float* grad;  // [ldp] reduced minibatch sum
    float* red;   // [nwarps][ldp]
    float* dots;  // [n] x_i . beta   (two-pass Gauss-Seidel only)
    float* rhat;  // [n] y_i - sigmoid(zhat_i)
    float* ys;    // [n]
    float* scratch;  // [32]
};

// One pass over the minibatch.  PASS_FULL: sum_i c_i * [x_i, 1];  PASS_BIAS: sum_i c_i only
// (stores the dots);  PASS_BETA: sum_i c_i' x_i with c_i' from the stored dots and the
// already-updated bias.  c_i = y_i - s(z_i), minus y_i - s(zhat_i) for control variates.
template <int G, int V>
__device__ __forceinline__ void logi_pass(const ChainArgs& a, const LogiSmem& sm, int pass, const Perm& perm,
                                          const int32_t* __restrict__ idx) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NG = blockDim.x / G;
    const int gid = tid / G, lg = tid % G;
    const int ldp = V * G * 4;
    const int qt = a.d >> 2, ct = a.d & 3;      // tail float4 index / slot of y in it
    const int lg_t = qt % G, v_t = qt / G;
    const float bias = sm.th[a.d], bias_hat = a.cv ? sm.hat[a.d] : 0.f;

    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float csum = 0.f;

    for (int j0 = 0; j0 < a.n; j0 += NG) {
        const int j = j0 + gid;
        const bool valid = j < a.n;
        float4 x[V];
        if (valid) {
            const uint32_t row = idx ? (uint32_t)idx[j] : perm_at(perm, (uint32_t)j);
            const float* __restrict__ xr = row_ptr(a, row);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int col = 4 * (lg + G * v);
                x[v] = (col < a.ld) ? ld4(xr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // extract y from the tail float4 and blank its slot
        float y = 0.f, rc = 0.f;
        const bool hat_dot = a.cv && !a.rhat_cached;   // the residual at a shared mode rides in the row (column d + 1)
        if (lg == lg_t) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v == v_t) { y = get4(x[v], ct); rc = get4(x[v], ct + 1); set4(x[v], ct, 0.f); }
        }
        float coef;
        if (pass == PASS_BETA) {
            const float yy = valid ? sm.ys[j] : 0.f;
            const float r = yy - sigmoidf(bias + (valid ? sm.dots[j] : 0.f));
            coef = a.cv ? r - (valid ? sm.rhat[j] : 0.f) : r;
        } else {
            y = __shfl_sync(0xffffffffu, y, lg_t, G);
            float dz = 0.f, dh = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int col = 4 * (lg + G * v);
                dz += dot4(x[v], *reinterpret_cast<const float4*>(sm.th + col));
                if (hat_dot) dh += dot4(x[v], *reinterpret_cast<const float4*>(sm.hat + col));
            }
            // the bias slot of th/hat meets a blanked x slot, so dz/dh are pure x.beta
            dz = group_sum<G>(dz);
            const float r = y - sigmoidf(bias + dz);
            float rh = 0.f;
            if (hat_dot) { dh = group_sum<G>(dh); rh = y - sigmoidf(bias_hat + dh); }
            else if (a.cv) rh = __shfl_sync(0xffffffffu, rc, lg_t, G);
            coef = a.cv ? r - rh : r;
            if (pass == PASS_BIAS && valid && lg == 0) { sm.dots[j] = dz; sm.rhat[j] = rh; sm.ys[j] = y; }
        }
        if (!valid) coef = 0.f;
        if (pass == PASS_BIAS) {
            if (lg == 0) csum += coef;
        } else {
            if (lg == lg_t) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (v == v_t) set4(x[v], ct, 1.0f);   // the bias "feature"
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                acc[v].x += coef * x[v].x; acc[v].y += coef * x[v].y;
                acc[v].z += coef * x[v].z; acc[v].w += coef * x[v].w;
            }
        }
    }

    if (pass == PASS_BIAS) {
        const float tot = block_sum(csum, sm.scratch);
        if (tid == 0) sm.grad[a.d] = tot;
        __syncthreads();
        return;
    }
    // groups of one warp -> lanes < G, then warps -> sm.grad
#pragma unroll
    for (int v = 0; v < V; ++v) {
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
            acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
            acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
            acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
            acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
        }
        if (lane < G) *reinterpret_cast<float4*>(sm.red + (size_t)warp * ldp + 4 * (lane + G * v)) = acc[v];
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int e = tid; e < ldp; e += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += sm.red[(size_t)w * ldp + e];
        sm.grad[e] = s;
    }
    __syncthreads();
}

template <int G, int V>
__global__ void __launch_bounds__(512) logistic_chain_kernel(const ChainArgs a) {
    SGM_DYN_SMEM(float, smem_f, 16);
    const int tid = threadIdx.x;
    const int ldp = V * G * 4;
    const int nw = blockDim.x >> 5;
    const int d = a.d;
    const float SQRT_HALF = 0.70710678118654752440f;
    LogiSmem sm;
    sm.th = smem_f; sm.hat = sm.th + ldp; sm.ghat = sm.hat + ldp; sm.mo = sm.ghat + ldp;
    sm.grad = sm.mo + ldp; sm.red = sm.grad + ldp; sm.scratch = sm.red + (size_t)nw * ldp;
    sm.dots = sm.scratch + 32; sm.rhat = sm.dots + a.n; sm.ys = sm.rhat + a.n;
    const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
    SGM_EMU_NOTE(a.n_shards > 1 ? "logistic_chain:sharded" : "logistic_chain");

    for (int c = blockIdx.x; c < a.C; c += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < ldp; e += blockDim.x) {
            const bool live = e <= d;
            sm.th[e] = live ? a.theta[(size_t)c * a.Pst + e] : 0.f;
            sm.mo[e] = live ? a.mom[(size_t)c * a.Pst + e] : 0.f;
            sm.hat[e] = (live && a.cv) ? a.theta_hat[(size_t)c * a.hat_stride + e] : 0.f;
            sm.ghat[e] = (live && a.cv) ? a.g_hat[(size_t)c * a.hat_stride + e] : 0.f;
        }
        float zeta_bias = a.thermo[c * MAX_TENSORS + 0], zeta_beta = a.thermo[c * MAX_TENSORS + 1];
        bool bad = false;
        __syncthreads();

        for (int it = 0; it < a.iters; ++it) {
            const uint64_t step = a.step0 + (uint64_t)it;
            // internal element e: tensor 0 = bias (e == d), tensor 1 = beta; flat index f
            auto noise_at = [&](int f, int sub) -> float {
                if (a.det) return a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + f];
                return normal_at((uint32_t)f, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1);
            };
            if (a.algo == ALGO_SGHMC) {  // refresh, R/dynamics.R:62
                for (int e = tid; e <= d; e += blockDim.x) {
                    const int t = (e == d) ? 0 : 1, f = (e == d) ? 0 : e + 1;
                    sm.mo[e] = 0.5f * (a.tc.sqrt_2eps[t] * noise_at(f, 0));
                }
                __syncthreads();
            }
            for (int l = 0; l < a.L; ++l) {
                const int sub = l + 1;
                const Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), a.index_tag, a.k0, a.k1);
                const int32_t* idx = a.det ? a.idx + (((size_t)c * a.iters + it) * a.L + l) * a.n : nullptr;
                const int passes = seidel ? 2 : 1;
                for (int pass = 0; pass < passes; ++pass) {
                    const int mode = seidel ? (pass == 0 ? PASS_BIAS : PASS_BETA) : PASS_FULL;
                    logi_pass<G, V>(a, sm, mode, perm, idx);
                    // fused update of the elements this pass owns
                    float uu = 0.f;
                    for (int e = tid; e <= d; e += blockDim.x) {
                        const int t = (e == d) ? 0 : 1, f = (e == d) ? 0 : e + 1;
                        if (seidel && t != pass) continue;
                        const float th = sm.th[e];
                        float g = a.corr * sm.grad[e];
                        if (a.use_prior) g += -signf(th);
                        if (a.cv) {
                            if (a.use_prior) g += signf(sm.hat[e]);
                            g += sm.ghat[e];
                        }
                        if (a.algo == ALGO_SGLD) {
                            sm.th[e] = th + (a.tc.half_eps[t] * g + SQRT_HALF * (a.tc.sqrt_2eps[t] * noise_at(f, sub)));
                        } else if (a.algo == ALGO_SGHMC) {
                            float nu = sm.mo[e];
                            nu += a.tc.eps[t] * g - a.tc.alpha[t] * nu + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * noise_at(f, sub));
                            sm.mo[e] = nu;
                            sm.th[e] = th + nu;
                        } else if (a.algo == ALGO_SGNHT) {
                            float u = sm.mo[e];
                            u += a.tc.eps[t] * g - u * (t == 0 ? zeta_bias : zeta_beta) + a.tc.sqrt_a[t] * (a.tc.sqrt_2eps[t] * noise_at(f, sub));
                            sm.mo[e] = u;
                            sm.th[e] = th + u;
                            if (t == 1) uu += u * u;
                        } else if (a.algo == ALGO_OPT) {
                            sm.th[e] = th + a.opt_lr * g;
                        } else {  // ALGO_GRAD
                            sm.mo[e] = g;
                        }
                    }
                    if (a.algo == ALGO_SGNHT) {  // thermostats, R/dynamics.R:108-116
                        const float tot = block_sum(uu, sm.scratch);
                        const float ub = sm.mo[d];
                        zeta_beta += tot / (float)d - a.tc.eps[1];
                        zeta_bias += ub * ub - a.tc.eps[0];
                    }
                    __syncthreads();
                }
            }
            for (int e = tid; e <= d; e += blockDim.x) {
                const float v = sm.th[e];
                bad |= !isfinite(v);
                if (a.store) a.store[((size_t)(a.store_pos + it) * a.C + c) * a.P + ((e == d) ? 0 : e + 1)] = v;
            }
        }
        for (int e = tid; e <= d; e += blockDim.x) {
            a.theta[(size_t)c * a.Pst + e] = sm.th[e];
            a.mom[(size_t)c * a.Pst + e] = sm.mo[e];
        }
        if (tid == 0) { a.thermo[c * MAX_TENSORS + 0] = zeta_bias; a.thermo[c * MAX_TENSORS + 1] = zeta_beta; }
        if (bad) *a.diverged = 1;
    }
}

// Control-variate full-data gradient at the mode(s): sum over rows [row0, row1) of
// (y_i - s(z_i)) [x_i, 1] for every chain's theta_hat (calcFullGrads,
// R/controlVariates.R:105-111; fullCalc, cV.R:96-97).  grid = (row chunks, modes);
// partial sums are combined with double-precision atomics into gsum[mode][Pst].
template <int G, int V>
__global__ void __launch_bounds__(512) logistic_fullgrad_kernel(const float* __restrict__ X, int ld, int d,
                                                                int64_t row0, int64_t row1, int rows_per_cta,
                                                                const float* __restrict__ theta_hat, int Pst,
                                                                double* __restrict__ gsum) {
    SGM_DYN_SMEM(float, smem_f, 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldp = V * G * 4;
    const int nw = blockDim.x >> 5;
    const int NG = blockDim.x / G, gid = tid / G, lg = tid % G;
    float* th = smem_f;
    float* red = th + ldp;
    const int m = blockIdx.y;
    for (int e = tid; e < ldp; e += blockDim.x) th[e] = (e <= d) ? theta_hat[(size_t)m * Pst + e] : 0.f;
    __syncthreads();
    const int qt = d >> 2, ct = d & 3, lg_t = qt % G, v_t = qt / G;
    const float bias = th[d];
    const int64_t r_begin = row0 + (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_end = (r_begin + rows_per_cta < row1) ? r_begin + rows_per_cta : row1;

    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t j0 = r_begin; j0 < r_end; j0 += NG) {
        const int64_t j = j0 + gid;
        const bool valid = j < r_end;
        float4 x[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int col = 4 * (lg + G * v);
            x[v] = (valid && col < ld) ? ld4(X + (size_t)j * ld + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float y = 0.f;
        if (lg == lg_t) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v == v_t) { y = get4(x[v], ct); set4(x[v], ct, 0.f); }
        }
        y = __shfl_sync(0xffffffffu, y, lg_t, G);
        float dz = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) dz += dot4(x[v], *reinterpret_cast<const float4*>(th + 4 * (lg + G * v)));
        dz = group_sum<G>(dz);
        const float coef = valid ? y - sigmoidf(bias + dz) : 0.f;
        if (lg == lg_t) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v == v_t) set4(x[v], ct, 1.0f);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            acc[v].x += coef * x[v].x; acc[v].y += coef * x[v].y;
            acc[v].z += coef * x[v].z; acc[v].w += coef * x[v].w;
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
            acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
            acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
            acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
            acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
        }
        if (lane < G) *reinterpret_cast<float4*>(red + (size_t)warp * ldp + 4 * (lane + G * v)) = acc[v];
    }
    __syncthreads();
    for (int e = tid; e <= d; e += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += red[(size_t)w * ldp + e];
        atomicAdd(gsum + (size_t)m * Pst + e, (double)s);
    }
}

// g_hat = prior gradient at the mode + reduced sum; also theta := theta_hat (cV.R:98-99,109).
__global__ void logistic_fullgrad_finish(const double* __restrict__ gsum, const float* __restrict__ theta_hat,
                                         float* __restrict__ g_hat, float* __restrict__ theta, int Pst, int d,
                                         int use_prior, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = i % Pst;
    if (e > d) { g_hat[i] = 0.f; return; }
    const float th = theta_hat[i];
    g_hat[i] = (float)gsum[i] + (use_prior ? -signf(th) : 0.f);
    theta[i] = th;
}

// estLogPost value on a minibatch, one CTA per chain (checkDivergence, R/update.R:197-207).
__global__ void logistic_logpost_kernel(const ChainArgs a, uint32_t word2, const float* __restrict__ theta,
                                        double* __restrict__ out) {
    __shared__ float scratch[32];
    const int c = blockIdx.x, tid = threadIdx.x;
    const float* th = theta + (size_t)c * a.Pst;
    const Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2, TAG_CHECK_INDEX, a.k0, a.k1);
    const int32_t* idx = (a.det == 1) ? a.idx + (size_t)c * a.n : nullptr;
    float ll = 0.f;
    for (int j = tid; j < a.n; j += blockDim.x) {
        const uint32_t row = (a.det == 2) ? (uint32_t)j : (idx ? (uint32_t)idx[j] : perm_at(perm, (uint32_t)j));   // det == 2: rows 0..n-1
        const float* xr = (a.det == 2) ? a.X + (size_t)row * a.ld : row_ptr(a, row);   // det == 2: a held-out set of this rank
        float z = 0.f;
        for (int k = 0; k < a.d; ++k) z += xr[k] * th[k];
        z += th[a.d];
        const float y = xr[a.d];
        // y log s + (1-y) log(1-s) in the overflow-free form  y z - log(1 + e^z)
        ll += y * z - (fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))));
    }
    ll = block_sum(ll, scratch);
    float lp = 0.f;
    if (a.use_prior) {
        for (int e = tid; e <= a.d; e += blockDim.x) lp += -fabsf(th[e]);
    }
    lp = block_sum(lp, scratch);
    if (tid == 0) out[c] = (double)(lp + a.corr * ll);
}

}  // namespace sgm
