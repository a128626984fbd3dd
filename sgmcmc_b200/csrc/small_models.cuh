// Gaussian-mean and two-component-mixture families: one warp per chain, state in
// registers, the whole sgmcmcStep loop (sample -> gather -> per-datum gradient ->
// warp-shuffle reduce -> x N/n -> fused SGLD/SGHMC/SGNHT update with Philox noise)
// inside one persistent kernel.
//
// Replaces, for these families, dataFeed/dataSelect (R/update.R:166-194),
// tf$gradients of estLogPost (R/setup.R:36-62,116-125), declareDynamics.*
// (R/dynamics.R:7-120) and the per-tensor Session.run loops of sgmcmcStep.*
// (R/update.R:40-81).  Models: vignettes/mvGauss.Rmd:37-54,
// tests/testthat/test{1d,3d}Gauss.R, vignettes/gaussMixture.Rmd:48-76.
#pragma once
#include "common.cuh"

namespace sgm {

constexpr int SMALL_MODEL_GAUSS = 1;
constexpr int SMALL_MODEL_MIX2 = 2;

// Sum over the minibatch of the per-datum log-likelihood gradients at `th`
// (and at `hat` when cv), all-reduced over the warp.  Rows are either the first n
// images of the keyed permutation (device sampling) or caller-supplied.
// SMEM: `base` is the CTA's shared-memory copy of the whole dataset (rows of ld floats); otherwise the dataset in global memory
template <int MODEL, bool SMEM, int D>
__device__ __forceinline__ void small_minibatch_sums(const ChainArgs& a, const float* __restrict__ base, int lane, const Perm& perm,
                                                     const int32_t* __restrict__ idx,
                                                     const float (&th)[(2 * D)], const float (&hat)[(2 * D)],
                                                     float (&s)[(2 * D)], float (&sh)[(2 * D)]) {
    const int d = a.d;
#pragma unroll
    for (int k = 0; k < (2 * D); ++k) { s[k] = 0.f; sh[k] = 0.f; }
    for (int j = lane; j < a.n; j += 32) {
        const uint32_t row = idx ? (uint32_t)idx[j] : perm_at(perm, (uint32_t)j);
        const float* __restrict__ xr = base + (size_t)row * a.ld;
        float x[D];
        if (SMEM) {
            if (a.ld == 2) {   // the vignette shape: one 8-byte shared load per row
                const float2 v = *reinterpret_cast<const float2*>(xr);
                x[0] = v.x; x[1] = v.y;
#pragma unroll
                for (int k = 2; k < D; ++k) x[k] = 0.f;
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) x[k] = (k < d) ? xr[k] : 0.f;
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = (k < d) ? __ldg(xr + k) : 0.f;
        }
        if (MODEL == SMALL_MODEL_GAUSS) {
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k < d) { s[k] += x[k] - th[k]; if (a.cv) sh[k] += x[k] - hat[k]; }
        } else {
            // responsibilities r_k = softmax_k(-0.5 |x - theta_k|^2) (equal weights cancel)
            float q1 = 0.f, q2 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k < d) {
                    const float e1 = x[k] - th[k], e2 = x[k] - th[D + k];
                    q1 += e1 * e1; q2 += e2 * e2;
                    if (a.cv) {
                        const float f1 = x[k] - hat[k], f2 = x[k] - hat[D + k];
                        h1 += f1 * f1; h2 += f2 * f2;
                    }
                }
            // p1 = w1 / (w1 + w2), w_k = exp(-q_k / 2)
            const float p1 = 1.0f / (1.0f + expf(0.5f * (q1 - q2)));
            const float p2 = 1.0f - p1;
            float g1 = 0.f, g2 = 0.f;
            if (a.cv) { g1 = 1.0f / (1.0f + expf(0.5f * (h1 - h2))); g2 = 1.0f - g1; }
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k < d) {
                    s[k] += p1 * (x[k] - th[k]);
                    s[D + k] += p2 * (x[k] - th[D + k]);
                    if (a.cv) {
                        sh[k] += g1 * (x[k] - hat[k]);
                        sh[D + k] += g2 * (x[k] - hat[D + k]);
                    }
                }
        }
    }
#pragma unroll
    for (int k = 0; k < (2 * D); ++k) {
        const bool live = (k % D) < d && (MODEL == SMALL_MODEL_MIX2 || k < D);
        if (live) { s[k] = warp_sum(s[k]); if (a.cv) sh[k] = warp_sum(sh[k]); }
    }
}

// d estLogPost / d theta (R/setup.R:116-125) from the minibatch sums.
template <int MODEL, int D>
__device__ __forceinline__ void small_grad_from_sums(const ChainArgs& a, const float (&th)[(2 * D)],
                                                     const float (&s)[(2 * D)], float (&g)[(2 * D)]) {
    const int d = a.d;
#pragma unroll
    for (int k = 0; k < (2 * D); ++k) {
        const int kk = k % D;
        g[k] = 0.f;
        if (kk >= d || (MODEL == SMALL_MODEL_GAUSS && k >= D)) continue;
        float ll = s[k];
        if (MODEL == SMALL_MODEL_GAUSS) {
            ll = 0.f;
#pragma unroll
            for (int m = 0; m < D; ++m)
                if (m < d) ll += a.prec[kk * SMALL_MAXD + m] * s[m];
        }
        const float pr = a.use_prior ? -th[k] * a.prior_inv_var : 0.f;
        g[k] = pr + a.corr * ll;
    }
}

// Register slot k <-> flat parameter index (mixture: two tensors of d entries each).
template <int MODEL, int D = SMALL_MAXD>
__device__ __forceinline__ int small_flat_of_slot(int k, int d) {
    if (MODEL == SMALL_MODEL_GAUSS) return k < d ? k : -1;
    const int t = k / D, kk = k % D;
    return kk < d ? t * d + kk : -1;
}

constexpr int SMALL_SMEM_THREADS = 512;                 // dataset-in-shared-memory variant: 16 chains (warps) per CTA, 2 CTAs per SM
constexpr size_t SMALL_SMEM_MAX_BYTES = 100 * 1024;     // dataset size up to which two CTAs (two copies) fit one SM
constexpr uint32_t SMALL_SMEM_CHUNK = 32 * 1024;        // bytes per TMA bulk copy
static inline size_t small_smem_bytes(int64_t N, int ld) { return (size_t)N * ld * sizeof(float) + 16; }

// D: the dimension the register arrays are sized for (2 for the vignettes' d <= 2: a quarter of the registers of the general D = 8
// instantiation, four times the resident warps).
// SMEM = true: the whole dataset (80 KB at the vignettes' N = 1e4, d = 2) is staged in shared memory ONCE per CTA by TMA bulk copies
// (mbarrier-tracked), and every gathered row is then one shared-memory load -- SURVEY.md section 7 ("rows of config 1-2 are 8 B: use
// float2 loads of an SMEM-resident dataset").  SMEM = false: rows come from global memory through the read-only path (datasets that
// do not fit, or whose byte size is not a multiple of 16).
// Launch bounds: the shared-memory variant is sized for TWO CTAs per SM (2 x 80 KB of dataset, 32 resident chains); with d <= 2 that
// means <= 64 registers per thread, which ptxas only honours when told (ptxas -v: Gaussian 64 registers + 28 B of spills without the
// bound, 47 and none with it; mixture 86 -> 64 with 18 B of spills outside the row loop: one CTA per SM -> two).
template <int MODEL, bool SMEM, int D>
__global__ void __launch_bounds__(SMEM ? SMALL_SMEM_THREADS : 128, (SMEM && D <= 2) ? 2 : 0) small_chain_kernel(const ChainArgs a) {
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    const int T = (MODEL == SMALL_MODEL_GAUSS) ? 1 : 2;
    const float SQRT_HALF = 0.70710678118654752440f;
    const float* __restrict__ rows_base = a.X;
    if (SMEM) {
        SGM_DYN_SMEM(unsigned char, small_smem, 128);
        const uint32_t data_s = smem_u32(small_smem);
        const uint32_t bytes = (uint32_t)a.N * (uint32_t)a.ld * 4u;                 // host guarantees: a multiple of 16, <= SMALL_SMEM_MAX_BYTES
        const uint32_t bar = data_s + ((bytes + 7u) & ~7u);
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_expect_tx(bar, bytes);
            for (uint32_t off = 0; off < bytes; off += SMALL_SMEM_CHUNK)
                bulk_g2s(data_s + off, reinterpret_cast<const unsigned char*>(a.X) + off, (bytes - off < SMALL_SMEM_CHUNK) ? bytes - off : SMALL_SMEM_CHUNK, bar);
        }
        __syncthreads();          // the barrier is initialised before anybody polls it
        mbar_wait(bar, 0);
        rows_base = reinterpret_cast<const float*>(small_smem);
        SGM_EMU_NOTE("small_chain:smem");
    }
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < a.C; c += warps_per_grid) {
        float th[(2 * D)], mo[(2 * D)], hat[(2 * D)], gh0[(2 * D)];
        float zeta[2];
#pragma unroll
        for (int k = 0; k < (2 * D); ++k) {
            const int f = small_flat_of_slot<MODEL, D>(k, a.d);
            th[k] = f >= 0 ? a.theta[(size_t)c * a.Pst + f] : 0.f;
            mo[k] = f >= 0 ? a.mom[(size_t)c * a.Pst + f] : 0.f;
            hat[k] = (f >= 0 && a.cv) ? a.theta_hat[(size_t)c * a.hat_stride + f] : 0.f;
            gh0[k] = (f >= 0 && a.cv) ? a.g_hat[(size_t)c * a.hat_stride + f] : 0.f;
        }
        zeta[0] = a.thermo[c * MAX_TENSORS + 0];
        zeta[1] = a.thermo[c * MAX_TENSORS + 1];
        bool bad = false;

        for (int it = 0; it < a.iters; ++it) {
            const uint64_t step = a.step0 + (uint64_t)it;
            auto noise_at = [&](int f, int sub) -> float {
                if (a.det) return a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + f];
                return normal_at((uint32_t)f, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1);
            };
            if (a.algo == ALGO_SGHMC) {  // momentum refresh, R/dynamics.R:62, R/update.R:52-57
#pragma unroll
                for (int k = 0; k < (2 * D); ++k) {
                    const int f = small_flat_of_slot<MODEL, D>(k, a.d);
                    if (f >= 0) mo[k] = 0.5f * (a.tc.sqrt_2eps[k / D] * noise_at(f, 0));
                }
            }
            for (int l = 0; l < a.L; ++l) {
                const int sub = l + 1;
                Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), a.index_tag, a.k0, a.k1);
                const int32_t* idx = a.det ? a.idx + (((size_t)c * a.iters + it) * a.L + l) * a.n : nullptr;
                // sgld / sghmc: one gradient evaluation per tensor (Gauss-Seidel, R/update.R:43-45,60-64);
                // sgnht / optimiser: all tensors from one evaluation (R/update.R:72-80, cV.R:87-88).
                const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
                const int passes = seidel ? T : 1;
                for (int pass = 0; pass < passes; ++pass) {
                    float s[(2 * D)], sh[(2 * D)], g[(2 * D)], gh[(2 * D)];
                    small_minibatch_sums<MODEL, SMEM, D>(a, rows_base, lane, perm, idx, th, hat, s, sh);
                    small_grad_from_sums<MODEL, D>(a, th, s, g);
                    if (a.cv) {
                        small_grad_from_sums<MODEL, D>(a, hat, sh, gh);
#pragma unroll
                        for (int k = 0; k < (2 * D); ++k) g[k] = (gh0[k] - gh[k]) + g[k];  // R/setup.R:58
                    }
#pragma unroll
                    for (int k = 0; k < (2 * D); ++k) {
                        const int t = k / D;
                        const int f = small_flat_of_slot<MODEL, D>(k, a.d);
                        if (f < 0 || (seidel && t != pass)) continue;
                        if (a.algo == ALGO_SGLD) {
                            const float z = noise_at(f, sub);
                            th[k] += a.tc.half_eps[t] * g[k] + SQRT_HALF * (a.tc.sqrt_2eps[t] * z);
                        } else if (a.algo == ALGO_SGHMC) {
                            const float z = noise_at(f, sub);
                            mo[k] += a.tc.eps[t] * g[k] - a.tc.alpha[t] * mo[k] + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * z);
                            th[k] += mo[k];
                        } else if (a.algo == ALGO_SGNHT) {
                            const float z = noise_at(f, sub);
                            mo[k] += a.tc.eps[t] * g[k] - mo[k] * zeta[t] + a.tc.sqrt_a[t] * (a.tc.sqrt_2eps[t] * z);
                        } else if (a.algo == ALGO_OPT) {
                            th[k] += a.opt_lr * g[k];
                        } else {  // ALGO_GRAD
                            mo[k] = g[k];
                        }
                    }
                    if (a.algo == ALGO_SGNHT) {  // theta pass then thermostat pass, R/update.R:75-80
                        float uu[2] = {0.f, 0.f};
#pragma unroll
                        for (int k = 0; k < (2 * D); ++k) {
                            if (small_flat_of_slot<MODEL, D>(k, a.d) < 0) continue;
                            th[k] += mo[k];
                            uu[k / D] += mo[k] * mo[k];
                        }
#pragma unroll
                        for (int t = 0; t < T; ++t) zeta[t] += uu[t] / (float)a.d - a.tc.eps[t];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < (2 * D); ++k) {
                const int f = small_flat_of_slot<MODEL, D>(k, a.d);
                if (f < 0) continue;
                bad |= !isfinite(th[k]);
                if (a.store && lane == (f & 31))
                    a.store[((size_t)(a.store_pos + it) * a.C + c) * a.P + f] = th[k];
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < (2 * D); ++k) {
                const int f = small_flat_of_slot<MODEL, D>(k, a.d);
                if (f < 0) continue;
                a.theta[(size_t)c * a.Pst + f] = th[k];
                a.mom[(size_t)c * a.Pst + f] = mo[k];
            }
            a.thermo[c * MAX_TENSORS + 0] = zeta[0];
            a.thermo[c * MAX_TENSORS + 1] = zeta[1];
            if (bad) *a.diverged = 1;
        }
    }
}

// estLogPost value (R/setup.R:116-125) incl. the densities' constants, one warp per chain.
template <int MODEL>
__global__ void small_logpost_kernel(const ChainArgs a, uint32_t word2, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= a.C) return;
    const int d = a.d;
    const float LOG_2PI = 1.8378770664093453f;
    float th[SMALL_MAXP];
#pragma unroll
    for (int k = 0; k < SMALL_MAXP; ++k) {
        const int f = small_flat_of_slot<MODEL>(k, d);
        th[k] = f >= 0 ? a.theta[(size_t)c * a.Pst + f] : 0.f;
    }
    Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2, TAG_CHECK_INDEX, a.k0, a.k1);
    const int32_t* idx = (a.det == 1) ? a.idx + (size_t)c * a.n : nullptr;
    float ll = 0.f;
    for (int j = lane; j < a.n; j += 32) {
        const uint32_t row = (a.det == 2) ? (uint32_t)j : (idx ? (uint32_t)idx[j] : perm_at(perm, (uint32_t)j));   // det == 2: rows 0..n-1 (held-out set)
        const float* xr = a.X + (size_t)row * a.ld;
        if (MODEL == SMALL_MODEL_GAUSS) {
            float q = 0.f;
            for (int k = 0; k < d; ++k)
                for (int m = 0; m < d; ++m) q += (xr[k] - th[k]) * a.prec[k * SMALL_MAXD + m] * (xr[m] - th[m]);
            ll += -0.5f * q;
        } else {
            float q1 = 0.f, q2 = 0.f;
            for (int k = 0; k < d; ++k) {
                const float e1 = xr[k] - th[k], e2 = xr[k] - th[SMALL_MAXD + k];
                q1 += e1 * e1; q2 += e2 * e2;
            }
            const float l1 = -0.5f * q1, l2 = -0.5f * q2, m = fmaxf(l1, l2);
            ll += m + logf(expf(l1 - m) + expf(l2 - m)) + logf(0.5f) - 0.5f * d * LOG_2PI;
        }
    }
    ll = warp_sum(ll);
    if (lane == 0) {
        float lp = 0.f;
        if (MODEL == SMALL_MODEL_GAUSS) {
            ll += (float)a.n * (0.5f * a.logdet_prec - 0.5f * d * LOG_2PI);
        }
        if (a.use_prior) {
            const float sd = rsqrtf(a.prior_inv_var);
#pragma unroll
            for (int k = 0; k < SMALL_MAXP; ++k)
                if (small_flat_of_slot<MODEL>(k, d) >= 0)
                    lp += -0.5f * th[k] * th[k] * a.prior_inv_var - logf(sd) - 0.5f * LOG_2PI;
        }
        out[c] = (double)(lp + a.corr * ll);
    }
}

}  // namespace sgm
