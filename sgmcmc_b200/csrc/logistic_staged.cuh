// Logistic regression, short rows (ld <= 128 floats): the whole minibatch is staged in
// shared memory by TMA bulk copies, one CTA per chain, lane-per-row arithmetic.
//
//   gather   every thread issues ONE cp.async.bulk (global -> shared, ld*4 bytes) for its own row of the
//            minibatch (index = keyed permutation, i.e. sampling without replacement) and the warp waits
//            on its own mbarrier: a warp starts computing as soon as ITS 32 rows have landed.
//   phase 1  lane-per-row dot products x.beta (and x.beta_hat for control variates) -> c_i = y - s(z)
//            (minus y - s(z_hat)); the rows are read with conflict-free 16-byte shared loads (odd
//            float4 row stride), the parameters with broadcast loads.
//   phase 2  column mapping: warp q sums c_i * x_i[4q..4q+3] over all rows (lanes stride rows), one
//            shuffle tree per warp -> gradient.  No per-thread accumulator arrays, no second gather.
//   update   fused SGLD / SGHMC / SGNHT / mode-finder step with Philox noise.
// The reference's per-tensor Gauss-Seidel order (bias, then beta; R/update.R:43-45,60-64) costs only a
// second sigmoid per row here: the dots stay in registers, the rows stay in shared memory.
//
// Replaces the same reference code as logistic.cuh (dataFeed/dataSelect R/update.R:166-194,
// tf$gradients R/setup.R:36-62, declareDynamics.* R/dynamics.R:7-120, sgmcmcStep.* R/update.R:40-81).
#pragma once
#include "common.cuh"
#include "logistic.cuh"

namespace sgm {

constexpr int LS_THREADS = 512;
constexpr int LS_WARPS = LS_THREADS / 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared::cta, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct LogiStagedSmem {
    float* rows;   // [n][S] S = odd number of float4 per row
    float* th;     // [ld]
    float* hat;    // [ld]
    float* ghat;   // [ld]
    float* mo;     // [ld]
    float* grad;   // [ld]
    float* coef;   // [n]
    float* scratch;  // [64]
    uint64_t* bars;  // [LS_WARPS]
};

__host__ __device__ inline int ls_row_stride(int ld) { return 4 * ((ld / 4) | 1); }
__host__ __device__ inline int ls_rounds(int n) { return (n + LS_THREADS - 1) / LS_THREADS; }
static inline size_t ls_smem_bytes(int ld, int n) {
    return sizeof(float) * ((size_t)n * ls_row_stride(ld) + 5 * (size_t)ld + (size_t)n + 64) + sizeof(uint64_t) * LS_WARPS + 16;
}

__global__ void __launch_bounds__(LS_THREADS, 1) logistic_staged_kernel(const ChainArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = a.d, ld = a.ld, n = a.n;
    const int S = ls_row_stride(ld), Q = ld >> 2;
    const int qt = d >> 2, ct = d & 3;   // float4 / slot holding y (and, for the parameters, the bias)
    const int rounds = ls_rounds(n);
    const float SQRT_HALF = 0.70710678118654752440f;
    LogiStagedSmem sm;
    sm.rows = reinterpret_cast<float*>(smem_raw);
    sm.th = sm.rows + (size_t)n * S; sm.hat = sm.th + ld; sm.ghat = sm.hat + ld; sm.mo = sm.ghat + ld; sm.grad = sm.mo + ld;
    sm.coef = sm.grad + ld; sm.scratch = sm.coef + n;
    sm.bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sm.scratch + 64) + 15) & ~(uintptr_t)15);
    if (tid < LS_WARPS) mbar_init(sm.bars + tid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;   // uses of this warp's mbarrier so far (its parity is the low bit)
    const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
    const uint32_t row_bytes = (uint32_t)ld * 4u;

    for (int c = blockIdx.x; c < a.C; c += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < ld; e += LS_THREADS) {
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
            auto noise_at = [&](int f, int sub) -> float {
                if (a.det) return a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + f];
                return normal_at((uint32_t)f, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1);
            };
            if (a.algo == ALGO_SGHMC) {
                for (int e = tid; e <= d; e += LS_THREADS) {
                    const int t = (e == d) ? 0 : 1, f = (e == d) ? 0 : e + 1;
                    sm.mo[e] = 0.5f * (a.tc.sqrt_2eps[t] * noise_at(f, 0));
                }
            }
            for (int l = 0; l < a.L; ++l) {
                const int sub = l + 1;
                const Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), a.index_tag, a.k0, a.k1);
                const int32_t* idx = a.det ? a.idx + (((size_t)c * a.iters + it) * a.L + l) * n : nullptr;
                // ---- gather: the shared rows were last read through the generic proxy
                __syncthreads();
                fence_proxy_async();
                float dotv[4], rhat[4], yv[4];   // per owned row (rounds <= 4)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    dotv[r] = 0.f; rhat[r] = 0.f; yv[r] = 0.f;
                    if (r >= rounds) continue;
                    const int j = r * LS_THREADS + tid;
                    const int j_warp0 = r * LS_THREADS + warp * 32;
                    int cnt = n - j_warp0; cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
                    if (cnt > 0) {
                        if (lane == 0) mbar_expect_tx(sm.bars + warp, (uint32_t)cnt * row_bytes);
                        __syncwarp();
                        if (j < n) {
                            const uint32_t row = idx ? (uint32_t)idx[j] : perm_at(perm, (uint32_t)j);
                            bulk_g2s(sm.rows + (size_t)j * S, a.X + (size_t)row * ld, row_bytes, sm.bars + warp);
                        }
                        mbar_wait(sm.bars + warp, phase & 1u);
                        ++phase;
                    }
                    // ---- phase 1 on the rows this warp just received
                    float dz = 0.f, dh = 0.f, y = 0.f;
                    if (j < n) {
                        const float4* xr = reinterpret_cast<const float4*>(sm.rows + (size_t)j * S);
#pragma unroll 2
                        for (int q = 0; q < Q; ++q) {
                            float4 x = xr[q];
                            if (q == qt) { y = get4(x, ct); set4(x, ct, 0.f); }
                            dz += dot4(x, reinterpret_cast<const float4*>(sm.th)[q]);
                            if (a.cv) dh += dot4(x, reinterpret_cast<const float4*>(sm.hat)[q]);
                        }
                    }
                    dotv[r] = dz; yv[r] = y;
                    rhat[r] = a.cv ? y - sigmoidf(sm.hat[d] + dh) : 0.f;
                }
                const int passes = seidel ? 2 : 1;
                for (int pass = 0; pass < passes; ++pass) {
                    // the bias slot of th is excluded from dz above (its x slot was blanked)
                    const float bias = sm.th[d];
                    float csum = 0.f;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int j = r * LS_THREADS + tid;
                        if (r < rounds && j < n) {
                            const float rr = yv[r] - sigmoidf(bias + dotv[r]);
                            const float cf = a.cv ? rr - rhat[r] : rr;
                            csum += cf;
                            sm.coef[j] = cf;
                        }
                    }
                    if (seidel && pass == 0) {
                        const float tot = block_sum(csum, sm.scratch);
                        if (tid == 0) sm.grad[d] = tot;
                        __syncthreads();
                    } else {
                        __syncthreads();
                        // ---- phase 2: warp q sums c_i * x_i[4q..4q+3]; lanes stride the rows
                        for (int q = warp; q < Q; q += LS_WARPS) {
                            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int j = lane; j < n; j += 32) {
                                float4 x = *reinterpret_cast<const float4*>(sm.rows + (size_t)j * S + 4 * q);
                                if (q == qt) set4(x, ct, 1.0f);   // the bias "feature" sits where y is stored
                                const float cf = sm.coef[j];
                                acc.x += cf * x.x; acc.y += cf * x.y; acc.z += cf * x.z; acc.w += cf * x.w;
                            }
                            acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
                            if (lane == 0) *reinterpret_cast<float4*>(sm.grad + 4 * q) = acc;
                        }
                        __syncthreads();
                    }
                    // ---- fused update of the tensor(s) this pass owns
                    float uu = 0.f;
                    for (int e = tid; e <= d; e += LS_THREADS) {
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
                        } else {
                            sm.mo[e] = g;
                        }
                    }
                    if (a.algo == ALGO_SGNHT) {
                        const float tot = block_sum(uu, sm.scratch);
                        const float ub = sm.mo[d];
                        zeta_beta += tot / (float)d - a.tc.eps[1];
                        zeta_bias += ub * ub - a.tc.eps[0];
                    }
                    __syncthreads();
                }
            }
            for (int e = tid; e <= d; e += LS_THREADS) {
                const float v = sm.th[e];
                bad |= !isfinite(v);
                if (a.store) a.store[((size_t)(a.store_pos + it) * a.C + c) * a.P + ((e == d) ? 0 : e + 1)] = v;
            }
        }
        __syncthreads();
        for (int e = tid; e <= d; e += LS_THREADS) {
            a.theta[(size_t)c * a.Pst + e] = sm.th[e];
            a.mom[(size_t)c * a.Pst + e] = sm.mo[e];
        }
        if (tid == 0) { a.thermo[c * MAX_TENSORS + 0] = zeta_bias; a.thermo[c * MAX_TENSORS + 1] = zeta_beta; }
        if (bad) *a.diverged = 1;
    }
}

}  // namespace sgm
