// Logistic regression, long rows (ld > 128 floats, e.g. d = 1024 -> 4112-byte rows): one CTA per chain,
// warp-specialised TMA ring.
//
//   producer  warp 0: computes the minibatch row indices 32 at a time (keyed permutation) and issues ONE
//             cp.async.bulk (TMA, SASS UBLKCP) per row into a ring of single-row shared-memory stages;
//             full / empty mbarriers per stage, the ring runs continuously across sub-steps and iterations.
//   consumers 12 warps, one row at a time each: 16-byte conflict-free shared loads, beta in registers,
//             shuffle-reduced dot, c_i = y_i - s(z_i) (- cached r_hat_i for control variates with a shared
//             mode), per-lane column accumulators; the stage is released as soon as the row sits in registers.
//   update    accumulators -> shared (the idle ring is reused) -> fused dynamics with Philox noise.
// In-flight depth is set by the ring (32 rows = 131 KB per SM), not by registers, so HBM latency is covered
// with one CTA per SM.  SGLD / SGHMC keep the reference's bias-then-beta order (R/update.R:43-45,60-64) by
// running the ring twice per sub-step (dots kept in shared memory, rows re-gathered from L2/HBM).
//
// Replaces the same reference code as logistic.cuh.
#pragma once
#include "common.cuh"
#include "logistic.cuh"
#include "logistic_staged.cuh"

namespace sgm {

constexpr int LR_CONSUMERS = 12;
constexpr int LR_THREADS = 32 * (1 + LR_CONSUMERS);
constexpr int LR_STAGES = 32;                      // rows in the ring
constexpr int LR_GROUP = 4;                        // rows per mbarrier pair (one wait / expect_tx per group)
constexpr int LR_NGROUPS = LR_STAGES / LR_GROUP;

static inline size_t lr_smem_bytes(int ld, int n) {
    return sizeof(float) * ((size_t)LR_STAGES * ld + 4 * (size_t)ld + 3 * (size_t)n + 64) + 16 * LR_STAGES + 32;
}

// V = float4 columns per lane (32 * V >= ld / 4); HAT_DOT: control variates without the cached residual
template <int V, bool HAT_DOT>
__global__ void __launch_bounds__(LR_THREADS, 1) logistic_ring_kernel(const ChainArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = a.d, ld = a.ld, n = a.n, Q = ld >> 2;
    const int qt = d >> 2, ct = d & 3;
    const float SQRT_HALF = 0.70710678118654752440f;
    float* ring = reinterpret_cast<float*>(smem_raw);          // [LR_STAGES][ld]
    float* th = ring + (size_t)LR_STAGES * ld;                  // [ld] beta, bias slot and padding zeroed
    float* hat = th + ld;
    float* grad = hat + ld;
    float* mo = grad + ld;
    float* dots = mo + ld;                                      // [n] x.beta      (two-pass order only)
    float* rh = dots + n;                                       // [n] r_hat
    float* ys = rh + n;                                         // [n]
    float* scratch = ys + n;                                    // [64]
    const uint32_t bars = (smem_u32(scratch + 64) + 7u) & ~7u;  // full[s] at +16 s, empty[s] at +16 s + 8
    const uint32_t ring_s = smem_u32(ring);
    const uint32_t row_bytes = (uint32_t)ld * 4u;
    if (tid < LR_NGROUPS) { mbar_init(bars + 16u * tid, 1); mbar_init(bars + 16u * tid + 8u, LR_GROUP); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
    const bool noisy = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC || a.algo == ALGO_SGNHT);
    const bool cached = a.cv && !HAT_DOT;
    uint32_t gg = 0;   // row groups pushed through the ring so far (producer and consumers count alike)

    for (int c = blockIdx.x; c < a.C; c += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < ld; e += LR_THREADS) {
            const bool live = e <= d;
            th[e] = (e < d) ? a.theta[(size_t)c * a.Pst + e] : 0.f;
            mo[e] = live ? a.mom[(size_t)c * a.Pst + e] : 0.f;
            hat[e] = (e < d && a.cv) ? a.theta_hat[(size_t)c * a.hat_stride + e] : 0.f;
        }
        if (tid == 0) {
            scratch[40] = a.theta[(size_t)c * a.Pst + d];
            scratch[41] = a.cv ? a.theta_hat[(size_t)c * a.hat_stride + d] : 0.f;
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
                for (int e = tid; e <= d; e += LR_THREADS) {
                    const int t = (e == d) ? 0 : 1, f = (e == d) ? 0 : e + 1;
                    mo[e] = 0.5f * (a.tc.sqrt_2eps[t] * noise_at(f, 0));
                }
                __syncthreads();
            }
            for (int l = 0; l < a.L; ++l) {
                const int sub = l + 1;
                const int32_t* idx = a.det ? a.idx + (((size_t)c * a.iters + it) * a.L + l) * n : nullptr;
                const int passes = seidel ? 2 : 1;
                for (int pass = 0; pass < passes; ++pass) {
                    const int mode = seidel ? (pass == 0 ? PASS_BIAS : PASS_BETA) : PASS_FULL;
                    const float bias = scratch[40], bias_hat = scratch[41];
                    float csum = 0.f;
                    float4 acc[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (warp == 0) {
                        // ================= producer =================
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the ring was just used as reduction scratch
                        const Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), a.index_tag, a.k0, a.k1);
                        for (int base = 0; base < n; base += 32) {
                            const int j = base + lane;
                            uint32_t my_row = 0;
                            if (j < n) my_row = idx ? (uint32_t)idx[j] : perm_at(perm, (uint32_t)j);
                            const int cnt = (n - base < 32) ? n - base : 32;
                            for (int k0 = 0; k0 < cnt; k0 += LR_GROUP) {
                                const uint32_t gi = gg + (uint32_t)((base + k0) / LR_GROUP);
                                const uint32_t s = gi % LR_NGROUPS, use = gi / LR_NGROUPS;
                                const int rows_here = (cnt - k0 < LR_GROUP) ? cnt - k0 : LR_GROUP;
                                if (use > 0) mbar_wait(bars + 16u * s + 8u, (use - 1u) & 1u);   // the previous occupants were released
                                if (lane == 0) {
                                    mbar_expect_tx(bars + 16u * s, (uint32_t)rows_here * row_bytes);
                                    if (rows_here < LR_GROUP)   // a short last group: arrive for the rows nobody will consume
                                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 16u * s + 8u), "r"(LR_GROUP - rows_here) : "memory");
                                }
#pragma unroll
                                for (int k = 0; k < LR_GROUP; ++k) {
                                    const uint32_t row = __shfl_sync(0xffffffffu, my_row, (k0 + k) & 31);
                                    if (lane == 0 && k < rows_here)
                                        bulk_g2s(ring_s + (s * LR_GROUP + k) * row_bytes, a.X + (size_t)row * ld, row_bytes, bars + 16u * s);
                                }
                            }
                        }
                    } else {
                        // ================= consumers =================
                        const int cw = warp - 1;
                        float4 tr[V];
#pragma unroll
                        for (int v = 0; v < V; ++v)
                            tr[v] = (lane + 32 * v < Q) ? *reinterpret_cast<const float4*>(th + 4 * (lane + 32 * v)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int j = cw; j < n; j += LR_CONSUMERS) {
                            const uint32_t gi = gg + (uint32_t)(j / LR_GROUP);
                            const uint32_t s = gi % LR_NGROUPS, use = gi / LR_NGROUPS;
                            mbar_wait(bars + 16u * s, use & 1u);
                            const uint32_t xa = ring_s + (s * LR_GROUP + (uint32_t)(j % LR_GROUP)) * row_bytes;
                            float4 x[V];
                            float y = 0.f, rc = 0.f;
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                const int q = lane + 32 * v;
                                x[v] = (q < Q) ? lds128(xa + q * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
                                if (q == qt) { y = get4(x[v], ct); if (cached) rc = get4(x[v], ct + 1); }
                            }
                            __syncwarp();
                            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 16u * s + 8u) : "memory");   // stage free
                            y = __shfl_sync(0xffffffffu, y, qt & 31);
                            float coef;
                            if (mode == PASS_BETA) {
                                const float r = ys[j] - sigmoid_fast(bias + dots[j]);
                                coef = a.cv ? r - rh[j] : r;
                            } else {
                                float dz = 0.f, dh = 0.f;
#pragma unroll
                                for (int v = 0; v < V; ++v) {
                                    dz = fdot4(x[v], tr[v], dz);
                                    if (HAT_DOT) { const int q = lane + 32 * v; if (q < Q) dh = fdot4(x[v], *reinterpret_cast<const float4*>(hat + 4 * q), dh); }
                                }
                                dz = warp_sum(dz);
                                float rhv = 0.f;
                                if (cached) rhv = __shfl_sync(0xffffffffu, rc, qt & 31);
                                if (HAT_DOT) { dh = warp_sum(dh); rhv = y - sigmoid_fast(bias_hat + dh); }
                                const float r = y - sigmoid_fast(bias + dz);
                                coef = a.cv ? r - rhv : r;
                                if (mode == PASS_BIAS && lane == 0) { dots[j] = dz; rh[j] = rhv; ys[j] = y; }
                            }
                            csum += coef;
                            if (mode != PASS_BIAS) {
#pragma unroll
                                for (int v = 0; v < V; ++v) {
                                    acc[v].x = fmaf(coef, x[v].x, acc[v].x); acc[v].y = fmaf(coef, x[v].y, acc[v].y);
                                    acc[v].z = fmaf(coef, x[v].z, acc[v].z); acc[v].w = fmaf(coef, x[v].w, acc[v].w);
                                }
                            }
                        }
                    }
                    gg += (uint32_t)((n + LR_GROUP - 1) / LR_GROUP);
                    __syncthreads();   // every row consumed: the ring is idle
                    // ---- reduce the consumers' accumulators through the idle ring: [LR_CONSUMERS][ld]
                    if (warp > 0) {
                        if (mode != PASS_BIAS) {
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                const int q = lane + 32 * v;
                                if (q < Q) *reinterpret_cast<float4*>(ring + (size_t)(warp - 1) * ld + 4 * q) = acc[v];
                            }
                        }
                        if (lane == 0) scratch[warp - 1] = csum;
                    }
                    __syncthreads();
                    float cs_tot = 0.f;
#pragma unroll
                    for (int w = 0; w < LR_CONSUMERS; ++w) cs_tot += scratch[w];
                    if (mode != PASS_BIAS) {
                        for (int e = tid; e < ld; e += LR_THREADS) {
                            float s2 = 0.f;
#pragma unroll
                            for (int w = 0; w < LR_CONSUMERS; ++w) s2 += ring[(size_t)w * ld + e];
                            grad[e] = s2;
                        }
                    }
                    __syncthreads();
                    // ---- fused update of the tensor(s) this pass owns (bias gradient = sum of the coefficients)
                    float uu = 0.f;
                    for (int e = tid; e <= d; e += LR_THREADS) {
                        const int t = (e == d) ? 0 : 1, f = (e == d) ? 0 : e + 1;
                        if (seidel && t != pass) continue;
                        const float thv = (e == d) ? bias : th[e];
                        const float hv = (e == d) ? bias_hat : hat[e];
                        float g = a.corr * ((e == d) ? cs_tot : grad[e]);
                        if (a.use_prior) g += -signf(thv);
                        if (a.cv) {
                            if (a.use_prior) g += signf(hv);
                            g += a.g_hat[(size_t)c * a.hat_stride + e];
                        }
                        float nv = thv;
                        const float z = noisy ? noise_at(f, sub) : 0.f;
                        if (a.algo == ALGO_SGLD) {
                            nv = thv + (a.tc.half_eps[t] * g + SQRT_HALF * (a.tc.sqrt_2eps[t] * z));
                        } else if (a.algo == ALGO_SGHMC) {
                            float nu = mo[e];
                            nu += a.tc.eps[t] * g - a.tc.alpha[t] * nu + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * z);
                            mo[e] = nu;
                            nv = thv + nu;
                        } else if (a.algo == ALGO_SGNHT) {
                            float u = mo[e];
                            u += a.tc.eps[t] * g - u * (t == 0 ? zeta_bias : zeta_beta) + a.tc.sqrt_a[t] * (a.tc.sqrt_2eps[t] * z);
                            mo[e] = u;
                            nv = thv + u;
                            if (t == 1) uu += u * u;
                        } else if (a.algo == ALGO_OPT) {
                            nv = thv + a.opt_lr * g;
                        } else {
                            mo[e] = g;
                        }
                        if (e == d) scratch[40] = nv; else th[e] = nv;
                    }
                    if (a.algo == ALGO_SGNHT) {
                        const float tot = block_sum(uu, scratch + 16);
                        const float ub = mo[d];
                        zeta_beta += tot / (float)d - a.tc.eps[1];
                        zeta_bias += ub * ub - a.tc.eps[0];
                    }
                    __syncthreads();
                }
            }
            for (int e = tid; e <= d; e += LR_THREADS) {
                const float v = (e == d) ? scratch[40] : th[e];
                bad |= !isfinite(v);
                if (a.store) a.store[((size_t)(a.store_pos + it) * a.C + c) * a.P + ((e == d) ? 0 : e + 1)] = v;
            }
        }
        __syncthreads();
        for (int e = tid; e <= d; e += LR_THREADS) {
            a.theta[(size_t)c * a.Pst + e] = (e == d) ? scratch[40] : th[e];
            a.mom[(size_t)c * a.Pst + e] = mo[e];
        }
        if (tid == 0) { a.thermo[c * MAX_TENSORS + 0] = zeta_bias; a.thermo[c * MAX_TENSORS + 1] = zeta_beta; }
        if (bad) *a.diverged = 1;
    }
}

}  // namespace sgm
