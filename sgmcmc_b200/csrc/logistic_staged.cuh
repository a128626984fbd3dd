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
#include <cuda.h>   // CUtensorMap
#include "common.cuh"
#include "logistic.cuh"

namespace sgm {

constexpr int LS_THREADS = 256;
constexpr int LS_WARPS = LS_THREADS / 32;
constexpr int LS_MAX_R = 4;   // rows per lane (n <= 1024)

// sigmoid with a full-precision exponential and a single-instruction reciprocal
__device__ __forceinline__ float sigmoid_fast(float z) { return __fdividef(1.0f, 1.0f + expf(-z)); }
__device__ __forceinline__ float fdot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}

__host__ __device__ inline int ls_row_stride4(int ld) { const int Q = ld >> 2; return (Q % 4 == 2) ? Q : (Q | 1); }
// coefficients [n] are followed by the scratch block, which holds the per-warp mbarriers (8-byte objects): n is rounded up to an even
// count there, or an ODD minibatch size would misalign them (a misaligned-address fault on the device; found by the emulated build)
__host__ __device__ inline int ls_coef_slots(int n) { return (n + 1) & ~1; }
static inline size_t ls_smem_bytes(int ld, int n) {
    return sizeof(float) * ((size_t)n * 4 * ls_row_stride4(ld) + 3 * (size_t)ld + (size_t)ls_coef_slots(n) + 64);
}

// scratch slots
constexpr int LS_BIAS = 40, LS_BIAS_HAT = 41, LS_UBIAS = 42, LS_PERM = 48;   // [48..51] and [56..59]: permutation keys of sub-steps s with s & 1 = 0 / 1

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

#define LS_TICK(slot)                                                                      \
    if (a.prof && tid == 0) { const long long t_ = clock64(); atomicAdd(a.prof + (slot), (unsigned long long)(t_ - t_prev)); t_prev = t_; }

// HAT_DOT: control variates whose residual at the mode is NOT cached in the rows (per-chain modes)
// R: rows per lane = ceil(n / 256)
template <bool HAT_DOT, int R>
__global__ void __launch_bounds__(LS_THREADS, 2) logistic_staged_kernel(const ChainArgs a, const __grid_constant__ CUtensorMap tmap) {
    SGM_DYN_SMEM(unsigned char, smem_raw, 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = a.d, ld = a.ld, n = a.n;
    const int Q = ld >> 2, S = ls_row_stride4(ld);
    const bool paired = (S & 1) == 0;      // unpadded rows: skewed lanes
    const int qt = d >> 2, ct = d & 3;     // float4 / slot holding y (and, for the parameters, the bias)
    const float SQRT_HALF = 0.70710678118654752440f;
    LogiStagedSmem sm;
    sm.rows = reinterpret_cast<float*>(smem_raw);
    sm.th = sm.rows + (size_t)n * 4 * S; sm.hat = sm.th + ld; sm.grad = sm.hat + ld; sm.coef = sm.grad + ld; sm.scratch = sm.coef + ls_coef_slots(n);
    const uint32_t rows_s = smem_u32(sm.rows), th_s = smem_u32(sm.th), hat_s = smem_u32(sm.hat), coef_s = smem_u32(sm.coef);
    const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
    const bool noisy = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC || a.algo == ALGO_SGNHT);
    const bool cached = a.cv && !HAT_DOT;   // r_hat rides in the row
    const int e = tid;                      // the parameter element this thread owns in the update (e <= d)
    const bool owner = e <= d;
    const int t_own = (e == d) ? 0 : 1, f_own = (e == d) ? 0 : e + 1;
    // gather mapping: one LDGSTS instruction covers rpi whole rows (lane = row-in-instruction * Q + float4 column)
    const int rpi = 32 / Q, g_row = lane / Q, g_q = lane - g_row * Q;
    const bool g_live = g_row < rpi;
    // phase-1 skew (unpadded rows only): row quads 4..7 (mod 8) start one float4 later, so the 8 lanes of a
    // shared-memory phase hit 8 different 16-byte bank groups although the row stride is even
    const int skew = paired ? ((tid >> 2) & 1) : 0;
    const uint32_t bar_s = smem_u32(sm.scratch + 16) + (uint32_t)warp * 8u;   // one mbarrier per warp (TMA gather mode)
    uint32_t bar_uses = 0;
    if (a.gather_tma) {
        if (lane == 0) mbar_init(bar_s, 1);
        fence_mbar_init();
    }
    long long t_prev = clock64();
    SGM_EMU_NOTE(a.gather_tma == 3 ? "logistic_staged:gather4" : (a.gather_tma == 0 ? "logistic_staged:cp.async" : "logistic_staged:bulk"));

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
        // Row indices are always computed one sub-step ahead (they do not depend on theta): the first ones here,
        // later ones while the previous minibatch is in flight.  Sub-step s of this launch = (it, l) = (s / L, s % L);
        // its permutation keys are published by the last warp in scratch slot s & 1, two sub-steps ahead.
        const int n_sub = a.iters * a.L;
        auto keys_of = [&](int s_) -> Perm {
            return make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(a.step0 + (uint64_t)(s_ / a.L), a.L, s_ % a.L + 1),
                             a.index_tag, a.k0, a.k1);
        };
        auto rows_of = [&](const Perm& p_, int s_, uint32_t (&out)[R]) {
            const int32_t* idx_ = a.det ? a.idx + ((size_t)c * n_sub + s_) * n : nullptr;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = r * LS_THREADS + tid;
                out[r] = 0;
                if (j < n && s_ < n_sub) out[r] = idx_ ? (uint32_t)idx_[j] : perm_at(p_, (uint32_t)j);
            }
        };
        uint32_t next_row[R];
        {
            Perm p0;
            if (!a.det) p0 = keys_of(0);
            rows_of(p0, 0, next_row);
            if (!a.det && warp == LS_WARPS - 1) {
                const Perm p1 = keys_of(1);
                if (lane < 4) reinterpret_cast<uint32_t*>(sm.scratch + LS_PERM + 8)[lane] = p1.key[lane];
            }
        }
        int s_cur = 0;

        for (int it = 0; it < a.iters; ++it) {
            const uint64_t step = a.step0 + (uint64_t)it;
            auto noise_own = [&](int sub) -> float {
                if (a.det) return a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + f_own];
                return normal_at((uint32_t)f_own, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1);
            };
            if (a.algo == ALGO_SGHMC && owner) mo_e = 0.5f * (a.tc.sqrt_2eps[t_own] * noise_own(0));   // refresh, R/dynamics.R:62
            for (int l = 0; l < a.L; ++l) {
                const int sub = l + 1;
                LS_TICK(0)   // store / loop overhead
                __syncthreads();   // previous minibatch fully consumed; th/hat/bias/keys published
                LS_TICK(1)   // sync before gather
                // ---- gather: each warp fetches the rows its lanes own (R x 32), everything in flight at once.
                // gather_tma: 1 = every round by TMA bulk copies, 0 = every round by 16-byte cp.async,
                //             2 = even rounds TMA / odd rounds cp.async, 3 = TMA tile::gather4 (4 rows per request; unpadded rows)
                auto via_tma = [&](int r) -> bool { return a.gather_tma == 1 || a.gather_tma == 3 || (a.gather_tma == 2 && (r & 1) == 0); };
                int tma_rows = 0;
                if (a.gather_tma) {
                    fence_proxy_async();
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (via_tma(r)) { int cnt = n - (r * LS_THREADS + warp * 32); tma_rows += cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt); }
                    if (lane == 0 && tma_rows > 0) mbar_expect_tx(bar_s, (uint32_t)tma_rows * (uint32_t)ld * 4u);
                    __syncwarp();
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int j = r * LS_THREADS + tid;
                    if (a.gather_tma == 3) {
                        // rows j..j+3 of a full group by one gather4 (issued by the lane with j % 4 == 0), leftovers singly
                        const int r1 = __shfl_down_sync(0xffffffffu, (int)next_row[r], 1);
                        const int r2 = __shfl_down_sync(0xffffffffu, (int)next_row[r], 2);
                        const int r3 = __shfl_down_sync(0xffffffffu, (int)next_row[r], 3);
                        if ((j | 3) < n) {
                            if ((lane & 3) == 0) tma_gather4(rows_s + (uint32_t)(j * S) * 16u, &tmap, 0, (int)next_row[r], r1, r2, r3, bar_s);
                        } else if (j < n) {
                            bulk_g2s(rows_s + (uint32_t)(j * S) * 16u, a.X + (size_t)next_row[r] * ld, (uint32_t)ld * 4u, bar_s);
                        }
                    } else if (via_tma(r)) {
                        if (j < n) bulk_g2s(rows_s + (uint32_t)(j * S) * 16u, a.X + (size_t)next_row[r] * ld, (uint32_t)ld * 4u, bar_s);
                    } else {
                        const int j0 = r * LS_THREADS + warp * 32;
                        const uint32_t my_row = next_row[r];
                        const uint32_t dst0 = rows_s + (uint32_t)((j0 + g_row) * S + g_q) * 16u;
                        for (int i = 0; i * rpi < 32; ++i) {
                            const int jr = i * rpi + g_row;
                            const uint32_t row = __shfl_sync(0xffffffffu, my_row, jr & 31);
                            if (g_live && jr < 32 && j0 + jr < n) {
                                const float* src = a.X + ((size_t)row * ld + 4 * g_q);
                                cp_async16(dst0 + (uint32_t)(i * rpi * S) * 16u, src);
                            }
                        }
                    }
                }
                cp_async_commit();
                LS_TICK(2)   // index computation + issue of the copies
                // ---- while the rows are in flight: this sub-step's noise and the next sub-step's permutation keys
                float z_own = 0.f;
                if (owner && noisy) z_own = noise_own(sub);
                if (!a.det && warp == LS_WARPS - 1 && s_cur + 2 < n_sub) {
                    const Perm p2 = keys_of(s_cur + 2);
                    if (lane < 4) reinterpret_cast<uint32_t*>(sm.scratch + LS_PERM + 8 * (s_cur & 1))[lane] = p2.key[lane];
                }
                {
                    Perm pn;
                    pn.N = a.N; pn.mask = (a.perm_bits >= 32) ? 0xFFFFFFFFu : ((1u << a.perm_bits) - 1u);
                    pn.shift = a.perm_bits / 2 > 0 ? a.perm_bits / 2 : 1;
#pragma unroll
                    for (int k = 0; k < 4; ++k) pn.key[k] = reinterpret_cast<const uint32_t*>(sm.scratch + LS_PERM + 8 * ((s_cur + 1) & 1))[k];
                    rows_of(pn, s_cur + 1, next_row);
                }
                ++s_cur;
                if (tma_rows > 0) { mbar_wait(bar_s, bar_uses & 1u); ++bar_uses; }
                cp_async_wait_all();
                __syncwarp();
                LS_TICK(3)   // wait for the rows
                // ---- phase 1: lane-per-row dots, R rows per lane walked together, two float4 columns per iteration
                uint32_t xa[R];
                float dz0[R], dz1[R], dh0[R], dh1[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int j = r * LS_THREADS + tid;
                    xa[r] = rows_s + (uint32_t)((j < n ? j : 0) * S) * 16u;   // out-of-range lanes recompute row 0 (discarded)
                    dz0[r] = dz1[r] = dh0[r] = dh1[r] = 0.f;
                }
                for (int k = 0; k + 1 < Q; k += 2) {
                    const int q0 = k + skew;
                    const int q1 = (q0 + 1 == Q) ? 0 : q0 + 1;
                    const float4 t0 = lds128(th_s + q0 * 16), t1 = lds128(th_s + q1 * 16);
                    float4 h0, h1;
                    if (HAT_DOT) { h0 = lds128(hat_s + q0 * 16); h1 = lds128(hat_s + q1 * 16); }
                    float4 x0[R], x1[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) { x0[r] = lds128(xa[r] + q0 * 16); x1[r] = lds128(xa[r] + q1 * 16); }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        dz0[r] = fdot4(x0[r], t0, dz0[r]); dz1[r] = fdot4(x1[r], t1, dz1[r]);
                        if (HAT_DOT) { dh0[r] = fdot4(x0[r], h0, dh0[r]); dh1[r] = fdot4(x1[r], h1, dh1[r]); }
                    }
                }
                if (Q & 1) {   // odd column count (padded stride, no skew): last column
                    const int q0 = Q - 1;
                    const float4 t0 = lds128(th_s + q0 * 16);
                    float4 h0;
                    if (HAT_DOT) h0 = lds128(hat_s + q0 * 16);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float4 x0 = lds128(xa[r] + q0 * 16);
                        dz0[r] = fdot4(x0, t0, dz0[r]);
                        if (HAT_DOT) dh0[r] = fdot4(x0, h0, dh0[r]);
                    }
                }
                float dz[R], yv[R], rhat[R];
                const float bias_hat = sm.scratch[LS_BIAS_HAT];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    dz[r] = dz0[r] + dz1[r];
                    const float4 tail = lds128(xa[r] + qt * 16);
                    yv[r] = get4(tail, ct);
                    rhat[r] = 0.f;
                    if (cached) rhat[r] = get4(tail, ct + 1);   // sits where beta is zero-padded: never enters dz
                    else if (HAT_DOT) rhat[r] = yv[r] - sigmoid_fast(bias_hat + (dh0[r] + dh1[r]));
                }
                LS_TICK(4)   // phase 1 dots
                const int passes = seidel ? 2 : 1;
                for (int pass = 0; pass < passes; ++pass) {
                    const float bias = sm.scratch[LS_BIAS];
                    float csum = 0.f;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int j = r * LS_THREADS + tid;
                        if (j < n) {
                            const float rr = yv[r] - sigmoid_fast(bias + dz[r]);
                            const float cf = a.cv ? rr - rhat[r] : rr;
                            csum += cf;
                            sts32(coef_s + j * 4, cf);
                        }
                    }
                    LS_TICK(5)   // sigmoid + coef
                    const float cs_tot = ls_block_sum(csum, sm.scratch);   // = gradient sum of the bias; also publishes coef
                    LS_TICK(6)   // block sum (barrier)
                    if (!(seidel && pass == 0)) {
                        // ---- phase 2: column sums of c_i * x_i (4 independent accumulators per lane)
                        const int rstep = paired ? 16 : 32;                 // rows between a lane's consecutive rows
                        const int rl = paired ? (lane >> 1) : lane;
                        const int cpar = paired ? (lane & 1) : 0;
                        const int cstep = paired ? 2 : 1;
                        for (int q0 = cstep * warp; q0 < Q; q0 += cstep * LS_WARPS) {
                            const uint32_t col = rows_s + (uint32_t)(q0 + cpar) * 16u;
                            float4 acc[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            int j = rl;
                            for (; j + 3 * rstep < n; j += 4 * rstep) {
                                float4 x[4]; float cf[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) { x[u] = lds128(col + (uint32_t)((j + u * rstep) * S) * 16u); cf[u] = lds32(coef_s + (j + u * rstep) * 4); }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    acc[u].x = fmaf(cf[u], x[u].x, acc[u].x); acc[u].y = fmaf(cf[u], x[u].y, acc[u].y);
                                    acc[u].z = fmaf(cf[u], x[u].z, acc[u].z); acc[u].w = fmaf(cf[u], x[u].w, acc[u].w);
                                }
                            }
                            for (; j < n; j += rstep) {
                                const float4 x = lds128(col + (uint32_t)(j * S) * 16u);
                                const float cf = lds32(coef_s + j * 4);
                                acc[0].x = fmaf(cf, x.x, acc[0].x); acc[0].y = fmaf(cf, x.y, acc[0].y);
                                acc[0].z = fmaf(cf, x.z, acc[0].z); acc[0].w = fmaf(cf, x.w, acc[0].w);
                            }
                            float4 t;
                            t.x = (acc[0].x + acc[1].x) + (acc[2].x + acc[3].x); t.y = (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y);
                            t.z = (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z); t.w = (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w);
#pragma unroll
                            for (int o = cstep; o < 32; o <<= 1) {
                                t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
                                t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
                            }
                            if (lane < cstep) *reinterpret_cast<float4*>(sm.grad + 4 * (q0 + cpar)) = t;
                        }
                        __syncthreads();
                    }
                    LS_TICK(7)   // phase 2 column sums (+ barrier)
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
                            th_e = th_e + (a.tc.half_eps[t] * g + SQRT_HALF * (a.tc.sqrt_2eps[t] * z_own));
                        } else if (a.algo == ALGO_SGHMC) {
                            mo_e += a.tc.eps[t] * g - a.tc.alpha[t] * mo_e + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * z_own);
                            th_e = th_e + mo_e;
                        } else if (a.algo == ALGO_SGNHT) {
                            mo_e += a.tc.eps[t] * g - mo_e * (t == 0 ? zeta_bias : zeta_beta) + a.tc.sqrt_a[t] * (a.tc.sqrt_2eps[t] * z_own);
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
                    LS_TICK(8)   // update
                    __syncthreads();
                    LS_TICK(9)   // sync after update
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
