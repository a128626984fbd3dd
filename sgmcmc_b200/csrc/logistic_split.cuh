// Logistic regression with FEW chains (or few modes): one thread-block CLUSTER per chain instead of one CTA per chain.
//
// A Markov chain is sequential: with one CTA per chain (logistic_staged.cuh / logistic.cuh) a user who runs the reference's way --
// ONE chain, or a handful -- uses 1 of 148 SMs per chain, and every step pays the full latency of its own gather (profiles/r01:
// 4.3 us per step at n = 500, d = 54; a 4 MB minibatch at d = 1024 through one SM's load path).  The same holds for the mode finder
// of the control-variate samplers when the chains share one mode (getMode / optUpdate, R/controlVariates.R:30-33,37-52: nItersOpt
// sequential SGD steps).  Here
//   * the S CTAs of a cluster each take n / S consecutive positions of every minibatch (the same keyed permutation, so the union is
//     exactly the minibatch the one-CTA kernels draw), sum their rows' contributions, and exchange the partial sums through
//     DISTRIBUTED SHARED MEMORY: every CTA publishes its vector in its own shared memory, one barrier.cluster, every CTA adds the S
//     vectors in rank order (identical sums everywhere) and applies the update to ITS OWN replica of the chain state -- the noise is
//     a function of (seed, chain, iteration, element), so the S replicas never drift apart;
//   * the rows do not depend on the state, so every warp keeps a ring of LSP_STAGES rounds of its rows in flight (one TMA bulk copy
//     per row, one mbarrier per warp and stage), running ahead ACROSS step boundaries: the gather latency leaves the critical
//     path, which is shared-memory dot products, one block reduction and one cluster barrier per gradient evaluation.
// The partial-sum buffers alternate from exchange to exchange (counted over everything a cluster works through): a CTA overwrites
// the buffer of exchange s at exchange s + 2, which it reaches only after the barrier of exchange s + 1, at which every peer has
// finished reading exchange s.  One cluster barrier per exchange suffices.
//
// Semantics are those of logistic.cuh, statement for statement (the reference's per-tensor Gauss-Seidel order of SGLD / SGHMC --
// bias pass, then beta pass on the same rows, R/update.R:43-45,60-64 --, control variates with the residual at a shared mode cached
// in the rows or recomputed, SGNHT thermostats, the mode finder's plain SGD, the validation gradient); only the order of the float32
// sums differs.  A dataset sharded over ranks and peer-mapped (common.cuh: row_ptr) is gathered the same way: the bulk copy's source
// is then a peer's memory and the copy crosses NVLink.
//
// Row -> thread mapping as in logistic.cuh: G lanes per row, lane lg owns the float4 columns lg + G v (v < V); the bias rides as a
// constant-1 feature in the slot y occupies.  Internal parameter order beta[0..d-1], bias at d.
#pragma once
#include "common.cuh"
#include "logistic.cuh"

namespace sgm {

constexpr int LSP_THREADS = 256;
constexpr int LSP_WARPS = LSP_THREADS / 32;
constexpr int LSP_STAGES = 4;          // rounds of rows in flight per warp
constexpr int LSP_MAX_SPLIT = 8;       // portable cluster size

// shared memory (floats): th | mo | hat | ghat [ldp each] | grad[2][ldp] | red[warps][ldp] | dots, rhat, ys [n_loc each] | scratch[32]
//                         | ring[warps][stages][32 / G rows][ld] | mbarriers[warps][stages]
static inline size_t lsp_smem_bytes(int G, int V, int ld, int n, int split) {
    const size_t ldp = (size_t)V * G * 4;
    const size_t n_loc = ((size_t)((n + split - 1) / split) + 3) & ~(size_t)3;
    return sizeof(float) * ((6 + LSP_WARPS) * ldp + 3 * n_loc + 32 + (size_t)LSP_WARPS * LSP_STAGES * (32 / G) * ld) + 8 * LSP_WARPS * LSP_STAGES;
}

template <int G, int V>
__global__ void __launch_bounds__(LSP_THREADS, 1) logistic_split_kernel(const ChainArgs a) {
    SGM_DYN_SMEM(float, lsp_smem, 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int RW = 32 / G;                 // rows per warp and round
    constexpr int NG = LSP_WARPS * RW;         // rows per CTA and round
    constexpr int ldp = V * G * 4;
    const int d = a.d, ld = a.ld, n = a.n;
    const int g_row = lane / G, lg = lane % G;
    const int qt = d >> 2, ct = d & 3, lg_t = qt % G, v_t = qt / G;   // float4 / slot holding y (and, for the parameters, the bias)
    const int S = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / S, n_clusters = gridDim.x / S;
    const float SQRT_HALF = 0.70710678118654752440f;
    // this CTA's positions of every minibatch: [j_lo, j_lo + cnt)
    const int n_loc = (n + S - 1) / S;
    const int n_loc4 = (n_loc + 3) & ~3;
    const int j_lo = rank * n_loc;
    const int cnt = (n - j_lo < 0) ? 0 : ((n - j_lo < n_loc) ? n - j_lo : n_loc);
    const int rps = (n_loc + NG - 1) / NG;     // rounds per pass over the minibatch, the same in every CTA of the cluster
    const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
    const int passes = seidel ? 2 : 1;
    const bool hat_dot = a.cv && !a.rhat_cached;   // else the residual at a shared mode rides in the row (column d + 1)
    float* th = lsp_smem;
    float* mo = th + ldp;
    float* hat = mo + ldp;
    float* ghat = hat + ldp;
    float* grad = ghat + ldp;                  // [2][ldp]
    float* red = grad + 2 * ldp;               // [warps][ldp]
    float* dots = red + LSP_WARPS * ldp;       // [n_loc] x . beta of the bias pass, kept for the beta pass
    float* rhat = dots + n_loc4;
    float* ys = rhat + n_loc4;
    float* scratch = ys + n_loc4;              // [32]
    float* ring = scratch + 32;
    const uint32_t ring_s = smem_u32(ring) + (uint32_t)(warp * LSP_STAGES * RW) * (uint32_t)ld * 4u;
    const uint32_t bars = smem_u32(ring + (size_t)LSP_WARPS * LSP_STAGES * RW * ld) + (uint32_t)(warp * LSP_STAGES) * 8u;
    const uint32_t grad_s = smem_u32(grad);
    const uint32_t row_bytes = (uint32_t)ld * 4u;
    if (lane == 0)
        for (int s = 0; s < LSP_STAGES; ++s) mbar_init(bars + 8u * s, 1);
    fence_mbar_init();
    __syncwarp();
    SGM_EMU_NOTE(a.algo == ALGO_OPT ? (a.n_shards > 1 ? "logistic_opt_split:sharded" : "logistic_opt_split") : "logistic_split:sample");
    uint32_t gg = 0;        // rounds consumed so far by this warp (stage = gg % LSP_STAGES, phase parity = (gg / LSP_STAGES) & 1)
    uint32_t tt = 0;        // exchanges finished so far by this CTA, over ALL its chains: the partial-sum buffer alternates with it
    bool bad = false;

    for (int c = cid; c < a.C; c += n_clusters) {
        __syncthreads();
        for (int e = tid; e < ldp; e += LSP_THREADS) {
            const bool live = e <= d;
            th[e] = live ? a.theta[(size_t)c * a.Pst + e] : 0.f;
            mo[e] = live ? a.mom[(size_t)c * a.Pst + e] : 0.f;
            hat[e] = (live && a.cv) ? a.theta_hat[(size_t)c * a.hat_stride + e] : 0.f;
            ghat[e] = (live && a.cv) ? a.g_hat[(size_t)c * a.hat_stride + e] : 0.f;
        }
        float zeta_bias = a.thermo[c * MAX_TENSORS + 0], zeta_beta = a.thermo[c * MAX_TENSORS + 1];
        __syncthreads();
        const int rounds_per_it = a.L * passes * rps;
        const int total = a.iters * rounds_per_it;       // rounds of this chain
        Perm perm; int perm_sub = -1;
        // round g of this chain -> (iteration, sub-step, pass, round of the pass); both passes of a sub-step fetch the same rows
        auto issue = [&](int g, uint32_t slot_no) {
            const int q = g / rps, r = g - q * rps;
            const int sub_abs = q / passes;                   // it * L + l
            const int it = sub_abs / a.L, l = sub_abs - it * a.L;
            const uint32_t st = slot_no % LSP_STAGES, bar = bars + 8u * st;
            if (!a.det && sub_abs != perm_sub) {
                perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(a.step0 + (uint64_t)it, a.L, l + 1), a.index_tag, a.k0, a.k1);
                perm_sub = sub_abs;
            }
            const int j0 = r * NG + warp * RW;                 // the warp's first local position of this round
            int live = cnt - j0; live = live < 0 ? 0 : (live > RW ? RW : live);
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)live * row_bytes);
            __syncwarp();
            if (lg == 0 && g_row < live) {
                const int jj = j_lo + j0 + g_row;              // position in the sub-step's minibatch
                const uint32_t row = a.det ? (uint32_t)a.idx[(((size_t)c * a.iters + it) * a.L + l) * n + jj] : perm_at(perm, (uint32_t)jj);
                bulk_g2s(ring_s + (st * RW + (uint32_t)g_row) * row_bytes, row_ptr(a, row), row_bytes, bar);
            }
        };
        for (int g = 0; g < LSP_STAGES && g < total; ++g) issue(g, gg + (uint32_t)g);

        float4 acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        float csum = 0.f;
        for (int g = 0; g < total; ++g, ++gg) {
            const int q = g / rps, r = g - q * rps;
            const int pass_i = q % passes, sub_abs = q / passes;
            const int it = sub_abs / a.L, l = sub_abs - it * a.L, sub = l + 1;
            const int pass = seidel ? (pass_i == 0 ? PASS_BIAS : PASS_BETA) : PASS_FULL;
            const uint64_t step = a.step0 + (uint64_t)it;
            auto noise_at = [&](int f, int sb) -> float {
                if (a.det) return a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sb * a.P : 0) + f];
                return normal_at((uint32_t)f, a.chain0 + (uint32_t)c, word2_of(step, a.L, sb), TAG_NOISE, a.k0, a.k1);
            };
            if (r == 0 && pass_i == 0 && l == 0 && a.algo == ALGO_SGHMC) {   // momentum refresh at the start of every sgmcmcStep, R/dynamics.R:62
                for (int e = tid; e <= d; e += LSP_THREADS) {
                    const int t = (e == d) ? 0 : 1, f = (e == d) ? 0 : e + 1;
                    mo[e] = 0.5f * (a.tc.sqrt_2eps[t] * noise_at(f, 0));
                }
                __syncthreads();
            }
            const uint32_t st = gg % LSP_STAGES;
            const int j0 = r * NG + warp * RW;
            const int jl = j0 + g_row;                         // local position of this lane group's row
            const bool valid = jl < cnt;
            mbar_wait(bars + 8u * st, (gg / LSP_STAGES) & 1u);
            const uint32_t xs = ring_s + (st * RW + (uint32_t)g_row) * row_bytes;
            float4 x[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int col = 4 * (lg + G * v);
                x[v] = (valid && col < ld) ? lds128(xs + 4u * (uint32_t)col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // y and the cached residual sit in the tail float4; blank the y slot (the parameters hold the bias there)
            float y = 0.f, rc = 0.f;
            if (lg == lg_t) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (v == v_t) { y = get4(x[v], ct); rc = get4(x[v], ct + 1); set4(x[v], ct, 0.f); }
            }
            const float bias = th[d], bias_hat = a.cv ? hat[d] : 0.f;
            float coef;
            if (pass == PASS_BETA) {
                const float r_ = (valid ? ys[jl] : 0.f) - sigmoidf(bias + (valid ? dots[jl] : 0.f));
                coef = a.cv ? r_ - (valid ? rhat[jl] : 0.f) : r_;
            } else {
                y = __shfl_sync(0xffffffffu, y, lg_t, G);
                float dz = 0.f, dh = 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int col = 4 * (lg + G * v);
                    dz += dot4(x[v], *reinterpret_cast<const float4*>(th + col));
                    if (hat_dot) dh += dot4(x[v], *reinterpret_cast<const float4*>(hat + col));
                }
                dz = group_sum<G>(dz);     // the bias slot of th / hat meets a blanked x slot: dz, dh are pure x . beta
                const float r_ = y - sigmoidf(bias + dz);
                float rh = 0.f;
                if (hat_dot) { dh = group_sum<G>(dh); rh = y - sigmoidf(bias_hat + dh); }
                else if (a.cv) rh = __shfl_sync(0xffffffffu, rc, lg_t, G);
                coef = a.cv ? r_ - rh : r_;
                if (pass == PASS_BIAS && valid && lg == 0) { dots[jl] = dz; rhat[jl] = rh; ys[jl] = y; }
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
            // the stage is consumed: refill it with the round LSP_STAGES ahead (which may belong to a later pass, sub-step or iteration)
            fence_proxy_async();
            __syncwarp();
            if (g + LSP_STAGES < total) issue(g + LSP_STAGES, gg + LSP_STAGES);
            if (r != rps - 1) continue;

            // ---- end of a pass: this CTA's partial sums -> the cluster -> the update of the elements the pass owns
            float* mine = grad + (tt & 1u) * ldp;
            if (pass == PASS_BIAS) {
                const float tot = block_sum(csum, scratch);
                csum = 0.f;
                if (tid == 0) mine[d] = tot;
            } else {
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
                    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                __syncthreads();
                for (int e = tid; e <= d; e += LSP_THREADS) {
                    float s = 0.f;
#pragma unroll
                    for (int w = 0; w < LSP_WARPS; ++w) s += red[(size_t)w * ldp + e];
                    mine[e] = s;
                }
            }
            cluster_arrive();      // release: the partial sums above are visible to the peers after their wait
            cluster_wait();
            float uu = 0.f;
            for (int e = tid; e <= d; e += LSP_THREADS) {
                const int t = (e == d) ? 0 : 1, f = (e == d) ? 0 : e + 1;
                if (seidel && t != pass_i) continue;
                float s = 0.f;
                for (int p = 0; p < S; ++p) s += ld_dsmem32(mapa(grad_s + ((tt & 1u) * (uint32_t)ldp + (uint32_t)e) * 4u, (uint32_t)p));
                const float cur = th[e];
                float gr = a.corr * s;
                if (a.use_prior) gr += -signf(cur);
                if (a.cv) {
                    if (a.use_prior) gr += signf(hat[e]);
                    gr += ghat[e];
                }
                if (a.algo == ALGO_SGLD) {
                    th[e] = cur + (a.tc.half_eps[t] * gr + SQRT_HALF * (a.tc.sqrt_2eps[t] * noise_at(f, sub)));
                } else if (a.algo == ALGO_SGHMC) {
                    float nu = mo[e];
                    nu += a.tc.eps[t] * gr - a.tc.alpha[t] * nu + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * noise_at(f, sub));
                    mo[e] = nu;
                    th[e] = cur + nu;
                } else if (a.algo == ALGO_SGNHT) {
                    float u = mo[e];
                    u += a.tc.eps[t] * gr - u * (t == 0 ? zeta_bias : zeta_beta) + a.tc.sqrt_a[t] * (a.tc.sqrt_2eps[t] * noise_at(f, sub));
                    mo[e] = u;
                    th[e] = cur + u;
                    if (t == 1) uu += u * u;
                } else if (a.algo == ALGO_OPT) {
                    th[e] = cur + a.opt_lr * gr;
                } else {  // ALGO_GRAD
                    mo[e] = gr;
                }
            }
            ++tt;
            if (a.algo == ALGO_SGNHT) {  // thermostats, R/dynamics.R:108-116 (every replica computes the same sums)
                const float tot = block_sum(uu, scratch);
                const float ub = mo[d];
                zeta_beta += tot / (float)d - a.tc.eps[1];
                zeta_bias += ub * ub - a.tc.eps[0];
            }
            __syncthreads();       // the new state before the next pass's dot products
            if (pass_i == passes - 1 && l == a.L - 1) {        // end of sgmcmcStep `it`
                for (int e = tid; e <= d; e += LSP_THREADS) {
                    const float v = th[e];
                    bad |= !isfinite(v);
                    if (a.store && rank == 0) a.store[((size_t)(a.store_pos + it) * a.C + c) * a.P + ((e == d) ? 0 : e + 1)] = v;
                }
            }
        }
        if (rank == 0) {
            for (int e = tid; e <= d; e += LSP_THREADS) {
                a.theta[(size_t)c * a.Pst + e] = th[e];
                if (a.algo != ALGO_OPT) a.mom[(size_t)c * a.Pst + e] = mo[e];
            }
            if (tid == 0 && a.algo != ALGO_OPT) { a.thermo[c * MAX_TENSORS + 0] = zeta_bias; a.thermo[c * MAX_TENSORS + 1] = zeta_beta; }
        }
    }
    if (bad) *a.diverged = 1;
    // nobody leaves while a peer may still read its partial sums
    cluster_arrive();
    cluster_wait();
}

}  // namespace sgm
