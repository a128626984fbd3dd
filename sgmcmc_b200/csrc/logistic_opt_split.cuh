// Mode finder of the control-variate samplers for FEW modes (one shared mode, or fewer modes than SMs): one thread-block
// CLUSTER per mode instead of one CTA per mode.
//
// getMode / optUpdate (R/controlVariates.R:30-33,37-52) is nItersOpt strictly sequential SGD steps, theta_hat += optStepsize *
// grad estLogPost(theta_hat) on a fresh minibatch each.  With one CTA per mode (logistic_staged.cuh / logistic.cuh in ALGO_OPT) a
// job that shares one mode runs that loop on 1 of 148 SMs, and every step pays the full latency of its own gather (profiles/
// r01: 4.3 us per step at n = 500, d = 54; a 4 MB minibatch at d = 1024 through one SM's load path).  Here
//   * the S CTAs of a cluster each take n / S consecutive positions of the step's minibatch (the same keyed permutation, so the
//     union is exactly the minibatch the one-CTA kernels draw), sum c_i [x_i, 1] over them, and exchange the P partial sums through
//     DISTRIBUTED SHARED MEMORY: every CTA publishes its vector in its own shared memory, one barrier.cluster, every CTA adds the
//     S vectors in rank order (identical sums everywhere, so the S replicas of theta_hat never drift apart) and applies the step;
//   * the rows do not depend on theta_hat, so every warp keeps a ring of OS_STAGES rounds of its rows in flight (one TMA bulk
//     copy per row, one mbarrier per warp and stage), running ahead ACROSS step boundaries: the gather latency leaves the
//     critical path, which is shared-memory dot products, one block reduction and one cluster barrier per step.
// The partial-sum buffers alternate from step to step (counted over all the modes a cluster works through): a CTA overwrites the
// buffer of its step s at step s + 2, which it reaches only after the barrier of step s + 1, at which every peer has finished
// reading step s.  One cluster barrier per step suffices.
//
// A dataset sharded over ranks and peer-mapped (common.cuh: row_ptr) is gathered the same way: the bulk copy's source is then a peer's
// memory and the copy crosses NVLink -- the scale-out config's shared mode is where one SM's load path hurt most.
//
// Row -> thread mapping as in logistic.cuh: G lanes per row, lane lg owns the float4 columns lg + G v (v < V); the bias rides
// as a constant-1 feature in the slot y occupies.  Internal parameter order beta[0..d-1], bias at d (logistic.cuh).
#pragma once
#include "common.cuh"
#include "logistic.cuh"

namespace sgm {

constexpr int OS_THREADS = 256;
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_STAGES = 4;          // rounds of rows in flight per warp
constexpr int OS_MAX_SPLIT = 8;       // portable cluster size

// shared memory: th[ldp] | grad[2][ldp] | red[warps][ldp] | ring[warps][stages][32 / G rows][ld] | mbarriers[warps][stages]
static inline size_t os_smem_bytes(int G, int V, int ld) {
    const size_t ldp = (size_t)V * G * 4;
    return sizeof(float) * ((3 + OS_WARPS) * ldp + (size_t)OS_WARPS * OS_STAGES * (32 / G) * ld) + 8 * OS_WARPS * OS_STAGES;
}

template <int G, int V>
__global__ void __launch_bounds__(OS_THREADS, 1) logistic_opt_split_kernel(const ChainArgs a) {
    SGM_DYN_SMEM(float, os_smem, 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int RW = 32 / G;                 // rows per warp and round
    constexpr int NG = OS_WARPS * RW;          // rows per CTA and round
    constexpr int ldp = V * G * 4;
    const int d = a.d, ld = a.ld, n = a.n;
    const int g_row = lane / G, lg = lane % G;
    const int qt = d >> 2, ct = d & 3, lg_t = qt % G, v_t = qt / G;   // float4 / slot holding y (and, for the parameters, the bias)
    const int S = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / S, n_clusters = gridDim.x / S;
    // this CTA's positions of the minibatch: [j_lo, j_lo + cnt)
    const int n_loc = (n + S - 1) / S;
    const int j_lo = rank * n_loc;
    const int cnt = (n - j_lo < 0) ? 0 : ((n - j_lo < n_loc) ? n - j_lo : n_loc);
    const int rps = (n_loc + NG - 1) / NG;     // rounds per step, the same in every CTA of the cluster
    float* th = os_smem;
    float* grad = th + ldp;                    // [2][ldp]
    float* red = grad + 2 * ldp;               // [warps][ldp]
    float* ring = red + OS_WARPS * ldp;
    const uint32_t ring_s = smem_u32(ring) + (uint32_t)(warp * OS_STAGES * RW) * (uint32_t)ld * 4u;
    const uint32_t bars = smem_u32(ring + (size_t)OS_WARPS * OS_STAGES * RW * ld) + (uint32_t)(warp * OS_STAGES) * 8u;
    const uint32_t grad_s = smem_u32(grad);
    const uint32_t row_bytes = (uint32_t)ld * 4u;
    if (lane == 0)
        for (int s = 0; s < OS_STAGES; ++s) mbar_init(bars + 8u * s, 1);
    fence_mbar_init();
    __syncwarp();
    SGM_EMU_NOTE(a.n_shards > 1 ? "logistic_opt_split:sharded" : "logistic_opt_split");
    uint32_t gg = 0;        // rounds consumed so far by this warp (stage = gg % OS_STAGES, phase parity = (gg / OS_STAGES) & 1)
    uint32_t tt = 0;        // steps finished so far by this CTA, over ALL its modes: the partial-sum buffer alternates with it (a new mode
                            // must not start on the buffer the previous mode's last step may still be read from)
    bool bad = false;

    for (int c = cid; c < a.C; c += n_clusters) {
        __syncthreads();
        for (int e = tid; e < ldp; e += OS_THREADS) th[e] = (e <= d) ? a.theta[(size_t)c * a.Pst + e] : 0.f;
        __syncthreads();
        const int total = a.iters * rps;       // rounds of this mode
        Perm perm; int perm_step = -1;
        // round g of this mode (step g / rps) -> stage of the warp's ring; rows past the CTA's share are skipped (an arrival without bytes)
        auto issue = [&](int g, uint32_t slot_no) {
            const int t = g / rps, r = g - t * rps;
            const uint32_t st = slot_no % OS_STAGES, bar = bars + 8u * st;
            if (!a.det && t != perm_step) {
                perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(a.step0 + (uint64_t)t, 1, 1), a.index_tag, a.k0, a.k1);
                perm_step = t;
            }
            const int j0 = r * NG + warp * RW;                 // the warp's first local position of this round
            int live = cnt - j0; live = live < 0 ? 0 : (live > RW ? RW : live);
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)live * row_bytes);
            __syncwarp();
            if (lg == 0 && g_row < live) {
                const int jj = j_lo + j0 + g_row;              // position in the step's minibatch
                const uint32_t row = a.det ? (uint32_t)a.idx[((size_t)c * a.iters + t) * n + jj] : perm_at(perm, (uint32_t)jj);
                bulk_g2s(ring_s + (st * RW + (uint32_t)g_row) * row_bytes, row_ptr(a, row), row_bytes, bar);
            }
        };
        for (int g = 0; g < OS_STAGES && g < total; ++g) issue(g, gg + (uint32_t)g);

        float4 acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g = 0; g < total; ++g, ++gg) {
            const int t = g / rps, r = g - t * rps;
            const uint32_t st = gg % OS_STAGES;
            const int j0 = r * NG + warp * RW;
            const bool valid = j0 + g_row < cnt;
            mbar_wait(bars + 8u * st, (gg / OS_STAGES) & 1u);
            const uint32_t xs = ring_s + (st * RW + (uint32_t)g_row) * row_bytes;
            float4 x[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int col = 4 * (lg + G * v);
                x[v] = (valid && col < ld) ? lds128(xs + 4u * (uint32_t)col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float y = 0.f;
            if (lg == lg_t) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (v == v_t) { y = get4(x[v], ct); set4(x[v], ct, 0.f); if (ct < 3) set4(x[v], ct + 1, 0.f); }   // y, and the cached-residual slot next to it
            }
            y = __shfl_sync(0xffffffffu, y, lg_t, G);
            float dz = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) dz += dot4(x[v], *reinterpret_cast<const float4*>(th + 4 * (lg + G * v)));
            dz = group_sum<G>(dz);     // the bias slot of th meets a blanked x slot: dz is x . beta
            const float coef = valid ? y - sigmoidf(th[d] + dz) : 0.f;
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
            // the stage is consumed: refill it with the round OS_STAGES ahead (which may belong to a later step)
            fence_proxy_async();
            __syncwarp();
            if (g + OS_STAGES < total) issue(g + OS_STAGES, gg + OS_STAGES);

            if (r == rps - 1) {
                // ---- end of step t: row groups of the warp -> warps -> this CTA's partial sums -> the cluster
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
                float* mine = grad + (tt & 1u) * ldp;
                for (int e = tid; e <= d; e += OS_THREADS) {
                    float s = 0.f;
#pragma unroll
                    for (int w = 0; w < OS_WARPS; ++w) s += red[(size_t)w * ldp + e];
                    mine[e] = s;
                }
                cluster_arrive();      // release: the partial sums above are visible to the peers after their wait
                cluster_wait();
                for (int e = tid; e <= d; e += OS_THREADS) {
                    float s = 0.f;
                    for (int p = 0; p < S; ++p) s += ld_dsmem32(mapa(grad_s + ((tt & 1u) * (uint32_t)ldp + (uint32_t)e) * 4u, (uint32_t)p));
                    const float cur = th[e];
                    float gr = a.corr * s;
                    if (a.use_prior) gr += -signf(cur);
                    const float nv = cur + a.opt_lr * gr;
                    th[e] = nv;
                    if (t == a.iters - 1) bad |= !isfinite(nv);
                }
                ++tt;
                __syncthreads();       // the new theta_hat before the next step's dot products
            }
        }
        if (rank == 0)
            for (int e = tid; e <= d; e += OS_THREADS) a.theta[(size_t)c * a.Pst + e] = th[e];
    }
    if (bad) *a.diverged = 1;
    // nobody leaves while a peer may still read its partial sums
    cluster_arrive();
    cluster_wait();
}

}  // namespace sgm
