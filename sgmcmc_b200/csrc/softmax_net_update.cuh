// Softmax-net family (softmax_net.cu): priors, the fused SGLD / SGHMC / SGNHT / mode-finder updates of the eight parameter tensors
// (streamed through shared memory for B, in registers / shared memory for the small ones), shared-memory layout of a chain.
#pragma once
#include "softmax_net_stages.cuh"

namespace sgm {
__device__ float nn_sumsq(const float* __restrict__ w, int count, float* red) {
    float s = 0.f;
    for (int i = threadIdx.x; i < count; i += NN_THREADS) { const float v = w[i]; s += v * v; }
    return nn_block_sum(s, red);
}

__device__ __forceinline__ void nn_tensor_range(const NNDims& dm, int t, int& lo, int& hi) {
    const int offs[9] = {dm.offA, dm.offB, dm.offa, dm.offb, dm.offl, dm.offl + 1, dm.offl + 2, dm.offl + 3, dm.offl + 4};
    lo = offs[t]; hi = offs[t + 1];
}

// prior gradient of internal element e at th; ssq = ||W||^2 for a lambda element
__device__ __forceinline__ float nn_prior_grad(const NNDims& dm, const float* __restrict__ th, int e, int t, float ssq) {
    if (t < 4) return -th[dm.offl + t] * th[e];
    int lo, hi;
    nn_tensor_range(dm, t - 4, lo, hi);
    return (float)(hi - lo) / (2.0f * th[e]) - 0.5f * ssq - 1.0f;
}

// operand staging region: max(S1: two stages of A hi/lo [4][128] + B hi/lo [4][Npad+2] chunks, S5: A hi/lo [16][128] + B hi/lo [16][65])
__host__ __device__ inline int nn_tile_floats(const NNDims& dm) {
    const int Npad = (dm.dh + 15) & ~15;
    const int s1 = 2 * (2 * (16 / 4) * NN_RB * 4 + 2 * (16 / 4) * (Npad + 2) * 4);
    const int s5 = 2 * (64 / 4) * NN_RB * 4 + 2 * (64 / 4) * 65 * 4;
    return s1 > s5 ? s1 : s5;
}

struct NNChainState {
    float* th; float* mo; const float* hat; const float* ghat; float* G; float* Ghat;
    float zeta[MAX_TENSORS];
    // ||W||^2 of the weight tensors: at theta when the last update of tensor t produced it (bit t of ssq_valid), at the mode
    // once per launch (theta_hat never moves while sampling)
    float ssq[4], ssq_hat[4];
    unsigned ssq_valid, ssq_hat_valid;
    // Gauss-Seidel fast path without control variates: the small tensors never round-trip through global memory on the critical
    // path.  theta of A / a / b lives in sm.A / sm.At / sm.a / sm.b (the copies S2 and S4 read), their momenta in registers
    // (small index si = tid + 512 slot), the four precisions and their momenta in every thread; all writes go through to global.
    bool fast;
    float mo_reg[3];
    float lam_th[4], lam_mo[4];
};

// small index <-> internal element: A first (e < offB), then a, b, lambdas (e >= offa)
__device__ __forceinline__ int nn_small_elem(const NNDims& dm, int si) { return si < dm.offB ? si : si + (dm.offa - dm.offB); }
__device__ __forceinline__ int nn_small_count(const NNDims& dm) { return dm.offB + (dm.P - dm.offa); }

// fused dynamics update of tensor t (R/dynamics.R) from the gradient sums in G (and Ghat for cv)
__device__ void nn_update_tensor(const ChainArgs& a, const NNDims& dm, const NNSmem& sm, NNChainState& st, int t, int c,
                                 uint64_t step, int sub, float* uu_out) {
    const float SQRT_HALF = 0.70710678118654752440f;
    int lo, hi;
    nn_tensor_range(dm, t, lo, hi);
    float ssq = 0.f, ssq_hat = 0.f;
    int ssq_count = 0;
    if (t >= 4 && a.use_prior) {
        int wlo, whi;
        nn_tensor_range(dm, t - 4, wlo, whi);
        ssq_count = whi - wlo;
        ssq = (st.ssq_valid >> (t - 4)) & 1u ? st.ssq[t - 4] : nn_sumsq(st.th + wlo, whi - wlo, sm.red);
        if (a.cv) {
            if (!((st.ssq_hat_valid >> (t - 4)) & 1u)) { st.ssq_hat[t - 4] = nn_sumsq(st.hat + wlo, whi - wlo, sm.red); st.ssq_hat_valid |= 1u << (t - 4); }
            ssq_hat = st.ssq_hat[t - 4];
        }
    }
    if (t < 4) st.ssq_valid &= ~(1u << t);   // this tensor moves now
    float uu = 0.f;
    const bool noisy = a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC || a.algo == ALGO_SGNHT;
    // hoisted uniforms: the precision this weight tensor's prior uses (t < 4) at theta and at the mode
    const float lam_t = (t < 4 && a.use_prior) ? (st.fast ? st.lam_th[t] : st.th[dm.offl + t]) : 0.f;
    const float lam_h = (t < 4 && a.use_prior && a.cv) ? st.hat[dm.offl + t] : 0.f;
    const bool vec = (lo & 3) == 0;   // 16-byte aligned tensor: whole float4 groups are loaded before any store
    // ---- large 16-byte aligned tensors (B): stream theta / momentum / gradient sums through shared memory with TMA bulk copies
    // (a deep ring of 2048-element chunks: ~100 KB in flight per SM without holding a register), update in shared memory,
    // bulk-store theta / momentum back.  The ring is the H | dZ1 | tile region, dead between a gradient evaluation and the next S1.
    const bool wr_th_ = a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC || a.algo == ALGO_OPT;
    const bool wr_mo_ = a.algo == ALGO_SGHMC || a.algo == ALGO_SGNHT || a.algo == ALGO_GRAD;
    const int narr = a.cv ? 6 : 3;
    constexpr int UCH = 4 * NN_THREADS;                         // elements per chunk: one float4 group per thread
    const uint32_t ustage = (uint32_t)narr * UCH * 4u;
    const int ustages = (int)((sm.ring_bytes) / ustage) < 8 ? (int)(sm.ring_bytes / ustage) : 8;
    if (t < 4 && (lo & 3) == 0 && ((hi - lo) & 3) == 0 && (hi - lo) >= 8 * UCH && ustages >= 3 && nn_stream_updates) {
        const int count = hi - lo, nchunks = (count + UCH - 1) / UCH;
        const uint32_t bars = smem_u32(sm.s1_bars);
        fence_proxy_async_global();   // this CTA's generic stores (gradient sums, state) -> the bulk reads
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int s2 = 0; s2 < ustages; ++s2) mbar_init(bars + 8 * s2, 1);
            fence_mbar_init();
        }
        __syncthreads();
        const float* src[6] = {st.th + lo, st.mo + lo, st.G + lo, a.cv ? st.hat + lo : nullptr, a.cv ? st.ghat + lo : nullptr, a.cv ? st.Ghat + lo : nullptr};
        auto issue = [&](int k) {
            const int s2 = k % ustages;
            const int n_el = (count - k * UCH < UCH) ? count - k * UCH : UCH;
            const uint32_t bytes = (uint32_t)n_el * 4u, bar = bars + 8 * s2, dst = sm.ring + (uint32_t)s2 * ustage;
            mbar_expect_tx(bar, bytes * (uint32_t)narr);
            for (int q = 0; q < narr; ++q)
                bulk_g2s(dst + (uint32_t)q * UCH * 4u, src[q] + (size_t)k * UCH, bytes, bar);
        };
        if (threadIdx.x == 0) for (int k = 0; k < ustages - 1 && k < nchunks; ++k) issue(k);
        float ssq_new = 0.f;
        for (int k = 0; k < nchunks; ++k) {
            const int s2 = k % ustages;
            const uint32_t stg = sm.ring + (uint32_t)s2 * ustage + 16u * threadIdx.x;
            const int e0 = lo + k * UCH + 4 * (int)threadIdx.x;
            mbar_wait(bars + 8 * s2, (uint32_t)(k / ustages) & 1u);
            if (e0 < hi) {
                float z4[4] = {0.f, 0.f, 0.f, 0.f};
                if (noisy && !a.det) normal4_at((uint32_t)(e0 >> 2), a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1, z4);
                if (noisy && a.det)
                    for (int q = 0; q < 4; ++q)
                        z4[q] = a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + nn_flat_of_internal(dm, e0 + q)];
                const float4 tv = lds128(stg), mv = lds128(stg + UCH * 4u), gv = lds128(stg + 2u * UCH * 4u);
                float th4[4] = {tv.x, tv.y, tv.z, tv.w}, mo4[4] = {mv.x, mv.y, mv.z, mv.w};
                const float g4[4] = {gv.x, gv.y, gv.z, gv.w};
                float hat4[4] = {0.f, 0.f, 0.f, 0.f}, ghat4[4] = {0.f, 0.f, 0.f, 0.f}, gh4[4] = {0.f, 0.f, 0.f, 0.f};
                if (a.cv) {
                    const float4 hv = lds128(stg + 3u * UCH * 4u), qv = lds128(stg + 4u * UCH * 4u), pv = lds128(stg + 5u * UCH * 4u);
                    hat4[0] = hv.x; hat4[1] = hv.y; hat4[2] = hv.z; hat4[3] = hv.w; ghat4[0] = qv.x; ghat4[1] = qv.y; ghat4[2] = qv.z; ghat4[3] = qv.w;
                    gh4[0] = pv.x; gh4[1] = pv.y; gh4[2] = pv.z; gh4[3] = pv.w;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {   // the same arithmetic, in the same order, as the register path below
                    const float th = th4[q];
                    float g = a.corr * g4[q];
                    if (a.use_prior) g += -lam_t * th;
                    if (a.cv) {
                        float gh = a.corr * gh4[q];
                        if (a.use_prior) gh += -lam_h * hat4[q];
                        g = (ghat4[q] - gh) + g;   // R/setup.R:58
                    }
                    const float z = z4[q];
                    if (a.algo == ALGO_SGLD) {
                        th4[q] = th + (a.tc.half_eps[t] * g + SQRT_HALF * (a.tc.sqrt_2eps[t] * z));
                    } else if (a.algo == ALGO_SGHMC) {
                        float nu = mo4[q];
                        nu += a.tc.eps[t] * g - a.tc.alpha[t] * nu + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * z);
                        mo4[q] = nu;
                        th4[q] = th + nu;
                    } else if (a.algo == ALGO_SGNHT) {
                        float u = mo4[q];
                        u += a.tc.eps[t] * g - u * st.zeta[t] + a.tc.sqrt_a[t] * (a.tc.sqrt_2eps[t] * z);
                        mo4[q] = u;
                        uu += u * u;
                    } else if (a.algo == ALGO_OPT) {
                        th4[q] = th + a.opt_lr * g;
                    } else {
                        mo4[q] = g;
                    }
                    ssq_new += th4[q] * th4[q];
                }
                if (wr_th_) sts128(stg, make_float4(th4[0], th4[1], th4[2], th4[3]));
                if (wr_mo_) sts128(stg + UCH * 4u, make_float4(mo4[0], mo4[1], mo4[2], mo4[3]));
            }
            fence_proxy_async();   // generic stores -> the bulk store's reads
            __syncthreads();
            if (threadIdx.x == 0) {
                const int n_el = (count - k * UCH < UCH) ? count - k * UCH : UCH;
                const uint32_t bytes = (uint32_t)n_el * 4u, srcs = sm.ring + (uint32_t)s2 * ustage;
                if (wr_th_) bulk_s2g(st.th + lo + (size_t)k * UCH, srcs, bytes);
                if (wr_mo_) bulk_s2g(st.mo + lo + (size_t)k * UCH, srcs + UCH * 4u, bytes);
                bulk_commit_group();
                // the stage of chunk k - 1 is free once ITS store has read shared memory: refill it with chunk k - 1 + ustages
                if (k >= 1 && k - 1 + ustages < nchunks) {
                    bulk_wait_group_read<1>();
                    issue(k - 1 + ustages);
                } else if (k == 0 && ustages - 1 < nchunks) {
                    issue(ustages - 1);   // the last stage was never used
                }
            }
        }
        if (threadIdx.x == 0) {
            bulk_wait_group<0>();
            fence_proxy_async_global();   // bulk stores (async proxy) -> later generic loads of the state
        }
        __syncthreads();
        if (wr_th_) {   // (SGNHT moves theta in a later pass: its ||W||^2 is not the one accumulated here)
            st.ssq[t] = nn_block_sum(ssq_new, sm.red);
            st.ssq_valid |= 1u << t;
        }
    } else {
    // software pipeline: the loads of the NEXT group of four elements are issued before the current group's noise / arithmetic /
    // stores (one group at a time left ~24 KB in flight per SM and the pass latency-bound, profiles/r01_c4_phases.txt)
    struct Grp { float th4[4], mo4[4], g4[4], gh4[4], hat4[4], ghat4[4]; };
    auto load_group = [&](int e0, Grp& q) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { q.th4[i] = q.mo4[i] = q.g4[i] = q.gh4[i] = q.hat4[i] = q.ghat4[i] = 0.f; }
        if (e0 >= hi) return;
        const int cnt = (hi - e0 < 4) ? hi - e0 : 4;
        if (vec && cnt == 4) {
            const float4 tv = *reinterpret_cast<const float4*>(st.th + e0), mv = *reinterpret_cast<const float4*>(st.mo + e0);
            q.th4[0] = tv.x; q.th4[1] = tv.y; q.th4[2] = tv.z; q.th4[3] = tv.w; q.mo4[0] = mv.x; q.mo4[1] = mv.y; q.mo4[2] = mv.z; q.mo4[3] = mv.w;
            if (t < 4) { const float4 gv = *reinterpret_cast<const float4*>(st.G + e0); q.g4[0] = gv.x; q.g4[1] = gv.y; q.g4[2] = gv.z; q.g4[3] = gv.w; }
            if (a.cv) {
                const float4 hv = *reinterpret_cast<const float4*>(st.hat + e0), qv = *reinterpret_cast<const float4*>(st.ghat + e0);
                q.hat4[0] = hv.x; q.hat4[1] = hv.y; q.hat4[2] = hv.z; q.hat4[3] = hv.w; q.ghat4[0] = qv.x; q.ghat4[1] = qv.y; q.ghat4[2] = qv.z; q.ghat4[3] = qv.w;
                if (t < 4) { const float4 gv = *reinterpret_cast<const float4*>(st.Ghat + e0); q.gh4[0] = gv.x; q.gh4[1] = gv.y; q.gh4[2] = gv.z; q.gh4[3] = gv.w; }
            }
        } else {
            for (int i = 0; i < cnt; ++i) {
                q.th4[i] = st.th[e0 + i]; q.mo4[i] = st.mo[e0 + i];
                if (t < 4) q.g4[i] = st.G[e0 + i];
                if (a.cv) { q.hat4[i] = st.hat[e0 + i]; q.ghat4[i] = st.ghat[e0 + i]; if (t < 4) q.gh4[i] = st.Ghat[e0 + i]; }
            }
        }
    };
    auto process = [&](int e0, Grp& cur) {
        if (e0 >= hi) return;
        float z4[4] = {0.f, 0.f, 0.f, 0.f};
        if (noisy && !a.det) {
            // Philox block = 4 consecutive INTERNAL elements (tensor ranges need not be 4-aligned)
            float zz[4];
            normal4_at((uint32_t)(e0 >> 2), a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1, zz);
            if ((e0 & 3) == 0) { z4[0] = zz[0]; z4[1] = zz[1]; z4[2] = zz[2]; z4[3] = zz[3]; }
            else {
                float z2[4];
                normal4_at((uint32_t)(e0 >> 2) + 1u, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1, z2);
                for (int q = 0; q < 4; ++q) { const int ee = (e0 & 3) + q; z4[q] = ee < 4 ? zz[ee] : z2[ee - 4]; }
            }
        }
        const int cnt = (hi - e0 < 4) ? hi - e0 : 4;
        float (&th4)[4] = cur.th4; float (&mo4)[4] = cur.mo4; float (&g4)[4] = cur.g4;
        float (&gh4)[4] = cur.gh4; float (&hat4)[4] = cur.hat4; float (&ghat4)[4] = cur.ghat4;
        if (noisy && a.det)
            for (int q = 0; q < cnt; ++q)
                z4[q] = a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + nn_flat_of_internal(dm, e0 + q)];
        // ---- arithmetic
        for (int q = 0; q < cnt; ++q) {
            const float th = th4[q];
            float g = (t < 4) ? a.corr * g4[q] : 0.f;
            if (a.use_prior) g += (t < 4) ? -lam_t * th : ((float)(ssq_count) / (2.0f * th) - 0.5f * ssq - 1.0f);
            if (a.cv) {
                float gh = (t < 4) ? a.corr * gh4[q] : 0.f;
                if (a.use_prior) gh += (t < 4) ? -lam_h * hat4[q] : ((float)(ssq_count) / (2.0f * hat4[q]) - 0.5f * ssq_hat - 1.0f);
                g = (ghat4[q] - gh) + g;   // R/setup.R:58
            }
            const float z = z4[q];
            if (a.algo == ALGO_SGLD) {
                th4[q] = th + (a.tc.half_eps[t] * g + SQRT_HALF * (a.tc.sqrt_2eps[t] * z));
            } else if (a.algo == ALGO_SGHMC) {
                float nu = mo4[q];
                nu += a.tc.eps[t] * g - a.tc.alpha[t] * nu + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * z);
                mo4[q] = nu;
                th4[q] = th + nu;
            } else if (a.algo == ALGO_SGNHT) {
                float u = mo4[q];
                u += a.tc.eps[t] * g - u * st.zeta[t] + a.tc.sqrt_a[t] * (a.tc.sqrt_2eps[t] * z);
                mo4[q] = u;
                uu += u * u;
            } else if (a.algo == ALGO_OPT) {
                th4[q] = th + a.opt_lr * g;
            } else {
                mo4[q] = g;
            }
        }
        // ---- stores
        const bool wr_th = a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC || a.algo == ALGO_OPT;
        const bool wr_mo = a.algo == ALGO_SGHMC || a.algo == ALGO_SGNHT || a.algo == ALGO_GRAD;
        if (vec && cnt == 4) {
            if (wr_th) *reinterpret_cast<float4*>(st.th + e0) = make_float4(th4[0], th4[1], th4[2], th4[3]);
            if (wr_mo) *reinterpret_cast<float4*>(st.mo + e0) = make_float4(mo4[0], mo4[1], mo4[2], mo4[3]);
        } else {
            for (int q = 0; q < cnt; ++q) { if (wr_th) st.th[e0 + q] = th4[q]; if (wr_mo) st.mo[e0 + q] = mo4[q]; }
        }
    };
    {   // two register sets in ping-pong (no moves of in-flight load results)
        Grp ga, gb;
        const int stride = 4 * NN_THREADS;
        load_group(lo + 4 * (int)threadIdx.x, ga);
        for (int e0 = lo + 4 * (int)threadIdx.x; e0 < hi; e0 += 2 * stride) {
            load_group(e0 + stride, gb);
            process(e0, ga);
            load_group(e0 + 2 * stride, ga);
            process(e0 + stride, gb);
        }
    }
    }
    if (uu_out) *uu_out = nn_block_sum(uu, sm.red);
    __syncthreads();
    NN_TICK(t == 1 ? 10 : (t >= 4 ? 11 : 5))
}

// ---- fused updates of the small tensors (fast path, see NNChainState): tensor t in {A = 0, a = 2, b = 3}
__device__ void nn_update_small(const ChainArgs& a, const NNDims& dm, const NNSmem& sm, NNChainState& st, int t, int c, uint64_t step, int sub) {
    const float SQRT_HALF = 0.70710678118654752440f;
    int lo, hi;
    nn_tensor_range(dm, t, lo, hi);
    const float lam = a.use_prior ? st.lam_th[t] : 0.f;
    float ssq_new = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int si = (int)threadIdx.x + NN_THREADS * k;
        const int e = nn_small_elem(dm, si);
        if (si < nn_small_count(dm) && e >= lo && e < hi) {
            float th, g;
            int ia = 0;
            if (t == 0) { ia = (e % dm.dh) * dm.dop + e / dm.dh; th = sm.A[ia]; g = sm.gA[ia]; }
            else if (t == 2) { th = sm.a[e - dm.offa]; g = sm.ga[e - dm.offa]; }
            else { th = sm.b[e - dm.offb]; g = sm.gb[e - dm.offb]; }
            g = a.corr * g;
            if (a.use_prior) g += -lam * th;
            const float z = a.det ? a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + nn_flat_of_internal(dm, e)]
                                  : normal_at((uint32_t)e, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1);
            if (a.algo == ALGO_SGLD) {
                th = th + (a.tc.half_eps[t] * g + SQRT_HALF * (a.tc.sqrt_2eps[t] * z));
            } else {   // ALGO_SGHMC
                float nu = st.mo_reg[k];
                nu += a.tc.eps[t] * g - a.tc.alpha[t] * nu + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * z);
                st.mo_reg[k] = nu;
                st.mo[e] = nu;
                th = th + nu;
            }
            st.th[e] = th;
            if (t == 0) { sm.A[ia] = th; sm.At[(e / dm.dh) * dm.dh + e % dm.dh] = th; }
            else if (t == 2) sm.a[e - dm.offa] = th;
            else sm.b[e - dm.offb] = th;
            ssq_new += th * th;
        }
    }
    st.ssq[t] = nn_block_sum(ssq_new, sm.red);   // (its barriers also publish the new values in shared memory)
    st.ssq_valid |= 1u << t;
    NN_TICK(5)
}

// precision t - 4 (scalar): every thread keeps the value and its momentum; thread 0 writes them through
__device__ void nn_update_lambda(const ChainArgs& a, const NNDims& dm, const NNSmem& sm, NNChainState& st, int t, int c, uint64_t step, int sub) {
    const float SQRT_HALF = 0.70710678118654752440f;
    const int w = t - 4, e = dm.offl + w;
    int wlo, whi;
    nn_tensor_range(dm, w, wlo, whi);
    float th = st.lam_th[w], g = 0.f;
    if (a.use_prior) {
        const float ssq = (st.ssq_valid >> w) & 1u ? st.ssq[w] : nn_sumsq(st.th + wlo, whi - wlo, sm.red);
        g = (float)(whi - wlo) / (2.0f * th) - 0.5f * ssq - 1.0f;
    }
    const float z = a.det ? a.noise[(size_t)c * a.noise_len + (a.algo == ALGO_SGHMC ? sub * a.P : 0) + nn_flat_of_internal(dm, e)]
                          : normal_at((uint32_t)e, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), TAG_NOISE, a.k0, a.k1);
    if (a.algo == ALGO_SGLD) {
        th = th + (a.tc.half_eps[t] * g + SQRT_HALF * (a.tc.sqrt_2eps[t] * z));
    } else {
        float nu = st.lam_mo[w];
        nu += a.tc.eps[t] * g - a.tc.alpha[t] * nu + a.tc.sqrt_alpha[t] * (a.tc.sqrt_2eps[t] * z);
        st.lam_mo[w] = nu;
        if (threadIdx.x == 0) st.mo[e] = nu;
        th = th + nu;
    }
    st.lam_th[w] = th;
    if (threadIdx.x == 0) st.th[e] = th;
    NN_TICK(11)
}

__host__ __device__ inline int nn_tile_floats_all(const NNDims& dm) {
    const int t1 = nn_tile_floats(dm);
    const int need = nn_s1_ring_bytes(dm) / 4 - 2 * NN_RB * dm.dh;   // the ring spans H | dZ1 | tile
    return t1 > need ? t1 : need;
}

__device__ void nn_carve(const NNDims& dm, float* base, NNSmem& sm) {
    // 1024-byte aligned start: the swizzled operand tiles need it (the launch reserves the slack)
    const uint32_t b0 = smem_u32(base);
    float* p = base + (((1024u - (b0 & 1023u)) & 1023u) >> 2);
    const int t1 = nn_tile_floats_all(dm);
    sm.ring = smem_u32(p);
    sm.ring_bytes = (uint32_t)(2 * NN_RB * dm.dh + t1) * 4u;
    sm.H = p; p += NN_RB * dm.dh;
    sm.dZ1 = p; p += NN_RB * dm.dh;
    sm.tile = p; p += t1;
    sm.dZ2 = p; p += NN_RB * dm.dop;
    sm.A = p; p += dm.dh * dm.dop;
    sm.At = p; p += dm.dop * dm.dh;
    sm.a = p; p += NN_MAXO;
    sm.b = p; p += NN_MAXH;
    sm.gb = p; p += NN_WARPS * dm.dh;
    sm.gA = p; p += dm.dh * dm.dop;
    sm.ga = p; p += NN_MAXO;
    sm.red = p; p += 64;
    sm.rows = reinterpret_cast<int*>(p); p += NN_RB;
    sm.tc_slot = reinterpret_cast<uint32_t*>(p); p += 4;
    sm.tc_bars = reinterpret_cast<uint64_t*>(p); p += 4;
    sm.s1_bars = reinterpret_cast<uint64_t*>(p);
}

static size_t nn_smem_bytes(const NNDims& dm) {
    const int t1 = nn_tile_floats_all(dm);
    size_t f = (size_t)2 * NN_RB * dm.dh + NN_RB * dm.dop + 2 * dm.dh * dm.dop + NN_MAXO + NN_MAXH + t1 +
               NN_WARPS * dm.dh + dm.dh * dm.dop + NN_MAXO + 64 + NN_RB + 4 + 4 + 32;
    return f * sizeof(float) + 1024;
}

}  // namespace sgm
