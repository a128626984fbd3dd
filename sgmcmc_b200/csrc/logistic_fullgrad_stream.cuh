// Control-variate full-data gradient at ONE mode, as a streaming kernel: persistent CTAs pull
// contiguous row tiles through a 3-stage TMA bulk-copy ring (one cp.async.bulk of ~32 KB per tile,
// mbarrier-tracked) and accumulate sum_i (y_i - s(x_i.beta_hat + bias_hat)) [x_i, 1] in registers across
// all of their tiles; one double-precision atomic per column at the very end.
//
// Replaces calcFullGrads / fullCalc (R/controlVariates.R:96-97,105-111) for the logistic family; HBM-bound
// (N * ld * 4 bytes read once).  Per-chain modes (M > 1) still use logistic_fullgrad_kernel.
#pragma once
#include "common.cuh"
#include "logistic.cuh"
#include "logistic_staged.cuh"

namespace sgm {

constexpr int FG_THREADS = 256;
constexpr int FG_WARPS = FG_THREADS / 32;
constexpr int FG_STAGES = 3;
constexpr int FG_SHORT_ROWS = 128;   // rows per tile, lane-per-row mapping
constexpr int FG_LONG_ROWS = 8;      // rows per tile, warp-per-row mapping

static inline size_t fg_smem_bytes(int ld, int tile_rows) {
    return sizeof(float) * ((size_t)FG_STAGES * tile_rows * ld + (size_t)ld + tile_rows + 64) + 8 * FG_STAGES + 16;
}

// V = float4 columns per lane in the long-row mapping (ignored when !LONG)
template <bool LONG, int V>
__global__ void __launch_bounds__(FG_THREADS, 2) logistic_fullgrad_stream_kernel(const float* __restrict__ X, int ld, int d, int64_t row0,
                                                                                 int64_t row1, int tile_rows, int n_tiles,
                                                                                 const float* __restrict__ theta_hat, double* __restrict__ gsum) {
    SGM_DYN_SMEM(unsigned char, smem_raw, 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Q = ld >> 2, qt = d >> 2, ct = d & 3;
    float* tiles = reinterpret_cast<float*>(smem_raw);
    float* th = tiles + (size_t)FG_STAGES * tile_rows * ld;   // beta_hat with the bias slot and padding zeroed
    float* coef = th + ld;
    float* scratch = coef + tile_rows;
    const uint32_t bars = (smem_u32(scratch + 64) + 7u) & ~7u;
    const uint32_t tiles_s = smem_u32(tiles), th_s = smem_u32(th), coef_s = smem_u32(coef);
    const uint32_t tile_bytes = (uint32_t)tile_rows * (uint32_t)ld * 4u;
    for (int e = tid; e < ld; e += FG_THREADS) th[e] = (e < d) ? theta_hat[e] : 0.f;
    const float bias = theta_hat[d];
    if (tid < FG_STAGES) mbar_init(bars + 8u * tid, 1);
    fence_mbar_init();
    __syncthreads();

    SGM_EMU_NOTE(LONG ? "fullgrad_stream:long" : "fullgrad_stream:short");
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto issue = [&](int k) {   // thread 0: tile k of this CTA into stage k % FG_STAGES
        const int64_t t = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
        const int64_t r_begin = row0 + t * tile_rows;
        int64_t rows = row1 - r_begin; if (rows > tile_rows) rows = tile_rows;
        const uint32_t bytes = (uint32_t)rows * (uint32_t)ld * 4u;
        const uint32_t bar = bars + 8u * (uint32_t)(k % FG_STAGES);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(tiles_s + (uint32_t)(k % FG_STAGES) * tile_bytes, X + (size_t)r_begin * ld, bytes, bar);
    };
    if (tid == 0) for (int k = 0; k < FG_STAGES && k < my_tiles; ++k) issue(k);

    const bool paired = !LONG && (Q % 4 == 2);
    const int skew = paired ? ((tid >> 2) & 1) : 0;
    float csum = 0.f;
    // short rows: one persistent float4 accumulator per (lane, column slot); long rows: V per lane
    float4 acc[LONG ? V : 4];
#pragma unroll
    for (int v = 0; v < (LONG ? V : 4); ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 thr[LONG ? V : 1];
    if (LONG) {
#pragma unroll
        for (int v = 0; v < V; ++v) thr[v] = (lane + 32 * v < Q) ? *reinterpret_cast<const float4*>(th + 4 * (lane + 32 * v)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (int k = 0; k < my_tiles; ++k) {
        const int64_t t = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
        const int64_t r_begin = row0 + t * tile_rows;
        int rows = (int)((row1 - r_begin) > tile_rows ? tile_rows : (row1 - r_begin));
        const uint32_t base = tiles_s + (uint32_t)(k % FG_STAGES) * tile_bytes;
        mbar_wait(bars + 8u * (uint32_t)(k % FG_STAGES), (uint32_t)(k / FG_STAGES) & 1u);
        if (LONG) {
            // warp-per-row: lanes stride the float4 columns of the row (conflict-free), beta_hat in registers
            for (int j = warp; j < rows; j += FG_WARPS) {
                const uint32_t xa = base + (uint32_t)(j * Q) * 16u;
                float4 x[LONG ? V : 1];
                float dz = 0.f, y = 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int q = lane + 32 * v;
                    x[v] = (q < Q) ? lds128(xa + q * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q == qt) y = get4(x[v], ct);
                    dz = fdot4(x[v], thr[v], dz);
                }
                dz = warp_sum(dz);
                y = __shfl_sync(0xffffffffu, y, qt & 31);
                const float cf = y - sigmoid_fast(bias + dz);
                csum += cf;   // every lane holds the same value; lane 0's copy is used
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    acc[v].x = fmaf(cf, x[v].x, acc[v].x); acc[v].y = fmaf(cf, x[v].y, acc[v].y);
                    acc[v].z = fmaf(cf, x[v].z, acc[v].z); acc[v].w = fmaf(cf, x[v].w, acc[v].w);
                }
            }
        } else {
            // phase 1: lane-per-row dots
            if (tid < tile_rows) {
                float cf = 0.f;
                if (tid < rows) {
                    const uint32_t xa = base + (uint32_t)(tid * Q) * 16u;
                    float dz0 = 0.f, dz1 = 0.f;
                    int kq = 0;
                    for (; kq + 1 < Q; kq += 2) {
                        const int q0 = kq + skew, q1 = (q0 + 1 == Q) ? 0 : q0 + 1;
                        const float4 x0 = lds128(xa + q0 * 16), x1 = lds128(xa + q1 * 16);
                        dz0 = fdot4(x0, lds128(th_s + q0 * 16), dz0);
                        dz1 = fdot4(x1, lds128(th_s + q1 * 16), dz1);
                    }
                    if (Q & 1) dz0 = fdot4(lds128(xa + (Q - 1) * 16), lds128(th_s + (Q - 1) * 16), dz0);
                    const float y = get4(lds128(xa + qt * 16), ct);
                    cf = y - sigmoid_fast(bias + (dz0 + dz1));
                    csum += cf;
                }
                sts32(coef_s + tid * 4, cf);
            }
            __syncthreads();
            // phase 2: column sums, accumulators persist across tiles
            const int rstep = paired ? 16 : 32, rl = paired ? (lane >> 1) : lane, cpar = paired ? (lane & 1) : 0, cstep = paired ? 2 : 1;
            int slot = 0;
            for (int q0 = cstep * warp; q0 < Q; q0 += cstep * FG_WARPS, ++slot) {
                const uint32_t col = base + (uint32_t)(q0 + cpar) * 16u;
                float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int j = rl; j < rows; j += rstep) {
                    const float4 x = lds128(col + (uint32_t)(j * Q) * 16u);
                    const float cf = lds32(coef_s + j * 4);
                    a4.x = fmaf(cf, x.x, a4.x); a4.y = fmaf(cf, x.y, a4.y); a4.z = fmaf(cf, x.z, a4.z); a4.w = fmaf(cf, x.w, a4.w);
                }
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2)
                    if (s2 == slot) { acc[s2].x += a4.x; acc[s2].y += a4.y; acc[s2].z += a4.z; acc[s2].w += a4.w; }
            }
        }
        __syncthreads();   // the stage is free again
        if (tid == 0 && k + FG_STAGES < my_tiles) {
            fence_proxy_async();
            issue(k + FG_STAGES);
        }
    }

    // ---- flush: lanes -> one value per column -> double atomics
    if (LONG) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int q = lane + 32 * v;
            if (q < Q) {
                const float vals[4] = {acc[v].x, acc[v].y, acc[v].z, acc[v].w};
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) { const int e = 4 * q + c4; if (e < d) atomicAdd(gsum + e, (double)vals[c4]); }
            }
        }
        if (lane == 0) atomicAdd(gsum + d, (double)csum);
    } else {
        const int cstep = paired ? 2 : 1, cpar = paired ? (lane & 1) : 0;
        int slot = 0;
        for (int q0 = cstep * warp; q0 < Q; q0 += cstep * FG_WARPS, ++slot) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) if (s2 == slot) t4 = acc[s2];
#pragma unroll
            for (int o = cstep; o < 32; o <<= 1) {
                t4.x += __shfl_xor_sync(0xffffffffu, t4.x, o); t4.y += __shfl_xor_sync(0xffffffffu, t4.y, o);
                t4.z += __shfl_xor_sync(0xffffffffu, t4.z, o); t4.w += __shfl_xor_sync(0xffffffffu, t4.w, o);
            }
            if (lane < cstep) {
                const float vals[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) { const int e = 4 * (q0 + cpar) + c4; if (e < d) atomicAdd(gsum + e, (double)vals[c4]); }
            }
        }
        csum = warp_sum(csum);
        if (lane == 0) atomicAdd(gsum + d, (double)csum);
    }
}

}  // namespace sgm
