// Control-variate full-data gradient at MANY modes (one per chain) on the 5th-generation tensor cores.
//
// Replaces calcFullGrads / fullCalc (R/controlVariates.R:96-97,105-111) for the logistic family when every chain
// carries its own mode (the reference's independent runs each call getMode, cV.R:37-52).  With M modes the pass is
// two dense contractions over the N x d data matrix ("CV full gradient, C modes"):
//     Z[i][m] = x_i . beta_m + bias_m            [N x d] x [d x M]          (pass 1, epilogue: R = y - sigmoid(Z))
//     G[m][k] = sum_i R[i][m] X[i][k]            [d x N] x [N x M]          (pass 2, split over row ranges)
// 4 N d M flops against N d 4 bytes: tensor-bound from a few dozen modes on.
//
// Both passes are ONE kernel, C[128 x 128] = A[128 x K] B[128 x K]^T:
//     pass 1: A = X (rows i), B = theta_hat (rows m), K = features: both tiles arrive K-major (128-byte swizzle);
//     pass 2: A = X^T (rows k), B = R^T (rows m; written by pass 1), K = data rows.  No transposed copy of the data exists: the
//             TMA box is 32 data rows x 128 features of X as it lies in HBM, and the warps that split A into tensor memory read
//             it column-wise (lane = feature, one 4-byte load per K element: 128 contiguous bytes per warp load, conflict free) --
//             the A operand goes through registers on its way to TMEM anyway, so the transposition is free.  (Round 1 kept a
//             second, transposed copy of X in HBM for this pass: + 51 GB at the c5 shard, and a 16-byte alignment rule on the
//             K coordinate that an odd shard start broke.)
// float32 accuracy on TF32 tensor cores by the 3-term split hi*hi + hi*lo + lo*hi (tcgen05.mma kind::tf32, accumulators in
// TMEM).  Pipeline per 32-wide K stage: warp 0 issues two TMA boxes (128-byte swizzle) into a raw stage; warps 2..7 round the
// raw B tile to TF32 in place (hi), write the residual (lo) into a lo stage and split the raw A tile into TENSOR MEMORY
// (tcgen05.st; A-from-TMEM MMAs fetch only B from shared memory, whose bandwidth bounds the SS form); warp 1 issues 12 MMAs.  The tensor core's
// float32 accumulation truncates, so K is cut into segments of 8 stages (96 accumulation steps) that alternate between two
// TMEM buffers; warps 8..15 drain a finished segment into float32 registers (round-to-nearest adds) while the next one runs,
// and write the item's output (pass 1: R^T; pass 2: double-precision atomics into the gradient sums).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace sgm {

constexpr int CVG_THREADS = 512;
constexpr int CVG_WARPS = CVG_THREADS / 32;
constexpr int CVG_TILE = 128;          // C tile: 128 A rows x 128 B rows
constexpr int CVG_KC = 32;             // K per stage (one 128-byte swizzle span of float32)
constexpr int CVG_NRAW = 4;            // raw stages (TMA destinations)
constexpr int CVG_NLO = 2;             // lo stages
constexpr int CVG_SEG = 8;             // stages per TMEM segment
constexpr int CVG_CONV_WARPS = CVG_WARPS - 2;
constexpr int CVG_DRAIN_WARP0 = 8;     // warps 8..15 also drain TMEM (two per lane quarter, 64 columns each)
constexpr uint32_t CVG_OP_BYTES = CVG_TILE * 128u;          // one operand tile: 128 rows x 128 bytes
constexpr uint32_t CVG_STAGE_BYTES = 2u * CVG_OP_BYTES;     // A | B
constexpr int CVG_BCONV_WARPS = CVG_DRAIN_WARP0 - 2;   // warps 2..7 split the B operand in shared memory; warps 8..15 move A into TMEM
constexpr uint32_t CVG_TMEM_A = 256;                   // TMEM columns: [0, 256) two accumulator buffers, [256, 384) two A buffers (hi 32 | lo 32)
constexpr size_t CVG_SMEM = (size_t)CVG_NRAW * CVG_STAGE_BYTES + (size_t)CVG_NLO * CVG_OP_BYTES + 1024 /* alignment */ + 256 /* barriers, tmem slot */;

struct CvgArgs {
    int epi;               // 0: residual pass, 1: gradient pass
    int k_chunks;          // stages per CTA (K extent of this CTA / 32, rounded up)
    int tiles_a, tiles_b;  // C tiles along the A rows / the B rows (modes); a work item = (tile_a, tile_b, K split)
    int n_items;           // tiles_a * tiles_b * splits; CTA b takes items b, b + gridDim.x, ...
    int a_row0;            // row coordinate of A-tile 0 in the A tensor (pass 1: first row of the block)
    int a_k0, b_k0;        // K coordinate offsets of split 0 into the A and B tensors (pass 2: global row / block-relative row)
    int k_per_split;       // K elements per split (multiple of 32 x 8)
    int k_end_a;           // exclusive K bound of the A tensor coordinates (pass 2: row1)
    int a_box_w;           // pass 2: features per row of the A tile in shared memory (the TMA box is a_box_w features x 32 data rows, no swizzle)
    int bias_row;          // pass 2: A row (feature index) that is the constant 1 -- its sums are the bias component of the gradient, sum_i R[i][m]
                           // (d when d % 128 != 0: the row is free, the tile is not full; -1: the sums come from cvg_rowsum_kernel instead)
    // pass 1 epilogue
    const float* X; int ld; int d;
    int64_t row_base;      // global row of A-tile 0 (this launch's row block)
    int64_t row_end;       // rows >= row_end write r = 0
    const float* theta_hat; int Pst; int M;
    float* Rt; int64_t rt_pitch;   // R^T [M][rt_pitch], column = rt_col0 + row - row_base
    int rt_col0;           // column of the block's first row (0: the R^T coordinates are block-relative, so every box start is 16-byte aligned)
    double* gsum;          // [M][Pst]: pass 1 adds the bias component (column d), pass 2 the feature components
};

// K-major, 128-byte swizzle: 128-byte rows, 8-row groups 1024 bytes apart; a K step of 8 TF32 = +32 bytes in the start address
__device__ __forceinline__ uint64_t cvg_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ float cvg_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }   // round to nearest, ties away

// work item -> (A tile, B tile, K split); consecutive items share the A tile (and run on different CTAs at the same time: L2 reuse)
struct CvgItem { int tile_a, tile_b, ka0, kb0, nchunks; };
__device__ __forceinline__ CvgItem cvg_item(const CvgArgs& a, int item) {
    CvgItem w;
    const int tiles = a.tiles_a * a.tiles_b;
    const int split = item / tiles, t = item % tiles;
    w.tile_a = t / a.tiles_b; w.tile_b = t % a.tiles_b;
    w.ka0 = a.a_k0 + split * a.k_per_split; w.kb0 = a.b_k0 + split * a.k_per_split;
    w.nchunks = a.k_chunks;
    if (a.epi == 1) {   // the last split may be short
        const int left = a.k_end_a - w.ka0;
        const int c = left > 0 ? (left + CVG_KC - 1) / CVG_KC : 0;
        if (c < w.nchunks) w.nchunks = c;
    }
    return w;
}

__global__ void __launch_bounds__(CVG_THREADS, 1) cvg_kernel(const CvgArgs a, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    SGM_DYN_SMEM(unsigned char, cvg_smem, 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t base = (smem_u32(cvg_smem) + 1023u) & ~1023u;
    const uint32_t raw0 = base, lo0 = base + CVG_NRAW * CVG_STAGE_BYTES;   // lo stages hold only the B residuals
    const uint32_t bars = lo0 + CVG_NLO * CVG_OP_BYTES;
    // barriers: full[NRAW] | rawfree[NRAW] | conv[NLO] | lofree[NLO] | tfull[2] | tfree[2]; then the TMEM slot
    const uint32_t b_full = bars, b_rawfree = b_full + 8 * CVG_NRAW, b_conv = b_rawfree + 8 * CVG_NRAW, b_lofree = b_conv + 8 * CVG_NLO,
                   b_tfull = b_lofree + 8 * CVG_NLO, b_tfree = b_tfull + 16, slot = b_tfree + 16;

    if (warp == 0) {
        tmem_alloc<512>(slot);
        tmem_relinquish();
    }
    if (tid == 0) {
        for (int s = 0; s < CVG_NRAW; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_rawfree + 8 * s, 1); }
        for (int l = 0; l < CVG_NLO; ++l) { mbar_init(b_conv + 8 * l, CVG_BCONV_WARPS); mbar_init(b_lofree + 8 * l, 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(b_tfull + 8 * t, 1); mbar_init(b_tfree + 8 * t, CVG_WARPS - CVG_DRAIN_WARP0); }
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = lds_u32(slot);
    SGM_EMU_NOTE(a.epi == 0 ? "cvg:pass1" : "cvg:pass2");

    // every role walks the same item sequence; g = running stage counter, sg = running segment counter (barrier phases never reset)
    if (warp == 0) {
        // ---- producer
        if (lane == 0) {
            uint32_t g = 0;
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
                const CvgItem w = cvg_item(a, item);
                for (int c = 0; c < w.nchunks; ++c, ++g) {
                    const uint32_t s = g % CVG_NRAW;
                    if (g >= CVG_NRAW) mbar_wait(b_rawfree + 8 * s, ((g / CVG_NRAW) - 1u) & 1u);
                    const uint32_t dst = raw0 + s * CVG_STAGE_BYTES, bar = b_full + 8 * s;
                    if (a.epi == 0) {
                        mbar_expect_tx(bar, CVG_STAGE_BYTES);
                        tma_box2(dst, &tmA, w.ka0 + c * CVG_KC, a.a_row0 + w.tile_a * CVG_TILE, bar);
                    } else {   // 32 data rows (K) x a_box_w features of X, rows as they lie in HBM
                        mbar_expect_tx(bar, (uint32_t)a.a_box_w * CVG_KC * 4u + CVG_OP_BYTES);
                        tma_box2(dst, &tmA, w.tile_a * CVG_TILE, w.ka0 + c * CVG_KC, bar);
                    }
                    tma_box2(dst + CVG_OP_BYTES, &tmB, w.kb0 + c * CVG_KC, w.tile_b * CVG_TILE, bar);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CVG_TILE >> 3) << 17) | ((uint32_t)(CVG_TILE >> 4) << 24);
        uint32_t g = 0, sg = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            const CvgItem w = cvg_item(a, item);
            for (int c = 0; c < w.nchunks; ++c, ++g) {
                const uint32_t s = g % CVG_NRAW, l = g % CVG_NLO, tb = sg & 1u;
                const bool seg_first = (c % CVG_SEG) == 0, seg_last = (c % CVG_SEG) == CVG_SEG - 1 || c == w.nchunks - 1;
                if (seg_first && sg >= 2) mbar_wait(b_tfree + 8 * tb, ((sg >> 1) - 1u) & 1u);   // buffer drained
                mbar_wait(b_conv + 8 * l, (g / CVG_NLO) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t td = tmem + tb * CVG_TILE;
                    const uint64_t bhi = cvg_desc(raw0 + s * CVG_STAGE_BYTES + CVG_OP_BYTES), blo = cvg_desc(lo0 + l * CVG_OP_BYTES);
                    const uint32_t ahi = tmem + CVG_TMEM_A + l * 64u, alo = ahi + 32u;
#pragma unroll
                    for (int k8 = 0; k8 < CVG_KC / 8; ++k8) {
                        const uint64_t ko = (uint64_t)(2 * k8);
                        umma_tf32_ts(td, ahi + 8u * k8, bhi + ko, idesc, !(seg_first && k8 == 0));
                        umma_tf32_ts(td, ahi + 8u * k8, blo + ko, idesc, true);
                        umma_tf32_ts(td, alo + 8u * k8, bhi + ko, idesc, true);
                    }
                    umma_commit(b_rawfree + 8 * s);
                    umma_commit(b_lofree + 8 * l);
                    if (seg_last) umma_commit(b_tfull + 8 * tb);
                }
                __syncwarp();
                if (seg_last) ++sg;
            }
        }
    } else if (warp < CVG_DRAIN_WARP0) {
        // ---- converters (warps 2..7).  B operand: hi = rna_tf32(x) in place, lo = rna_tf32(x - hi) into the lo stage.
        // A operand: the swizzled raw tile -> hi / lo straight into TENSOR MEMORY (tcgen05.st), so the MMAs fetch only B from
        // shared memory.  A warp may touch only the TMEM lanes 32 (warp % 4) .. +31: quarters 2 and 3 have two converter warps
        // (16 K columns each), quarters 0 and 1 one (all 32 columns); the B slices are sized to even the load out (10 vs 12 float4).
        const int q = warp & 3;
        const bool heavy = (warp == 4 || warp == 5);              // both column halves of its quarter
        const int hb0 = heavy ? 0 : (warp >= 6 ? 1 : 0);
        const int nhalf = heavy ? 2 : 1;
        const int lw = heavy ? warp - 4 : (warp < 4 ? warp - 2 : warp - 4);   // light warps 2,3,6,7 -> 0..3; heavy 4,5 -> 0,1
        const int b_first = heavy ? 768 + lw * 32 + lane : lw * 32 + lane;
        const int b_step = heavy ? 64 : 128, b_rounds = heavy ? 4 : 6;
        const int r = 32 * q + lane;
        uint32_t g = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            const CvgItem w = cvg_item(a, item);
            for (int c = 0; c < w.nchunks; ++c, ++g) {
                const uint32_t s = g % CVG_NRAW, l = g % CVG_NLO;
                const uint32_t raw = raw0 + s * CVG_STAGE_BYTES, braw = raw + CVG_OP_BYTES, lo = lo0 + l * CVG_OP_BYTES;
                if (g >= CVG_NLO) mbar_wait(b_lofree + 8 * l, ((g / CVG_NLO) - 1u) & 1u);
                mbar_wait(b_full + 8 * s, (g / CVG_NRAW) & 1u);
                // all shared loads of this thread first (one latency per stage), then the splits and the stores
                float4 xb[6], xa[8];
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    if (u < b_rounds) xb[u] = lds128(braw + 16u * (uint32_t)(b_first + u * b_step));
                if (a.epi == 0) {
                    const uint32_t rowp = raw + (uint32_t)r * 128u;
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc)
                        if (cc < 4 * nhalf) {
                            const uint32_t chunk = (uint32_t)((4 * hb0 + cc) ^ (r & 7));
                            xa[cc] = lds128(rowp + 16u * chunk);
                        }
                } else {
                    // pass 2: the tile is [32 data rows][a_box_w features]; this thread's A row is feature r, its K elements one per data row
                    const bool have = r < a.a_box_w;
                    const bool ones = (w.tile_a * CVG_TILE + r) == a.bias_row;   // the bias "feature": 1 for every data row (rows past the range meet R = 0)
                    const uint32_t colp = raw + 4u * (uint32_t)r, pitch = 4u * (uint32_t)a.a_box_w;
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc)
                        if (cc < 4 * nhalf) {
                            const uint32_t k0 = (uint32_t)(4 * (4 * hb0 + cc));
                            xa[cc].x = have ? lds32(colp + (k0 + 0u) * pitch) : 0.f; xa[cc].y = have ? lds32(colp + (k0 + 1u) * pitch) : 0.f;
                            xa[cc].z = have ? lds32(colp + (k0 + 2u) * pitch) : 0.f; xa[cc].w = have ? lds32(colp + (k0 + 3u) * pitch) : 0.f;
                            if (ones) xa[cc] = make_float4(1.f, 1.f, 1.f, 1.f);
                        }
                }
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    if (u < b_rounds) {
                        const uint32_t off = 16u * (uint32_t)(b_first + u * b_step);
                        const float hx = cvg_tf32(xb[u].x), hy = cvg_tf32(xb[u].y), hz = cvg_tf32(xb[u].z), hw = cvg_tf32(xb[u].w);
                        const float lx = cvg_tf32(xb[u].x - hx), ly = cvg_tf32(xb[u].y - hy), lz = cvg_tf32(xb[u].z - hz), lq = cvg_tf32(xb[u].w - hw);
                        sts128(braw + off, make_float4(hx, hy, hz, hw));
                        sts128(lo + off, make_float4(lx, ly, lz, lq));
                    }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    if (hh < nhalf) {
                        float hi[16], lv[16];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const float4 x = xa[4 * hh + cc];
                            hi[4 * cc + 0] = cvg_tf32(x.x); hi[4 * cc + 1] = cvg_tf32(x.y); hi[4 * cc + 2] = cvg_tf32(x.z); hi[4 * cc + 3] = cvg_tf32(x.w);
                            lv[4 * cc + 0] = cvg_tf32(x.x - hi[4 * cc + 0]); lv[4 * cc + 1] = cvg_tf32(x.y - hi[4 * cc + 1]);
                            lv[4 * cc + 2] = cvg_tf32(x.z - hi[4 * cc + 2]); lv[4 * cc + 3] = cvg_tf32(x.w - hi[4 * cc + 3]);
                        }
                        const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + CVG_TMEM_A + l * 64u + 16u * (uint32_t)(hb0 + hh);
                        tmem_st16(ta, hi);
                        tmem_st16(ta + 32u, lv);
                    }
                tmem_wait_st();
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_conv + 8 * l);
            }
        }
    } else {
        // ---- drain + epilogue warps (8..15): finished K segments TMEM -> float32 registers, then the item's output; never on the
        // converters' path, so the next item's stages keep flowing during an epilogue
        const int q = warp & 3, hb = (warp - CVG_DRAIN_WARP0) >> 2;
        const uint32_t tq = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * hb);
        float acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.f;
        uint32_t sg = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            const CvgItem w = cvg_item(a, item);
            const int nseg = (w.nchunks + CVG_SEG - 1) / CVG_SEG;
            for (int sl = 0; sl < nseg; ++sl, ++sg) {
                const uint32_t tb = sg & 1u;
                mbar_wait(b_tfull + 8 * tb, (sg >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    uint32_t rr[16];
                    tmem_ld16(tq + tb * CVG_TILE + 16u * gq, rr);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[16 * gq + j] += __uint_as_float(rr[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_tfree + 8 * tb);
            }
            // ---- epilogue: this thread owns C row 32 q + lane, columns 64 hb .. 64 hb + 63 of the tile
            const int arow = w.tile_a * CVG_TILE + 32 * q + lane;
            const int bcol0 = w.tile_b * CVG_TILE + 64 * hb;
            if (a.epi == 0) {
                const int64_t row = a.row_base + arow;
                const bool live = row < a.row_end;
                const float y = live ? __ldg(a.X + (size_t)row * a.ld + a.d) : 0.f;
                const float bias_l = (bcol0 + lane < a.M) ? __ldg(a.theta_hat + (size_t)(bcol0 + lane) * a.Pst + a.d) : 0.f;
                const float bias_h = (bcol0 + 32 + lane < a.M) ? __ldg(a.theta_hat + (size_t)(bcol0 + 32 + lane) * a.Pst + a.d) : 0.f;
#pragma unroll
                for (int j = 0; j < 64; ++j) {   // fully unrolled: acc[] stays in registers
                    const int m = bcol0 + j;
                    const float bias = __shfl_sync(0xffffffffu, j < 32 ? bias_l : bias_h, j & 31);
                    if (m < a.M) {   // warp-uniform
                        const float rv = live ? y - sigmoidf(acc[j] + bias) : 0.f;
                        if (a.rt_col0 + arow < a.rt_pitch) a.Rt[(size_t)m * a.rt_pitch + a.rt_col0 + arow] = rv;
                    }
                    acc[j] = 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const int m = bcol0 + j;
                    if ((arow < a.d || arow == a.bias_row) && m < a.M) atomicAdd(a.gsum + (size_t)m * a.Pst + arow, (double)acc[j]);
                    acc[j] = 0.f;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// bias component of the gradient: gsum[m][d] += sum_i R^T[m][i]   (one warp per 4096-column slice of a row of R^T).  Only when d is a
// multiple of 128: otherwise pass 2 carries a constant-1 row in its last, partly empty feature tile and this read of R^T is saved.
__global__ void cvg_rowsum_kernel(const float* __restrict__ Rt, int64_t pitch, int64_t rows, int M, int Pst, int d, double* __restrict__ gsum) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t slices = (rows + 4095) / 4096;
    if (wid >= (int64_t)M * slices) return;
    const int m = (int)(wid / slices);
    const int64_t i0 = (wid % slices) * 4096, i1 = (i0 + 4096 < rows) ? i0 + 4096 : rows;
    float s = 0.f;
    for (int64_t i = i0 + lane; i < i1; i += 32) s += Rt[(size_t)m * pitch + i];
    s = warp_sum(s);
    if (lane == 0) atomicAdd(gsum + (size_t)m * Pst + d, (double)s);
}

}  // namespace sgm
