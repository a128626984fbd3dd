// 784-100-10 softmax-softmax network family (vignettes/nn.Rmd:58-79): one CTA per chain,
// one persistent launch for `iters` sgmcmcSteps.
//
// Replaces, for this family, dataFeed/dataSelect (R/update.R:166-194), tf$gradients of
// estLogPost (R/setup.R:36-62,116-125), declareDynamics.* (R/dynamics.R:7-120) and the
// per-tensor Session.run loops of sgmcmcStep.* (R/update.R:40-81).
//
// Forward / backward, per minibatch row block (<= 128 rows):
//   S1  Z1 = X B + b,  H = softmax_row(Z1)                       [n x 784] x [784 x 100]
//   S2  Z2 = H A + a,  Pr = softmax_row(Z2),  dZ2 = Y - Pr * rowsum(Y)
//   S3  grad A = H^T dZ2,  grad a = colsum(dZ2)
//   S4  dH = dZ2 A^T,  dZ1 = H * (dH - rowsum(dH * H)),  grad b = colsum(dZ1)
//   S5  grad B = X^T dZ1                                          [784 x n] x [n x 100]
// The reference updates the 8 parameter tensors one Session.run at a time (Gauss-Seidel,
// R/update.R:43-45,60-64): A, B, a, b, lambdaA, lambdaB, lambdaa, lambdab.  Because Z1/H
// depend only on (B, b), the exact sequence needs S1 twice and S5 once per sub-step:
//   S1 S2 S3 -> A;  S2 S4 S5 -> B;  S1 S2 S3 -> a;  S2 S4 -> b;  lambdas (prior only).
//
// Internal parameter layout per chain (float32, stride Pst): A[k][l] row-major, B[j][k]
// row-major, a, b, 4 lambdas; the flat (ABI) order keeps R's column-major element order
// inside A and B.  Dynamics noise is indexed by INTERNAL element (Philox block = 4
// consecutive internal elements) when drawn on the device, by flat index when supplied.
#include "softmax_net.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>   // CUtensorMap

namespace sgm {

constexpr int NN_THREADS = 512;
constexpr int NN_WARPS = NN_THREADS / 32;
constexpr int NN_RB = 128;      // rows per block
constexpr int NN_MAXH = 128;    // hidden width limit (lanes x 4)
constexpr int NN_MAXO = 16;     // output width limit

// optional phase timing (debug, ChainArgs::prof): thread 0 of every CTA adds clock64() deltas per stage
__device__ unsigned long long* nn_prof = nullptr;
__device__ int nn_stream_updates = 1;   // 0: register-staged update of every tensor (SGM_NN_UPDATE=regs)
__device__ int nn_s5_tmem = 0;          // 1: S5 with the X tile on the tensor-memory side (SGM_NN_S5=tmem; NOT YET RUN ON A GPU)
__device__ long long nn_t_prev[1024];
#define NN_TICK(slot)                                                                                   \
    if (nn_prof && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(nn_prof + (slot), (unsigned long long)(t_ - nn_t_prev[blockIdx.x & 1023])); nn_t_prev[blockIdx.x & 1023] = t_; }

struct NNDims {
    int din, dh, dop;  // dop = output width padded to a multiple of 4
    int dout;
    int offA, offB, offa, offb, offl;  // internal == flat tensor offsets
    int P;
};

struct SoftmaxNetWork {
    NNDims dm;
    int n, chains, device;
    float* G = nullptr;      // [chains][Pst] gradient sums at theta
    float* Ghat = nullptr;   // [chains][Pst] gradient sums at theta_hat (cv)
    float* part = nullptr;   // full-gradient partials
    size_t part_cap = 0;
    size_t smem = 0;
    // TMA descriptors, encoded on first use and cached by (base pointer, extent, stride)
    struct MapEntry { const void* base; long long rows; long long stride; CUtensorMap map; };
    MapEntry maps[8];
    int n_maps = 0;
    // kernel selectors (softmax_net_set_option); 0 = default
    int opt_update = 0;   // 1: register-staged update of every tensor
    int opt_stage1 = 0;   // 1: register-staged forward-GEMM operands (no TMA)
    int opt_stage5 = 0;   // 1: S5 with the X tile on the tensor-memory side (only in builds with -DNN_ENABLE_S5_TMEM)
};

// TMA descriptors of one gradient evaluation: the dataset (rows gathered four at a time) and the parameter array that
// holds this evaluation's B (3-D: input, hidden unit, chain); x == nullptr selects the register-staged path
struct NNMaps {
    const CUtensorMap* x;
    const CUtensorMap* b;
    int bcoord;
};

// internal element order == flat (ABI) order: every tensor in R's column-major element order (A[k + dh*l], B[j + din*n])
__device__ __forceinline__ int nn_flat_of_internal(const NNDims&, int e) { return e; }
__device__ __forceinline__ int nn_tensor_of(const NNDims& dm, int e) {
    if (e < dm.offB) return 0;
    if (e < dm.offa) return 1;
    if (e < dm.offb) return 2;
    if (e < dm.offl) return 3;
    return 4 + (e - dm.offl);
}

// tcgen05 state of one CTA: TMEM base, two stage barriers (operand buffers) and their commit counts
struct TcCtx {
    uint32_t tmem;      // TMEM base address (512 columns allocated for the kernel's lifetime)
    uint32_t bar[2];    // shared-space addresses of the stage mbarriers
    uint32_t commits[2];
};

struct NNSmem {
    float* H;     // [NN_RB][dh]
    float* dZ1;   // [NN_RB][dh]
    float* dZ2;   // [NN_RB][dop]
    float* A;     // [dh][dop] (zero padded)
    float* At;    // [dop][dh] transposed copy (stage 4 reads it with conflict-free float4 loads)
    float* a;     // [dop]
    float* b;     // [dh]
    float* tile;  // max(S1 tiles, S5 tile)
    uint32_t ring;      // shared-space address (1024-byte aligned) of the H | dZ1 | tile region = the S1 operand ring
    uint32_t ring_bytes;
    uint64_t* s1_bars;  // [16] barriers of the S1 pipeline
    float* gb;    // [NN_WARPS][dh] partial column sums
    float* gA;    // [dh*dop] gradient of A (row block accumulation)
    float* ga;    // [dop]
    float* red;   // [64]
    int* rows;    // [NN_RB]
    uint32_t* tc_slot;  // [4]: TMEM base written by tcgen05.alloc; 16-byte aligned
    uint64_t* tc_bars;  // [2]
};

__device__ __forceinline__ float nn_block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = (threadIdx.x < NN_WARPS) ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) { t = warp_sum(t); if (threadIdx.x == 0) red[32] = t; }
    __syncthreads();
    return red[32];
}

// load A (padded), a, b of one parameter vector into shared memory; `which` selects the tensors that changed since the last
// load (bit 0: A, bit 1: a, bit 2: b) -- inside a Gauss-Seidel sub-step only one small tensor moves between evaluations
enum : int { LOAD_A = 1, LOAD_a = 2, LOAD_b = 4, LOAD_ALL = 7 };
__device__ void nn_load_small(const NNDims& dm, const NNSmem& sm, const float* __restrict__ th, int which = LOAD_ALL) {
    if (which & LOAD_A)
        for (int i = threadIdx.x; i < dm.dh * dm.dop; i += NN_THREADS) {
            const int l = i / dm.dh, k = i % dm.dh;   // coalesced over k (A is stored [k + dh * l])
            const float av = (l < dm.dout) ? th[dm.offA + k + dm.dh * l] : 0.f;
            sm.A[k * dm.dop + l] = av;
            sm.At[l * dm.dh + k] = av;
        }
    if (which & LOAD_a)
        for (int i = threadIdx.x; i < dm.dop; i += NN_THREADS) sm.a[i] = (i < dm.dout) ? th[dm.offa + i] : 0.f;
    if (which & LOAD_b)
        for (int i = threadIdx.x; i < dm.dh; i += NN_THREADS) sm.b[i] = th[dm.offb + i];
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05 helpers (5th-generation tensor cores: operands in shared memory, accumulator in tensor memory)
// ---------------------------------------------------------------------------------------------------------
// Shared-memory operand descriptor, K-major, no swizzle: the operand is a grid of 8-row x 16-byte core matrices;
// the 16-byte chunk (kg, row) lives at kg * lbo + row * 16, so LBO = distance between the two K chunks of one MMA,
// SBO = distance between 8-row groups = 128 B.  (Bit layout: cute/arch/mma_sm100_desc.hpp, descriptor version 1.)
__device__ __forceinline__ uint64_t nn_umma_desc(uint32_t addr, uint32_t lbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
// Instruction descriptor for kind::tf32, float32 accumulate, both operands K-major
__device__ __forceinline__ uint32_t nn_umma_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void nn_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// float32 -> (hi, lo), both rounded to nearest TF32 (cvt.rna): hi*hi' + hi*lo' + lo*hi' recovers float32-level
// products with unbiased errors (a truncating split leaves same-signed residuals that add up coherently over K)
__device__ __forceinline__ float nn_tf32(float x) {
    // round to nearest (ties away) on the 13 dropped mantissa bits == cvt.rna.tf32.f32, but on the integer pipe
    // (the conversion instruction issues at a fraction of the ALU rate and the split runs once per staged element)
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void nn_split4(const float4 x, float4& hi, float4& lo) {
    hi.x = nn_tf32(x.x); hi.y = nn_tf32(x.y); hi.z = nn_tf32(x.z); hi.w = nn_tf32(x.w);
    lo.x = nn_tf32(x.x - hi.x); lo.y = nn_tf32(x.y - hi.y); lo.z = nn_tf32(x.z - hi.z); lo.w = nn_tf32(x.w - hi.w);
}

// kernel prologue / epilogue: TMEM allocation (all 512 columns: one CTA per SM), stage barriers
__device__ void nn_tc_init(const NNSmem& sm, TcCtx& tc) {
    if ((threadIdx.x >> 5) == 0) {
        tmem_alloc<512>(smem_u32(sm.tc_slot));
        tmem_relinquish();
    }
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(sm.tc_bars), 1);
        mbar_init(smem_u32(sm.tc_bars + 1), 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tc.tmem = *sm.tc_slot;
    tc.bar[0] = smem_u32(sm.tc_bars); tc.bar[1] = smem_u32(sm.tc_bars + 1);
    tc.commits[0] = tc.commits[1] = 0;
}
__device__ void nn_tc_fini(const TcCtx& tc) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc<512>(tc.tmem);
}

// S1 epilogue, shared by the register-staged and the TMA-fed pipelines: TMEM -> registers (K segments summed), + bias -> sm.H,
// then the row softmax in place
__device__ void nn_stage1_epilogue(const NNDims& dm, const NNSmem& sm, int nb, const TcCtx& tc, int nseg, int Npad) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // epilogue, all 16 warps: warp w reads TMEM lanes 32 (w % 4) .. +31 (= rows) for the 16-column groups w / 4, w / 4 + 4, ...,
    // sums the K segments in registers, adds the bias and parks Z1 in sm.H; then one warp per row does the softmax in place
    {
        const int q = warp & 3, row = 32 * q + lane;
        const uint32_t tbase = tc.tmem + ((uint32_t)(32 * q) << 16);
        for (int c0 = 16 * (warp >> 2); c0 < Npad; c0 += 64) {
            float v[16];
            nn_tmem_ld16(tbase + (uint32_t)c0, v);
            for (int sg = 1; sg < nseg; ++sg) {
                float w[16];
                nn_tmem_ld16(tbase + (uint32_t)(sg * Npad + c0), w);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += w[i];
            }
            if (row < nb) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    if (c0 + i < dm.dh)
                        *reinterpret_cast<float4*>(sm.H + row * dm.dh + c0 + i) =
                            make_float4(v[i] + sm.b[c0 + i], v[i + 1] + sm.b[c0 + i + 1], v[i + 2] + sm.b[c0 + i + 2], v[i + 3] + sm.b[c0 + i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    {
        const int ncg = dm.dh >> 2;
        const bool live = lane < ncg;
        for (int row = warp; row < nb; row += NN_WARPS) {
            float4 z = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (live) z = *reinterpret_cast<const float4*>(sm.H + row * dm.dh + 4 * lane);
            float m = fmaxf(fmaxf(z.x, z.y), fmaxf(z.z, z.w));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) { e.x = expf(z.x - m); e.y = expf(z.y - m); e.z = expf(z.z - m); e.w = expf(z.w - m); }
            const float inv = 1.0f / warp_sum(e.x + e.y + e.z + e.w);
            if (live) *reinterpret_cast<float4*>(sm.H + row * dm.dh + 4 * lane) = make_float4(e.x * inv, e.y * inv, e.z * inv, e.w * inv);
        }
    }
    __syncthreads();
}

constexpr int TC_KC = 16;   // K per operand stage of S1 (two k8 MMAs x three split terms)
constexpr int TC_SEG = 4;   // K is cut into TC_SEG segments with separate TMEM accumulators, summed in registers: the tensor
                            // core's float32 accumulation truncates, and the bias grows with the number of accumulation steps

// S1 on the tensor cores: Z1[128 x Npad] = X[128 x din] * B^T (B stored [hidden][input] = K-major) as 3xTF32 tcgen05.mma
// with the accumulator in TMEM; operand stages are written by all threads (float4 global loads -> hi/lo split -> 16-byte
// shared stores in the core-matrix layout), double-buffered against the asynchronous MMAs; the epilogue (bias, row
// softmax) reads TMEM with tcgen05.ld, one row per thread, and writes H to shared memory.
__device__ void nn_stage1_tc(const NNDims& dm, const NNSmem& sm, const float* __restrict__ X, int ld, int nb,
                             const float* __restrict__ th, TcCtx& tc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Npad = (dm.dh + 15) & ~15;
    const int LBOA = NN_RB * 16;                   // bytes between K chunks of the A operand (128 rows x 16 B)
    const int rowsB = Npad + 2;                    // +2 rows: K chunks land 8 banks apart (conflict-free 16-byte stores)
    const int LBOB = rowsB * 16;
    const int KG = TC_KC / 4;                      // 16-byte K chunks per stage
    const int A_FLOATS = KG * NN_RB * 4, B_FLOATS = KG * rowsB * 4, STAGE = 2 * A_FLOATS + 2 * B_FLOATS;
    const float* __restrict__ B = th + dm.offB;    // [dh][din]
    const uint32_t idesc = nn_umma_idesc(NN_RB, Npad);
    // zero the padding rows of the B operand (hidden units dh..Npad-1) in both stages once per call
    __syncthreads();
    for (int t = tid; t < 2 * 2 * KG * (rowsB - dm.dh); t += NN_THREADS) {
        const int st = t / (2 * KG * (rowsB - dm.dh)), r0 = t % (2 * KG * (rowsB - dm.dh));
        const int half = r0 / (KG * (rowsB - dm.dh)), r1 = r0 % (KG * (rowsB - dm.dh));
        const int kg = r1 / (rowsB - dm.dh), n = dm.dh + r1 % (rowsB - dm.dh);
        *reinterpret_cast<float4*>(sm.tile + st * STAGE + 2 * A_FLOATS + half * B_FLOATS + (kg * rowsB + n) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int xr = tid % NN_RB, xkg = tid / NN_RB;   // 512 threads = 128 rows x 4 chunks: one float4 of X per thread per stage
    const float* xrow = X + (size_t)sm.rows[xr < nb ? xr : 0] * ld;
    const int nchunks = (dm.din + TC_KC - 1) / TC_KC;
    const int seg_chunks = (nchunks + TC_SEG - 1) / TC_SEG;
    const int nseg = (nchunks + seg_chunks - 1) / seg_chunks;
    // software pipeline: the global loads of chunk c + 1 are issued before chunk c's barrier / MMAs
    const int bt0 = tid, bt1 = tid + NN_THREADS;                 // this thread's (up to two) B float4 slots per stage
    const bool b0_live = bt0 < dm.dh * KG, b1_live = bt1 < dm.dh * KG;
    const float* b0_ptr = B + (size_t)(bt0 / KG) * dm.din + 4 * (bt0 % KG);
    const float* b1_ptr = B + (size_t)(bt1 / KG) * dm.din + 4 * (bt1 % KG);
    // register pipeline three chunks deep: p0 is the chunk about to be staged, p1 / p2 the following two
    struct Pre { float4 x, b0, b1; };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](int cc) -> Pre {
        Pre pr{z4, z4, z4};
        if (cc < nchunks) {
            const int j0 = cc * TC_KC;
            if (xkg < KG && xr < nb && j0 + 4 * xkg < dm.din) pr.x = __ldg(reinterpret_cast<const float4*>(xrow + j0 + 4 * xkg));
            if (b0_live && j0 + 4 * (bt0 % KG) < dm.din) pr.b0 = *reinterpret_cast<const float4*>(b0_ptr + j0);
            if (b1_live && j0 + 4 * (bt1 % KG) < dm.din) pr.b1 = *reinterpret_cast<const float4*>(b1_ptr + j0);
        }
        return pr;
    };
    Pre p0 = fetch(0), p1 = fetch(1), p2 = fetch(2);
    // shared-memory descriptors of stage 0 (stage 1 = + STAGE bytes in the address field, which counts 16-byte units)
    uint64_t d_ahi[TC_KC / 8], d_alo[TC_KC / 8], d_bhi[TC_KC / 8], d_blo[TC_KC / 8];
    {
        const uint32_t a_hi = smem_u32(sm.tile), a_lo = a_hi + A_FLOATS * 4, b_hi = a_hi + 2 * A_FLOATS * 4, b_lo = b_hi + B_FLOATS * 4;
#pragma unroll
        for (int k8 = 0; k8 < TC_KC / 8; ++k8) {
            d_ahi[k8] = nn_umma_desc(a_hi + (uint32_t)(2 * k8) * LBOA, LBOA); d_alo[k8] = nn_umma_desc(a_lo + (uint32_t)(2 * k8) * LBOA, LBOA);
            d_bhi[k8] = nn_umma_desc(b_hi + (uint32_t)(2 * k8) * LBOB, LBOB); d_blo[k8] = nn_umma_desc(b_lo + (uint32_t)(2 * k8) * LBOB, LBOB);
        }
    }
    const uint32_t stage_step = (uint32_t)(STAGE * 4) >> 4;
    auto do_chunk = [&](int c, Pre& pr) {
        const int buf = c & 1;
        float* stage = sm.tile + buf * STAGE;
        // the MMAs that read this buffer two chunks ago must have completed
        if (tc.commits[buf] > 0) mbar_wait(tc.bar[buf], (tc.commits[buf] - 1u) & 1u);
        {
            float4 hi, lo;
            const float4 xv = pr.x, bv0 = pr.b0, bv1 = pr.b1;
            if (xkg < KG) {   // A operand: X rows
                nn_split4(xv, hi, lo);
                *reinterpret_cast<float4*>(stage + (xkg * NN_RB + xr) * 4) = hi;
                *reinterpret_cast<float4*>(stage + A_FLOATS + (xkg * NN_RB + xr) * 4) = lo;
            }
            if (b0_live) {    // B operand: hidden unit n, chunk kg
                nn_split4(bv0, hi, lo);
                *reinterpret_cast<float4*>(stage + 2 * A_FLOATS + ((bt0 % KG) * rowsB + bt0 / KG) * 4) = hi;
                *reinterpret_cast<float4*>(stage + 2 * A_FLOATS + B_FLOATS + ((bt0 % KG) * rowsB + bt0 / KG) * 4) = lo;
            }
            if (b1_live) {
                nn_split4(bv1, hi, lo);
                *reinterpret_cast<float4*>(stage + 2 * A_FLOATS + ((bt1 % KG) * rowsB + bt1 / KG) * 4) = hi;
                *reinterpret_cast<float4*>(stage + 2 * A_FLOATS + B_FLOATS + ((bt1 % KG) * rowsB + bt1 / KG) * 4) = lo;
            }
        }
        pr = fetch(c + 3);   // refill the SAME registers (no rotation moves: a move of an in-flight load result would stall on it)
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            const int seg = c / seg_chunks;
            const uint32_t td = tc.tmem + (uint32_t)(seg * Npad);
            const bool fresh = (c % seg_chunks) == 0;
#pragma unroll
            for (int k8 = 0; k8 < TC_KC / 8; ++k8) {
                // descriptors were built once before the loop: only the buffer's base differs (added to the 14-bit address field)
                const uint64_t bo = (uint64_t)(buf ? stage_step : 0u);
                umma_tf32_ss(td, d_ahi[k8] + bo, d_bhi[k8] + bo, idesc, !(fresh && k8 == 0));
                umma_tf32_ss(td, d_ahi[k8] + bo, d_blo[k8] + bo, idesc, true);
                umma_tf32_ss(td, d_alo[k8] + bo, d_bhi[k8] + bo, idesc, true);
            }
            umma_commit(tc.bar[buf]);
        }
        tc.commits[buf] += 1;
    };
    for (int c = 0; c < nchunks; c += 3) {
        do_chunk(c, p0);
        if (c + 1 < nchunks) do_chunk(c + 1, p1);
        if (c + 2 < nchunks) do_chunk(c + 2, p2);
    }
    NN_TICK(8)    // S1: staging loop
    // all MMAs complete in order: wait for the last commit of each buffer
#pragma unroll
    for (int b2 = 0; b2 < 2; ++b2)
        if (tc.commits[b2] > 0) mbar_wait(tc.bar[b2], (tc.commits[b2] - 1u) & 1u);
    tc_fence_after();
    NN_TICK(9)    // S1: drain
    nn_stage1_epilogue(dm, sm, nb, tc, nseg, Npad);
}

// ---------------------------------------------------------------------------------------------------------
// S1, TMA-fed: no operand byte passes through a register before it is in shared memory.
//   warp 14 (producer): per K chunk of 32 inputs, TMA tile::gather4 requests bring the minibatch rows' 128-byte segments (four rows
//            per request) and one 3-D box brings B[0..dh)[j0..j0+32) of this chain, both written with the 128-byte swizzle the
//            tensor core reads (K-major SW128), completion counted on the stage's mbarrier;
//   warps 0..13 (converters): round the raw stage to TF32 in place (= hi operands) and write the residuals to a lo stage of the
//            same layout (element-wise, so the swizzle never has to be computed), then fence to the async proxy;
//   warp 15 (issuer): 4 x (hi*hi + hi*lo + lo*hi) tcgen05.mma per stage, commits release the raw and the lo stage.
// The register-staged version of this loop paid one global-memory latency per chunk: ptxas drains every outstanding load at
// fence.proxy.async (profiles/r01_c4_s1_scoreboard.txt).
// ---------------------------------------------------------------------------------------------------------
// K-major operand descriptor, 128-byte swizzle: rows are 128 bytes, 8-row groups 1024 bytes apart (SBO); a K step of 8 TF32
// advances the start address by 32 bytes inside the swizzle span
__device__ __forceinline__ uint64_t nn_umma_desc_sw128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// S1 operand ring (TMA path): raw stages (X 128 rows x 128 B | B round8(dh) rows x 128 B, 128-byte swizzle) and lo stages of the
// same shape; the tensor core reads Npad = round16(dh) rows of B, i.e. up to 1 KB past a stage (garbage in, ignored columns out)
constexpr int S1_KC = 32;       // floats per operand row per stage = one 128-byte swizzle span
constexpr int S1_NRAW = 4;      // raw stages in flight (TMA destinations; rounded to TF32 in place = the hi operands)
constexpr int S1_NLO = 2;       // lo stages
__host__ __device__ inline int nn_s1_stage_bytes(const NNDims& dm) { return NN_RB * 128 + ((dm.dh + 7) / 8) * 1024; }
__host__ __device__ inline int nn_s1_ring_bytes(const NNDims& dm) { return (S1_NRAW + S1_NLO) * nn_s1_stage_bytes(dm) + 2048; }

constexpr int S1_CONV_WARPS = NN_WARPS - 2;

__device__ void nn_stage1_tma(const NNDims& dm, const NNSmem& sm, const NNMaps& mp, int nb, TcCtx& tc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Npad = (dm.dh + 15) & ~15;
    const uint32_t XB = NN_RB * 128u, BB = (uint32_t)((dm.dh + 7) / 8) * 1024u, STG = XB + BB;
    const int nchunks = (dm.din + S1_KC - 1) / S1_KC;
    const int seg_chunks = (nchunks + TC_SEG - 1) / TC_SEG;
    const int nseg = (nchunks + seg_chunks - 1) / seg_chunks;
    SGM_EMU_NOTE("nn_stage1_tma");
    const uint32_t raw0 = sm.ring, lo0 = sm.ring + S1_NRAW * STG;
    // barriers: full[S1_NRAW] | rawfree[S1_NRAW] | conv[S1_NLO] | lofree[S1_NLO] | done
    const uint32_t bars = smem_u32(sm.s1_bars);
    const uint32_t b_full = bars, b_rawfree = bars + 8 * S1_NRAW, b_conv = bars + 16 * S1_NRAW, b_lofree = b_conv + 8 * S1_NLO, b_done = b_lofree + 8 * S1_NLO;
    fence_proxy_async_global();   // this CTA's parameter updates (generic proxy) -> the TMA reads below
    __syncthreads();   // the previous users of H | dZ1 | tile are done
    if (tid == 0) {
        for (int s = 0; s < S1_NRAW; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_rawfree + 8 * s, 1); }
        for (int l = 0; l < S1_NLO; ++l) { mbar_init(b_conv + 8 * l, S1_CONV_WARPS); mbar_init(b_lofree + 8 * l, 1); }
        mbar_init(b_done, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == NN_WARPS - 2) {
        // ---- producer
        const int ngroups = (nb + 3) >> 2;
        const uint32_t tx = (uint32_t)ngroups * 512u + (uint32_t)dm.dh * 128u;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % S1_NRAW;
            if (c >= S1_NRAW) mbar_wait(b_rawfree + 8 * s, (uint32_t)((c / S1_NRAW) - 1) & 1u);
            const uint32_t dst = raw0 + (uint32_t)s * STG, bar = b_full + 8 * s;
            if (lane == 0) mbar_expect_tx(bar, tx);
            __syncwarp();
            for (int g = lane; g < ngroups; g += 32) {
                const int r = 4 * g;
                tma_gather4(dst + (uint32_t)g * 512u, mp.x, c * S1_KC, sm.rows[r], sm.rows[r + 1 < nb ? r + 1 : 0], sm.rows[r + 2 < nb ? r + 2 : 0],
                               sm.rows[r + 3 < nb ? r + 3 : 0], bar);
            }
            if (lane == 31) tma_box3(dst + XB, mp.b, c * S1_KC, 0, mp.bcoord, bar);
        }
    } else if (warp == NN_WARPS - 1) {
        // ---- MMA issuer
        const uint32_t idesc = nn_umma_idesc(NN_RB, Npad);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % S1_NRAW, l = c % S1_NLO;
            mbar_wait(b_conv + 8 * l, (uint32_t)(c / S1_NLO) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t td = tc.tmem + (uint32_t)((c / seg_chunks) * Npad);
                const bool fresh = (c % seg_chunks) == 0;
                const uint64_t ahi = nn_umma_desc_sw128(raw0 + (uint32_t)s * STG), bhi = nn_umma_desc_sw128(raw0 + (uint32_t)s * STG + XB);
                const uint64_t alo = nn_umma_desc_sw128(lo0 + (uint32_t)l * STG), blo = nn_umma_desc_sw128(lo0 + (uint32_t)l * STG + XB);
#pragma unroll
                for (int k8 = 0; k8 < S1_KC / 8; ++k8) {
                    const uint64_t ko = (uint64_t)(2 * k8);   // 32 bytes per K step, in 16-byte units
                    umma_tf32_ss(td, ahi + ko, bhi + ko, idesc, !(fresh && k8 == 0));
                    umma_tf32_ss(td, ahi + ko, blo + ko, idesc, true);
                    umma_tf32_ss(td, alo + ko, bhi + ko, idesc, true);
                }
                umma_commit(b_rawfree + 8 * s);
                umma_commit(b_lofree + 8 * l);
                if (c == nchunks - 1) umma_commit(b_done);
            }
            __syncwarp();
        }
    } else {
        // ---- converters: hi = rna_tf32(x) in place, lo = rna_tf32(x - hi) at the same offset of the lo stage
        const int nvec = (int)(STG >> 4);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % S1_NRAW, l = c % S1_NLO;
            const uint32_t raw = raw0 + (uint32_t)s * STG, lo = lo0 + (uint32_t)l * STG;
            const bool prof = nn_prof && tid == 0;   // sub-counters of the S1 slot: wait for the lo stage (MMA pace) / for the loads
            long long t0 = prof ? clock64() : 0;
            if (c >= S1_NLO) mbar_wait(b_lofree + 8 * l, (uint32_t)((c / S1_NLO) - 1) & 1u);
            if (prof) { const long long t1 = clock64(); atomicAdd(nn_prof + 9, (unsigned long long)(t1 - t0)); t0 = t1; }
            mbar_wait(b_full + 8 * s, (uint32_t)(c / S1_NRAW) & 1u);
            if (prof) atomicAdd(nn_prof + 8, (unsigned long long)(clock64() - t0));
            for (int i = tid; i < nvec; i += 32 * S1_CONV_WARPS) {
                const float4 x = lds128(raw + 16u * (uint32_t)i);
                float4 hi, lw;
                nn_split4(x, hi, lw);
                sts128(raw + 16u * (uint32_t)i, hi);
                sts128(lo + 16u * (uint32_t)i, lw);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_conv + 8 * l);
        }
    }
    mbar_wait(b_done, 0u);
    tc_fence_after();
    nn_stage1_epilogue(dm, sm, nb, tc, nseg, Npad);
}

// S2: dZ2[r][l] = Y[r][l] - Pr[r][l] * rowsum(Y[r]);  returns (via loglik) sum_r sum_l Y log Pr when asked
__device__ float nn_stage2(const NNDims& dm, const NNSmem& sm, const float* __restrict__ X, int ld, int nb, bool want_ll) {
    const int tid = threadIdx.x;
    const int r = tid >> 2, lq = tid & 3;   // 128 rows x 4 quads of outputs
    float ll = 0.f;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    const bool qlive = 4 * lq < dm.dop;
    if (r < nb && qlive) {
        const float* h = sm.H + r * dm.dh;
        const float* Ap = sm.A + 4 * lq;
#pragma unroll 2
        for (int k = 0; k < dm.dh; k += 4) {   // dh % 4 == 0; same summation order as the scalar loop
            const float4 hv = *reinterpret_cast<const float4*>(h + k);
            const float4 a0 = *reinterpret_cast<const float4*>(Ap + (k + 0) * dm.dop), a1 = *reinterpret_cast<const float4*>(Ap + (k + 1) * dm.dop);
            const float4 a2 = *reinterpret_cast<const float4*>(Ap + (k + 2) * dm.dop), a3 = *reinterpret_cast<const float4*>(Ap + (k + 3) * dm.dop);
            z[0] += hv.x * a0.x; z[1] += hv.x * a0.y; z[2] += hv.x * a0.z; z[3] += hv.x * a0.w;
            z[0] += hv.y * a1.x; z[1] += hv.y * a1.y; z[2] += hv.y * a1.z; z[3] += hv.y * a1.w;
            z[0] += hv.z * a2.x; z[1] += hv.z * a2.y; z[2] += hv.z * a2.z; z[3] += hv.z * a2.w;
            z[0] += hv.w * a3.x; z[1] += hv.w * a3.y; z[2] += hv.w * a3.z; z[3] += hv.w * a3.w;
        }
    }
    float y[4] = {0.f, 0.f, 0.f, 0.f};
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int l = 4 * lq + c;
        const bool ok = r < nb && l < dm.dout;
        z[c] = ok ? z[c] + sm.a[l] : -INFINITY;
        if (ok) y[c] = __ldg(X + (size_t)sm.rows[r] * ld + dm.din + l);
        m = fmaxf(m, z[c]);
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    float e[4], s = 0.f, ys = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) { e[c] = (z[c] > -INFINITY) ? expf(z[c] - m) : 0.f; s += e[c]; ys += y[c]; }
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
    ys += __shfl_xor_sync(0xffffffffu, ys, 1); ys += __shfl_xor_sync(0xffffffffu, ys, 2);
    if (r < nb && qlive) {
        float4 o;
        float* op = &o.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int l = 4 * lq + c;
            const float pr = e[c] / s;
            op[c] = (l < dm.dout) ? y[c] - pr * ys : 0.f;
            if (want_ll && l < dm.dout && y[c] != 0.f) ll += y[c] * logf(pr);
        }
        *reinterpret_cast<float4*>(sm.dZ2 + r * dm.dop + 4 * lq) = o;
    }
    __syncthreads();
    return ll;
}

// S3: gA[k][l] (+)= sum_r H[r][k] dZ2[r][l];  ga[l] (+)= sum_r dZ2[r][l]
__device__ void nn_stage3(const NNDims& dm, const NNSmem& sm, int nb, bool first) {
    // one (hidden unit k, output quad) per thread: 4 independent accumulators, rows walked in order (same sums as a scalar loop)
    const int nq = dm.dop >> 2;
    for (int i = threadIdx.x; i < dm.dh * nq; i += NN_THREADS) {
        const int k = i / nq, lq = i % nq;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const float* hp = sm.H + k;
        const float* zp = sm.dZ2 + 4 * lq;
#pragma unroll 4
        for (int r = 0; r < nb; ++r) {
            const float hv = hp[r * dm.dh];
            const float4 dz = *reinterpret_cast<const float4*>(zp + r * dm.dop);
            s0 += hv * dz.x; s1 += hv * dz.y; s2 += hv * dz.z; s3 += hv * dz.w;
        }
        float4* o = reinterpret_cast<float4*>(sm.gA + k * dm.dop + 4 * lq);
        if (first) *o = make_float4(s0, s1, s2, s3);
        else { const float4 pv = *o; *o = make_float4(pv.x + s0, pv.y + s1, pv.z + s2, pv.w + s3); }
    }
    for (int l = NN_THREADS - 1 - threadIdx.x; l < dm.dop; l += NN_THREADS) {   // the last threads (idle above)
        float s = 0.f;
        for (int r = 0; r < nb; ++r) s += sm.dZ2[r * dm.dop + l];
        sm.ga[l] = first ? s : sm.ga[l] + s;
    }
    __syncthreads();
}

// S4: dZ1[r][k] = H (dH - rowsum(dH H)), dH = dZ2 A^T;  gb[k] (+)= sum_r dZ1[r][k]  (warp per row)
__device__ void nn_stage4(const NNDims& dm, const NNSmem& sm, int nb, bool first, float* __restrict__ gb_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncg = dm.dh >> 2;
    const bool live = lane < ncg;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = warp; r < nb; r += NN_WARPS) {
        float dh[4] = {0.f, 0.f, 0.f, 0.f};
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            h = *reinterpret_cast<const float4*>(sm.H + r * dm.dh + 4 * lane);
            for (int l = 0; l < dm.dout; ++l) {
                const float dz = sm.dZ2[r * dm.dop + l];
                const float4 av = *reinterpret_cast<const float4*>(sm.At + l * dm.dh + 4 * lane);
                dh[0] = fmaf(dz, av.x, dh[0]); dh[1] = fmaf(dz, av.y, dh[1]); dh[2] = fmaf(dz, av.z, dh[2]); dh[3] = fmaf(dz, av.w, dh[3]);
            }
        }
        const float s = warp_sum(dh[0] * h.x + dh[1] * h.y + dh[2] * h.z + dh[3] * h.w);
        if (live) {
            const float4 o = make_float4(h.x * (dh[0] - s), h.y * (dh[1] - s), h.z * (dh[2] - s), h.w * (dh[3] - s));
            *reinterpret_cast<float4*>(sm.dZ1 + r * dm.dh + 4 * lane) = o;
            cs[0] += o.x; cs[1] += o.y; cs[2] += o.z; cs[3] += o.w;
        }
    }
    if (live) *reinterpret_cast<float4*>(sm.gb + warp * dm.dh + 4 * lane) = make_float4(cs[0], cs[1], cs[2], cs[3]);
    __syncthreads();
    for (int k = threadIdx.x; k < dm.dh; k += NN_THREADS) {
        float s = 0.f;
        for (int w = 0; w < NN_WARPS; ++w) s += sm.gb[w * dm.dh + k];
        const float tot = first ? s : gb_out[k] + s;
        gb_out[k] = tot;
        sm.gb[k] = tot;   // column k of the partials is read by this thread only: the total stays in shared memory for the fused small update
    }
    __syncthreads();
}

constexpr int TC5_NT = 64;    // input columns (j) per S5 tile = MMA N
constexpr int TC5_KH = 64;    // minibatch rows (K) per staging step

// S5 on the tensor cores: GB[n][j] (+)= sum_r dZ1[r][n] X[row_r][j]  as  D[128 x 64] = dZ1^T[128 x K] * (X^T)[64 x K]^T per
// tile of 64 inputs, K = minibatch rows in halves of 64, 3xTF32.  Up to 8 tiles (512 TMEM columns) accumulate before one
// epilogue pass writes them out with 16-byte stores.  The A operand (dZ1^T, hi/lo) is re-staged from shared memory per
// (tile group, K half); the B operand (X^T tile) is gathered from global memory with register prefetch.
__device__ void nn_stage5_tc(const NNDims& dm, const NNSmem& sm, const float* __restrict__ X, int ld, int nb, bool first,
                             float* __restrict__ GB, TcCtx& tc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int KG = TC5_KH / 4;                          // 16-byte K chunks per staging step
    const int LBOA = NN_RB * 16;                        // A operand: 128 rows (hidden units, zero padded)
    const int rowsB = TC5_NT + 1, LBOB = rowsB * 16;    // +1 row: the 4-byte scatter below is bank-conflict free
    const int A_FLOATS = KG * NN_RB * 4, B_FLOATS = KG * rowsB * 4;
    float* Ahi = sm.tile; float* Alo = Ahi + A_FLOATS; float* Bhi = Alo + A_FLOATS; float* Blo = Bhi + B_FLOATS;
    const uint32_t idesc = nn_umma_idesc(NN_RB, TC5_NT);
    const int ntiles = (dm.din + TC5_NT - 1) / TC5_NT;
    const int halves = (nb + TC5_KH - 1) / TC5_KH;
    // B staging: thread -> (row r of the half, float4 column jq): 64 x 16 = 1024 float4 = 2 per thread
    const int sr = tid % TC5_KH, sq0 = tid / TC5_KH;     // jq = sq0 and sq0 + 8
    float4 xv0, xv1;
    auto prefetch = [&](int t, int h) {
        const int r = h * TC5_KH + sr, j0 = t * TC5_NT;
        xv0 = make_float4(0.f, 0.f, 0.f, 0.f); xv1 = xv0;
        if (r < nb) {
            const float* xrow = X + (size_t)sm.rows[r] * ld + j0;
            if (j0 + 4 * sq0 < dm.din) xv0 = __ldg(reinterpret_cast<const float4*>(xrow + 4 * sq0));
            if (j0 + 4 * (sq0 + 8) < dm.din) xv1 = __ldg(reinterpret_cast<const float4*>(xrow + 4 * (sq0 + 8)));
        }
    };
    auto scatter = [&](const float4 v, int jq) {          // X[r][4jq..4jq+3] -> operand rows j = 4jq + i, chunk sr/4, slot sr%4
        float4 hi, lo;
        nn_split4(v, hi, lo);
        const int base = ((sr >> 2) * rowsB + 4 * jq) * 4 + (sr & 3);
        Bhi[base] = hi.x; Bhi[base + 4] = hi.y; Bhi[base + 8] = hi.z; Bhi[base + 12] = hi.w;
        Blo[base] = lo.x; Blo[base + 4] = lo.y; Blo[base + 8] = lo.z; Blo[base + 12] = lo.w;
    };
    uint64_t e_ahi[TC5_KH / 8], e_alo[TC5_KH / 8], e_bhi[TC5_KH / 8], e_blo[TC5_KH / 8];
#pragma unroll
    for (int k8 = 0; k8 < TC5_KH / 8; ++k8) {
        e_ahi[k8] = nn_umma_desc(smem_u32(Ahi) + (uint32_t)(2 * k8) * LBOA, LBOA); e_alo[k8] = nn_umma_desc(smem_u32(Alo) + (uint32_t)(2 * k8) * LBOA, LBOA);
        e_bhi[k8] = nn_umma_desc(smem_u32(Bhi) + (uint32_t)(2 * k8) * LBOB, LBOB); e_blo[k8] = nn_umma_desc(smem_u32(Blo) + (uint32_t)(2 * k8) * LBOB, LBOB);
    }
    for (int t0 = 0; t0 < ntiles; t0 += 8) {
        const int t1 = (t0 + 8 < ntiles) ? t0 + 8 : ntiles;
        for (int h = 0; h < halves; ++h) {
            prefetch(t0, h);
            // every MMA issued so far has completed before the A operand is overwritten
            if (tc.commits[0] > 0) mbar_wait(tc.bar[0], (tc.commits[0] - 1u) & 1u);
            __syncthreads();
            for (int u = tid; u < NN_RB * KG; u += NN_THREADS) {   // A: (hidden unit n, chunk kg) <- dZ1[64h + 4kg .. +3][n]
                const int n = u % NN_RB, kg = u / NN_RB;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
                if (n < dm.dh) {
                    const int r = h * TC5_KH + 4 * kg;
                    v.x = (r + 0 < nb) ? sm.dZ1[(r + 0) * dm.dh + n] : 0.f; v.y = (r + 1 < nb) ? sm.dZ1[(r + 1) * dm.dh + n] : 0.f;
                    v.z = (r + 2 < nb) ? sm.dZ1[(r + 2) * dm.dh + n] : 0.f; v.w = (r + 3 < nb) ? sm.dZ1[(r + 3) * dm.dh + n] : 0.f;
                }
                nn_split4(v, hi, lo);
                *reinterpret_cast<float4*>(Ahi + (kg * NN_RB + n) * 4) = hi;
                *reinterpret_cast<float4*>(Alo + (kg * NN_RB + n) * 4) = lo;
            }
            for (int t = t0; t < t1; ++t) {
                // the MMAs of the previous tile read the B buffer
                if (t > t0 && tc.commits[0] > 0) mbar_wait(tc.bar[0], (tc.commits[0] - 1u) & 1u);
                scatter(xv0, sq0);
                scatter(xv1, sq0 + 8);
                if (t + 1 < t1) prefetch(t + 1, h);
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (warp == 0 && elect_one()) {
                    tc_fence_after();
                    const uint32_t td = tc.tmem + (uint32_t)((t - t0) * TC5_NT);
#pragma unroll
                    for (int k8 = 0; k8 < TC5_KH / 8; ++k8) {
                        umma_tf32_ss(td, e_ahi[k8], e_bhi[k8], idesc, (h > 0) || (k8 > 0));
                        umma_tf32_ss(td, e_ahi[k8], e_blo[k8], idesc, true);
                        umma_tf32_ss(td, e_alo[k8], e_bhi[k8], idesc, true);
                    }
                    umma_commit(tc.bar[0]);
                }
                tc.commits[0] += 1;
            }
        }
        // ---- epilogue of the group: TMEM lane = hidden unit n, columns = inputs of the group's tiles; all 16 warps
        // (warp w reads lanes 32 (w % 4) .. +31, the tiles w / 4, w / 4 + 4 of the group)
        mbar_wait(tc.bar[0], (tc.commits[0] - 1u) & 1u);
        tc_fence_after();
        {
            const int q = warp & 3, n = 32 * q + lane;
            const uint32_t tbase = tc.tmem + ((uint32_t)(32 * q) << 16);
            for (int t = t0 + (warp >> 2); t < t1; t += 4) {
                for (int c0 = 0; c0 < TC5_NT; c0 += 16) {
                    float v[16];
                    nn_tmem_ld16(tbase + (uint32_t)((t - t0) * TC5_NT + c0), v);
                    const int j = t * TC5_NT + c0;
                    if (n < dm.dh && j < dm.din) {   // din is a multiple of 4, so float4 groups are all-or-nothing
                        float* dst = GB + (size_t)n * dm.din + j;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            if (j + 4 * q4 < dm.din) {
                                float4 o = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                                float4* d4 = reinterpret_cast<float4*>(dst + 4 * q4);
                                if (!first) { const float4 pv = *d4; o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w; }
                                *d4 = o;
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
}

#ifdef NN_ENABLE_S5_TMEM
// ---------------------------------------------------------------------------------------------------------
// S5, TMA-fed, X on the tensor-memory side -- written at the end of round 1, NOT YET RUN ON A GPU.  Compiled only with
// -DNN_ENABLE_S5_TMEM (SGM_BUILD_DEFS=-DNN_ENABLE_S5_TMEM python __graft_entry__.py --force), then selected by SGM_NN_S5=tmem.
//   GB^T[j][n] = sum_r X[row_r][j] dZ1[r][n]      M = 128 inputs j per tile, N = Npad hidden units, K = Kp = round8(nb) rows
//   * B operand: dZ1^T staged ONCE per call, K-major no-swizzle core matrices, hi / lo (same layout the register-staged S5 uses for A);
//   * A operand: the X tile arrives as four gather4 blocks [Kp rows][32 inputs] (the SW128 tiles S1 loads); block q belongs to the
//     warps of lane quarter q, each thread transposes its lane's column (one ld.shared.32 per row: a row of a block is one
//     128-byte line, so a warp reads 32 distinct banks) into TMEM columns r (hi) / 128 + r (lo) with tcgen05.st;
//   * 3 x Kp / 8 A-from-TMEM MMAs per tile into one of two TMEM accumulators; the epilogue of tile t - 1 (lane = input j: coalesced
//     stores to GB[n][j]) runs while the MMAs of tile t execute; the TMA requests of tile t + 2 go out as soon as tile t is in TMEM.
// TMEM: accumulators [0, 2 Npad), A hi [256, 384), A lo [384, 512).  Shared memory (inside the H | dZ1 | tile ring; dZ1 is read
// before anything overwrites it): two raw X stages of 4 x Kp x 128 B at the start, dZ1^T hi / lo behind them.
// ---------------------------------------------------------------------------------------------------------
// the two raw stages must end behind dZ1 (H | dZ1 come first in the ring; dZ1 is read while the B operand is written behind the stages)
__device__ __forceinline__ bool nn_s5_tmem_fits(const NNDims& dm, int nb, uint32_t ring_bytes) {
    const int Kp = (nb + 7) & ~7, Npad = (dm.dh + 15) & ~15;
    const long stages = 2L * 4 * Kp * 128, need = stages + 2L * (Kp / 4) * Npad * 16;
    return Kp <= 128 && Npad <= 112 && need <= (long)ring_bytes && stages >= 2L * NN_RB * dm.dh * 4;
}

__device__ SGM_NOINLINE void nn_stage5_tmem(const NNDims& dm, const NNSmem& sm, const CUtensorMap* xmap, int nb, bool first,
                                            float* __restrict__ GB, TcCtx& tc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = warp & 3, cg = warp >> 2;
    const int Kp = (nb + 7) & ~7, Npad = (dm.dh + 15) & ~15, KG = Kp >> 2;
    const uint32_t BLK = (uint32_t)Kp * 128u, STG = 4u * BLK;
    const uint32_t x0 = sm.ring, bop = sm.ring + 2u * STG;            // raw X stages | dZ1^T hi | dZ1^T lo
    const uint32_t LBOB = (uint32_t)Npad * 16u, PLANE = (uint32_t)KG * LBOB;
    const int ntiles = (dm.din + 127) >> 7;
    SGM_EMU_NOTE("nn_stage5_tmem");
    const uint32_t bars = smem_u32(sm.s1_bars);
    const uint32_t b_full = bars, b_dfull = bars + 16;               // full[2] | dfull[2]
    const uint32_t idesc = nn_umma_idesc(NN_RB, Npad);
    // ---- B operand from dZ1 (shared, [r][dh]); rows >= nb and hidden units >= dh are zero
    __syncthreads();
    for (int u = tid; u < Npad * KG; u += NN_THREADS) {
        const int n = u % Npad, kg = u / Npad, r = 4 * kg;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
        if (n < dm.dh) {
            v.x = (r + 0 < nb) ? sm.dZ1[(r + 0) * dm.dh + n] : 0.f; v.y = (r + 1 < nb) ? sm.dZ1[(r + 1) * dm.dh + n] : 0.f;
            v.z = (r + 2 < nb) ? sm.dZ1[(r + 2) * dm.dh + n] : 0.f; v.w = (r + 3 < nb) ? sm.dZ1[(r + 3) * dm.dh + n] : 0.f;
        }
        nn_split4(v, hi, lo);
        sts128(bop + (uint32_t)(kg * Npad + n) * 16u, hi);
        sts128(bop + PLANE + (uint32_t)(kg * Npad + n) * 16u, lo);
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);
        fence_mbar_init();
    }
    fence_proxy_async();      // operand stores -> the tensor core
    __syncthreads();                                                   // ... and dZ1 may now be overwritten by the X stages
    auto issue_tile = [&](int t) {   // warp 0: 4 blocks x Kp / 4 gather4 requests
        const uint32_t dst = x0 + (uint32_t)(t & 1) * STG, bar = b_full + 8 * (t & 1);
        if (lane == 0) mbar_expect_tx(bar, STG);
        __syncwarp();
        for (int i = lane; i < 4 * KG; i += 32) {
            const int b = i / KG, g = i % KG, r = 4 * g;
            tma_gather4(dst + (uint32_t)b * BLK + (uint32_t)g * 512u, xmap, 128 * t + 32 * b, sm.rows[r < nb ? r : 0], sm.rows[r + 1 < nb ? r + 1 : 0],
                           sm.rows[r + 2 < nb ? r + 2 : 0], sm.rows[r + 3 < nb ? r + 3 : 0], bar);
        }
    };
    if (warp == 0) { issue_tile(0); if (ntiles > 1) issue_tile(1); }
    auto epilogue = [&](int t) {     // TMEM lane = input j of tile t, columns = hidden units
        mbar_wait(b_dfull + 8 * (t & 1), (uint32_t)(t >> 1) & 1u);
        tc_fence_after();
        const int j = 128 * t + 32 * q + lane;
        const uint32_t tb = tc.tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)((t & 1) * Npad);
        for (int c0 = 32 * cg; c0 < 32 * cg + 32 && c0 < Npad; c0 += 16) {
            float v[16];
            nn_tmem_ld16(tb + (uint32_t)c0, v);
            if (j < dm.din) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int n = c0 + i;
                    if (n < dm.dh) {
                        float* dst = GB + (size_t)n * dm.din + j;
                        *dst = first ? v[i] : *dst + v[i];
                    }
                }
            }
        }
        tc_fence_before();
    };
    for (int t = 0; t < ntiles; ++t) {
        const uint32_t xs = x0 + (uint32_t)(t & 1) * STG + (uint32_t)q * BLK;
        mbar_wait(b_full + 8 * (t & 1), (uint32_t)(t >> 1) & 1u);
        // the MMAs of tile t - 1 read the TMEM A buffer: they must be done before it is rewritten
        if (t > 0) { mbar_wait(b_dfull + 8 * ((t - 1) & 1), (uint32_t)((t - 1) >> 1) & 1u); tc_fence_after(); }
        // transpose: this thread's input column (lane) of block q, rows 32 cg .. 32 cg + 31 -> TMEM columns
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int r0 = 32 * cg + 16 * h2;
            if (r0 < Kp) {   // warp-uniform
                float hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int r = r0 + i;
                    float x = 0.f;
                    if (r < Kp) x = lds32(xs + (uint32_t)r * 128u + ((uint32_t)((lane >> 2) ^ (r & 7)) << 4) + ((uint32_t)(lane & 3) << 2));
                    hi[i] = nn_tf32(x);
                    lo[i] = nn_tf32(x - hi[i]);
                }
                const uint32_t ta = tc.tmem + ((uint32_t)(32 * q) << 16) + 256u + (uint32_t)r0;
                tmem_st16(ta, hi);
                tmem_st16(ta + 128u, lo);
            }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            if (t + 2 < ntiles) issue_tile(t + 2);     // stage t & 1 is free: its tile lives in TMEM now
            tc_fence_after();
            if (elect_one()) {
                const uint32_t td = tc.tmem + (uint32_t)((t & 1) * Npad);
                for (int k8 = 0; k8 < (Kp >> 3); ++k8) {
                    const uint64_t bhi = nn_umma_desc(bop + (uint32_t)(2 * k8) * LBOB, LBOB), blo = nn_umma_desc(bop + PLANE + (uint32_t)(2 * k8) * LBOB, LBOB);
                    const uint32_t ahi = tc.tmem + 256u + 8u * (uint32_t)k8, alo = ahi + 128u;
                    umma_tf32_ts(td, ahi, bhi, idesc, k8 > 0);
                    umma_tf32_ts(td, ahi, blo, idesc, true);
                    umma_tf32_ts(td, alo, bhi, idesc, true);
                }
                umma_commit(b_dfull + 8 * (t & 1));
            }
            __syncwarp();
        }
        if (t > 0) epilogue(t - 1);                      // overlaps the MMAs of tile t
    }
    epilogue(ntiles - 1);
    __syncthreads();
}

#endif  // NN_ENABLE_S5_TMEM

enum : int { NEED_A = 1, NEED_B = 2, NEED_a = 4, NEED_b = 8, NEED_ALL = 15 };

// One gradient evaluation of the log-likelihood SUM over the minibatch at `th`, for the tensors in
// `need`, written to Gout (internal layout).  h_valid: sm.H already holds softmax(X B + b) for th.
// Rows come from idx (deterministic), the keyed permutation, or the contiguous range [row0, row0+n).
__device__ float nn_eval(const NNDims& dm, const NNSmem& sm, const float* __restrict__ X, int ld, int n,
                         const Perm* perm, const int32_t* __restrict__ idx, int64_t row0,
                         const float* __restrict__ th, int need, bool h_valid, bool want_ll, float* __restrict__ Gout, TcCtx& tc,
                         const NNMaps& mp, int reload = LOAD_ALL) {
    nn_load_small(dm, sm, th, reload);
    float ll = 0.f;
    const int nblocks = (n + NN_RB - 1) / NN_RB;
    for (int rb = 0; rb < nblocks; ++rb) {
        const int r0 = rb * NN_RB;
        const int nb = (n - r0 < NN_RB) ? n - r0 : NN_RB;
        const bool first = rb == 0;
        if (!(h_valid && nblocks == 1)) {
            __syncthreads();
            for (int r = threadIdx.x; r < NN_RB; r += NN_THREADS)
                sm.rows[r] = (r < nb) ? (idx ? idx[r0 + r] : (perm ? (int)perm_at(*perm, (uint32_t)(r0 + r)) : (int)(row0 + r0 + r))) : 0;
            __syncthreads();
            NN_TICK(6)
            if (mp.x) nn_stage1_tma(dm, sm, mp, nb, tc);
            else nn_stage1_tc(dm, sm, X, ld, nb, th, tc);
            NN_TICK(0)
        }
        ll += nn_stage2(dm, sm, X, ld, nb, want_ll);
        NN_TICK(1)
        if (need & (NEED_A | NEED_a)) nn_stage3(dm, sm, nb, first);
        NN_TICK(2)
        if (need & (NEED_b | NEED_B)) nn_stage4(dm, sm, nb, first, Gout + dm.offb);
        NN_TICK(3)
#ifdef NN_ENABLE_S5_TMEM   // compiled out by default: even an untaken call changes the register allocation of the default kernels
        if ((need & NEED_B) && mp.x && nn_s5_tmem && nn_s5_tmem_fits(dm, nb, sm.ring_bytes)) nn_stage5_tmem(dm, sm, mp.x, nb, first, Gout + dm.offB, tc);
        else
#endif
        if (need & NEED_B) nn_stage5_tc(dm, sm, X, ld, nb, first, Gout + dm.offB, tc);
        NN_TICK(4)
    }
    if (need & (NEED_A | NEED_a)) {
        for (int i = threadIdx.x; i < dm.dh * dm.dout; i += NN_THREADS) {
            const int k = i / dm.dout, l = i % dm.dout;
            Gout[dm.offA + k + dm.dh * l] = sm.gA[k * dm.dop + l];
        }
        for (int l = threadIdx.x; l < dm.dout; l += NN_THREADS) Gout[dm.offa + l] = sm.ga[l];
    }
    __syncthreads();
    return ll;
}

// sum of squares of one weight tensor (for the lambda gradients / prior value)
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

__global__ void __launch_bounds__(NN_THREADS, 1) softmax_net_chain_kernel(const ChainArgs a, const NNDims dm, float* Gall, float* Ghatall,
                                                                          const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmTh,
                                                                          const __grid_constant__ CUtensorMap tmHat, int use_tma) {
    SGM_DYN_SMEM(float, smem_f, 16);
    NNSmem sm;
    nn_carve(dm, smem_f, sm);
    const bool seidel = (a.algo == ALGO_SGLD || a.algo == ALGO_SGHMC);
    TcCtx tc;
    nn_tc_init(sm, tc);
    if (threadIdx.x == 0) nn_t_prev[blockIdx.x & 1023] = clock64();
    for (int c = blockIdx.x; c < a.C; c += gridDim.x) {
        NNChainState st;
        st.th = a.theta + (size_t)c * a.Pst; st.mo = a.mom + (size_t)c * a.Pst;
        st.hat = a.cv ? a.theta_hat + (size_t)c * a.hat_stride : nullptr;
        st.ghat = a.cv ? a.g_hat + (size_t)c * a.hat_stride : nullptr;
        st.G = Gall + (size_t)c * a.Pst; st.Ghat = Ghatall ? Ghatall + (size_t)c * a.Pst : nullptr;
        for (int t = 0; t < MAX_TENSORS; ++t) st.zeta[t] = a.thermo[c * MAX_TENSORS + t];
        st.ssq_valid = 0u; st.ssq_hat_valid = 0u;
        st.fast = seidel && !a.cv && nn_stream_updates;
        if (st.fast) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int si = (int)threadIdx.x + NN_THREADS * k;
                st.mo_reg[k] = si < nn_small_count(dm) ? st.mo[nn_small_elem(dm, si)] : 0.f;
            }
            for (int w4 = 0; w4 < 4; ++w4) { st.lam_th[w4] = st.th[dm.offl + w4]; st.lam_mo[w4] = st.mo[dm.offl + w4]; }
        }
        bool small_valid = false;   // shared copies of A / a / b belong to this chain's theta
        const NNMaps mpT{use_tma ? &tmX : nullptr, &tmTh, c}, mpH{use_tma ? &tmX : nullptr, &tmHat, a.hat_stride ? c : 0};
        for (int it = 0; it < a.iters; ++it) {
            const uint64_t step = a.step0 + (uint64_t)it;
            if (a.algo == ALGO_SGHMC) {  // refresh (R/dynamics.R:62)
                for (int e = threadIdx.x; e < a.P; e += NN_THREADS) {
                    const int t = nn_tensor_of(dm, e);
                    const float z = a.det ? a.noise[(size_t)c * a.noise_len + nn_flat_of_internal(dm, e)]
                                          : normal_at((uint32_t)e, a.chain0 + (uint32_t)c, word2_of(step, a.L, 0), TAG_NOISE, a.k0, a.k1);
                    st.mo[e] = 0.5f * (a.tc.sqrt_2eps[t] * z);
                }
                if (st.fast) {   // the register copies of the small momenta take the same refreshed values
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int si = (int)threadIdx.x + NN_THREADS * k;
                        if (si < nn_small_count(dm)) {
                            const int e = nn_small_elem(dm, si), t = nn_tensor_of(dm, e);
                            const float z = a.det ? a.noise[(size_t)c * a.noise_len + nn_flat_of_internal(dm, e)]
                                                  : normal_at((uint32_t)e, a.chain0 + (uint32_t)c, word2_of(step, a.L, 0), TAG_NOISE, a.k0, a.k1);
                            st.mo_reg[k] = 0.5f * (a.tc.sqrt_2eps[t] * z);
                        }
                    }
                    for (int w4 = 0; w4 < 4; ++w4) {
                        const int e = dm.offl + w4;
                        const float z = a.det ? a.noise[(size_t)c * a.noise_len + nn_flat_of_internal(dm, e)]
                                              : normal_at((uint32_t)e, a.chain0 + (uint32_t)c, word2_of(step, a.L, 0), TAG_NOISE, a.k0, a.k1);
                        st.lam_mo[w4] = 0.5f * (a.tc.sqrt_2eps[4 + w4] * z);
                    }
                }
                __syncthreads();
            }
            for (int l = 0; l < a.L; ++l) {
                const int sub = l + 1;
                const Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2_of(step, a.L, sub), a.index_tag, a.k0, a.k1);
                const int32_t* idx = a.det ? a.idx + (((size_t)c * a.iters + it) * a.L + l) * a.n : nullptr;
                const Perm* pp = a.det ? nullptr : &perm;
                if (a.cv) nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.hat, NEED_ALL, false, false, st.Ghat, tc, mpH);
                if (seidel && st.fast) {
                    // the small tensors are updated in shared memory / registers (written through): nothing to reload between evaluations
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_A, false, false, st.G, tc, mpT, small_valid ? 0 : LOAD_ALL);
                    small_valid = true;
                    nn_update_small(a, dm, sm, st, 0, c, step, sub);
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_B, true, false, st.G, tc, mpT, 0);
                    nn_update_tensor(a, dm, sm, st, 1, c, step, sub, nullptr);
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_a, false, false, st.G, tc, mpT, 0);
                    nn_update_small(a, dm, sm, st, 2, c, step, sub);
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_b, true, false, st.G, tc, mpT, 0);
                    nn_update_small(a, dm, sm, st, 3, c, step, sub);
                    for (int t = 4; t < 8; ++t) nn_update_lambda(a, dm, sm, st, t, c, step, sub);
                } else if (seidel) {
                    // A   (shared A / a / b are current from the previous sub-step except b, unless the mode's were loaded)
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_A, false, false, st.G, tc, mpT, (a.cv || !small_valid) ? LOAD_ALL : LOAD_b);
                    small_valid = true;
                    nn_update_tensor(a, dm, sm, st, 0, c, step, sub, nullptr);
                    // B: H still valid (depends on B, b only)
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_B, true, false, st.G, tc, mpT, LOAD_A);
                    nn_update_tensor(a, dm, sm, st, 1, c, step, sub, nullptr);
                    // a: B changed -> recompute H
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_a, false, false, st.G, tc, mpT, 0);
                    nn_update_tensor(a, dm, sm, st, 2, c, step, sub, nullptr);
                    // b
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_b, true, false, st.G, tc, mpT, LOAD_a);
                    nn_update_tensor(a, dm, sm, st, 3, c, step, sub, nullptr);
                    for (int t = 4; t < 8; ++t) nn_update_tensor(a, dm, sm, st, t, c, step, sub, nullptr);
                } else {
                    nn_eval(dm, sm, a.X, a.ld, a.n, pp, idx, 0, st.th, NEED_ALL, false, false, st.G, tc, mpT);
                    float uu[MAX_TENSORS];
                    // All tensors move from ONE evaluation.  The mode finder writes theta in place, so the
                    // precisions (whose gradient reads ||W||^2) go first and the weights see the OLD precisions.
                    float lam_old[4], lam_new[4];
                    for (int t = 0; t < 4; ++t) lam_old[t] = st.th[dm.offl + t];
                    for (int t = 7; t >= 4; --t) nn_update_tensor(a, dm, sm, st, t, c, step, sub, &uu[t]);
                    if (a.algo == ALGO_OPT) {
                        for (int t = 0; t < 4; ++t) lam_new[t] = st.th[dm.offl + t];
                        __syncthreads();
                        if (threadIdx.x == 0) for (int t = 0; t < 4; ++t) st.th[dm.offl + t] = lam_old[t];
                        __syncthreads();
                    }
                    for (int t = 3; t >= 0; --t) nn_update_tensor(a, dm, sm, st, t, c, step, sub, &uu[t]);
                    if (a.algo == ALGO_OPT) {
                        if (threadIdx.x == 0) for (int t = 0; t < 4; ++t) st.th[dm.offl + t] = lam_new[t];
                        __syncthreads();
                    }
                    if (a.algo == ALGO_SGNHT) {  // theta += u, then thermostats (R/update.R:75-80)
                        for (int e = threadIdx.x; e < a.P; e += NN_THREADS) st.th[e] += st.mo[e];
                        for (int t = 0; t < 8; ++t) {
                            int lo, hi;
                            nn_tensor_range(dm, t, lo, hi);
                            st.zeta[t] += uu[t] / (float)(hi - lo) - a.tc.eps[t];
                        }
                        __syncthreads();
                    }
                }
            }
            // the last updates of the step (the precisions: thread 0 writes them through to global memory) must be visible to the
            // threads that store the state below -- without this barrier the stored chain carried the PREVIOUS step's lambdaB / lambdaa /
            // lambdab while getParams was right (a race found by the emulated build: tests/test_emu_parity.py::test_streamed_sampling)
            __syncthreads();
            bool bad = false;
            for (int e = threadIdx.x; e < a.P; e += NN_THREADS) {
                const float v = st.th[e];
                bad |= !isfinite(v);
                if (a.store) a.store[((size_t)(a.store_pos + it) * a.C + c) * a.P + nn_flat_of_internal(dm, e)] = v;
            }
            if (bad) *a.diverged = 1;
            __syncthreads();
            NN_TICK(7)
        }
        if (threadIdx.x == 0)
            for (int t = 0; t < MAX_TENSORS; ++t) a.thermo[c * MAX_TENSORS + t] = st.zeta[t];
    }
    nn_tc_fini(tc);
}

// full-data gradient partials: CTA (x, m) sums rows [row0 + x*rows_per, ...) at theta_hat[m] into part[(m*gx + x)][Pst]
__global__ void __launch_bounds__(NN_THREADS, 1) softmax_net_fullgrad_kernel(const ChainArgs a, const NNDims dm, int64_t row0, int64_t row1,
                                                                             int rows_per, const float* __restrict__ theta_hat, float* __restrict__ part,
                                                                             const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmHat, int use_tma) {
    SGM_DYN_SMEM(float, smem_f, 16);
    NNSmem sm;
    nn_carve(dm, smem_f, sm);
    const int m = blockIdx.y;
    const int64_t lo = row0 + (int64_t)blockIdx.x * rows_per;
    int64_t hi = lo + rows_per; if (hi > row1) hi = row1;
    float* G = part + ((size_t)m * gridDim.x + blockIdx.x) * a.Pst;
    if (lo >= hi) { for (int e = threadIdx.x; e < a.Pst; e += NN_THREADS) G[e] = 0.f; return; }
    TcCtx tc;
    nn_tc_init(sm, tc);
    nn_eval(dm, sm, a.X, a.ld, (int)(hi - lo), nullptr, nullptr, lo, theta_hat + (size_t)m * a.Pst, NEED_ALL, false, false, G, tc,
            NNMaps{use_tma ? &tmX : nullptr, &tmHat, m});
    for (int e = dm.offl + threadIdx.x; e < a.Pst; e += NN_THREADS) G[e] = 0.f;
    nn_tc_fini(tc);
}

__global__ void softmax_net_part_reduce(const float* __restrict__ part, int gx, int Pst, double* __restrict__ gsum, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int m = i / Pst, e = i % Pst;
    double s = 0.0;
    for (int x = 0; x < gx; ++x) s += (double)part[((size_t)m * gx + x) * Pst + e];
    gsum[i] = s;
}

__global__ void __launch_bounds__(NN_THREADS) softmax_net_finish_kernel(const ChainArgs a, const NNDims dm, const double* __restrict__ gsum,
                                                                        const float* __restrict__ theta_hat, float* __restrict__ g_hat,
                                                                        float* __restrict__ theta) {
    __shared__ float red[64];
    const int m = blockIdx.x;
    const float* th = theta_hat + (size_t)m * a.Pst;
    float ssq[4];
    for (int t = 0; t < 4; ++t) {
        int lo, hi;
        nn_tensor_range(dm, t, lo, hi);
        ssq[t] = nn_sumsq(th + lo, hi - lo, red);
    }
    for (int e = threadIdx.x; e < a.P; e += NN_THREADS) {
        const int t = nn_tensor_of(dm, e);
        float g = (t < 4) ? (float)gsum[(size_t)m * a.Pst + e] : 0.f;
        if (a.use_prior) g += nn_prior_grad(dm, th, e, t, t >= 4 ? ssq[t - 4] : 0.f);
        g_hat[(size_t)m * a.Pst + e] = g;
        theta[(size_t)m * a.Pst + e] = th[e];
    }
}

__global__ void __launch_bounds__(NN_THREADS, 1) softmax_net_logpost_kernel(const ChainArgs a, const NNDims dm, uint32_t word2,
                                                                            const float* __restrict__ theta, float* Gscratch,
                                                                            double* __restrict__ out, const __grid_constant__ CUtensorMap tmX,
                                                                            const __grid_constant__ CUtensorMap tmTh, int use_tma) {
    SGM_DYN_SMEM(float, smem_f, 16);
    NNSmem sm;
    nn_carve(dm, smem_f, sm);
    const int c = blockIdx.x;
    const float* th = theta + (size_t)c * a.Pst;
    const Perm perm = make_perm(a.N, a.perm_bits, a.chain0 + (uint32_t)c, word2, TAG_CHECK_INDEX, a.k0, a.k1);
    const int32_t* idx = (a.det == 1) ? a.idx + (size_t)c * a.n : nullptr;   // det == 2: contiguous rows 0..n-1 (held-out set)
    TcCtx tc;
    nn_tc_init(sm, tc);
    float ll = nn_eval(dm, sm, a.X, a.ld, a.n, a.det ? nullptr : &perm, idx, 0, th, 0, false, true, Gscratch + (size_t)c * a.Pst, tc,
                       NNMaps{use_tma ? &tmX : nullptr, &tmTh, c});
    ll = nn_block_sum(ll, sm.red);
    float lp = 0.f;
    if (a.use_prior) {
        const float LOG_2PI = 1.8378770664093453f;
        for (int t = 0; t < 4; ++t) {
            int lo, hi;
            nn_tensor_range(dm, t, lo, hi);
            const float ssq = nn_sumsq(th + lo, hi - lo, sm.red);
            const float lam = th[dm.offl + t];
            lp += (float)(hi - lo) * (0.5f * logf(lam) - 0.5f * LOG_2PI) - 0.5f * lam * ssq - lam;
        }
    }
    if (threadIdx.x == 0) out[c] = (double)(lp + a.corr * ll);
    nn_tc_fini(tc);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*NNEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static NNEncodeFn nn_encode_fn() {
    static NNEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { cudaGetLastError(); return nullptr; }
        fn = (NNEncodeFn)p;
    }
    return fn;
}

// dataset map: 2-D (input column, row), box = 32 columns x 1 row (tile::gather4 fetches four such rows), 128-byte swizzle;
// the column extent is din, so the K tail past the last input (labels, padding) reads as zeros
static const CUtensorMap* nn_x_map(SoftmaxNetWork* w, const float* X, long long N, int ld) {
    for (int i = 0; i < w->n_maps; ++i)
        if (w->maps[i].base == X && w->maps[i].rows == N && w->maps[i].stride == ld) return &w->maps[i].map;
    NNEncodeFn fn = nn_encode_fn();
    if (!fn || !X) return nullptr;
    if (w->n_maps == 8) w->n_maps = 0;
    SoftmaxNetWork::MapEntry& e = w->maps[w->n_maps];
    const cuuint64_t gdim[2] = {(cuuint64_t)w->dm.din, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)S1_KC, 1};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return nullptr;
    e.base = X; e.rows = N; e.stride = ld;
    ++w->n_maps;
    return &e.map;
}

// parameter map: 3-D (input j, hidden unit k, chain) over the B block of a [count][stride] state array, box = 32 x dh x 1
static const CUtensorMap* nn_b_map(SoftmaxNetWork* w, const float* state, long long count, long long stride) {
    if (stride == 0 || count < 1) { count = 1; stride = 4 * ((w->dm.P + 3) / 4); }
    for (int i = 0; i < w->n_maps; ++i)
        if (w->maps[i].base == state && w->maps[i].rows == -count && w->maps[i].stride == stride) return &w->maps[i].map;
    NNEncodeFn fn = nn_encode_fn();
    if (!fn || !state) return nullptr;
    if (w->n_maps == 8) w->n_maps = 0;
    SoftmaxNetWork::MapEntry& e = w->maps[w->n_maps];
    const cuuint64_t gdim[3] = {(cuuint64_t)w->dm.din, (cuuint64_t)w->dm.dh, (cuuint64_t)count};
    const cuuint64_t gstride[2] = {(cuuint64_t)w->dm.din * 4, (cuuint64_t)stride * 4};
    const cuuint32_t box[3] = {(cuuint32_t)S1_KC, (cuuint32_t)w->dm.dh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(state) + w->dm.offB, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return nullptr;
    e.base = state; e.rows = -count; e.stride = stride;
    ++w->n_maps;
    return &e.map;
}

static bool nn_tma_enabled(const SoftmaxNetWork* w) {
    if (w->opt_stage1 == 1) return false;
    return (w->dm.offB % 4) == 0 && (size_t)nn_s1_ring_bytes(w->dm) <= (size_t)(2 * NN_RB * w->dm.dh + nn_tile_floats_all(w->dm)) * 4;
}

static bool nn_dims(NNDims& dm, int d_in, int d_hidden, int d_out) {
    if (d_in % 4 || d_hidden % 4 || d_hidden > NN_MAXH || d_out > NN_MAXO || d_out < 1 || (d_hidden * d_out) % 4) return false;
    dm.din = d_in; dm.dh = d_hidden; dm.dout = d_out; dm.dop = 4 * ((d_out + 3) / 4);
    dm.offA = 0; dm.offB = d_hidden * d_out; dm.offa = dm.offB + d_in * d_hidden; dm.offb = dm.offa + d_out; dm.offl = dm.offb + d_hidden;
    dm.P = dm.offl + 4;
    return true;
}

// host-only description of the launches of this family (sgm_debug_plan): no device is touched
int softmax_net_plan(int d_in, int d_hidden, int d_out, SoftmaxNetPlan* out) {
    NNDims dm;
    if (!nn_dims(dm, d_in, d_hidden, d_out)) return -1;
    out->threads = NN_THREADS;
    out->smem_bytes = nn_smem_bytes(dm);
    out->tmem_cols = 512;
    out->Pst = 4 * ((dm.P + 3) / 4);
    out->offB = dm.offB;
    SoftmaxNetWork w; w.dm = dm;
    out->tma_fed = nn_tma_enabled(&w) ? 1 : 0;
    out->s1_box_inner = S1_KC;
    out->s1_ring_bytes = (size_t)nn_s1_ring_bytes(dm);
    out->ring_capacity_bytes = (size_t)(2 * NN_RB * dm.dh + nn_tile_floats_all(dm)) * 4;
    return 0;
}

int softmax_net_create(SoftmaxNetWork** out, int d_in, int d_hidden, int d_out, int n, int chains, int device) {
    SoftmaxNetWork* w = new SoftmaxNetWork();
    NNDims& dm = w->dm;
    if (!nn_dims(dm, d_in, d_hidden, d_out)) { delete w; return (int)cudaErrorInvalidValue; }
    w->n = n; w->chains = chains; w->device = device;
    w->smem = nn_smem_bytes(dm);
    const int Pst = 4 * ((dm.P + 3) / 4);
    cudaError_t e = cudaMalloc(&w->G, (size_t)chains * Pst * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&w->Ghat, (size_t)chains * Pst * sizeof(float));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(softmax_net_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(softmax_net_fullgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(softmax_net_logpost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->smem);
    if (e != cudaSuccess) { softmax_net_destroy(w); return (int)e; }
    *out = w;
    return 0;
}

int softmax_net_set_option(SoftmaxNetWork* w, const char* key, int value) {
    if (!w || !key || value < 0 || value > 1) return -1;
    if (!strcmp(key, "nn_update")) { w->opt_update = value; return 0; }
    if (!strcmp(key, "nn_stage1")) { w->opt_stage1 = value; return 0; }
#ifdef NN_ENABLE_S5_TMEM
    if (!strcmp(key, "nn_stage5")) { w->opt_stage5 = value; return 0; }
#endif
    return -1;
}

void softmax_net_destroy(SoftmaxNetWork* w) {
    if (!w) return;
    if (w->G) cudaFree(w->G);
    if (w->Ghat) cudaFree(w->Ghat);
    if (w->part) cudaFree(w->part);
    delete w;
}

void softmax_net_tables(const SoftmaxNetWork* w, int* flat_of, int* tensor_of, int Pst) {
    const NNDims& dm = w->dm;
    for (int e = 0; e < Pst; ++e) {
        if (e >= dm.P) { flat_of[e] = -1; tensor_of[e] = 0; continue; }
        flat_of[e] = e;
        tensor_of[e] = e < dm.offB ? 0 : (e < dm.offa ? 1 : (e < dm.offb ? 2 : (e < dm.offl ? 3 : 4 + (e - dm.offl))));
    }
}

int softmax_net_launch(SoftmaxNetWork* w, const ChainArgs& a, cudaStream_t stream, int* launches) {
    cudaMemcpyToSymbolAsync(nn_prof, &a.prof, sizeof(a.prof), 0, cudaMemcpyHostToDevice, stream);
    {
        const int on = w->opt_update == 1 ? 0 : 1;
        cudaMemcpyToSymbolAsync(nn_stream_updates, &on, sizeof(on), 0, cudaMemcpyHostToDevice, stream);
        const int s5on = w->opt_stage5 == 1 ? 1 : 0;   // opt-in until it has run on a GPU
        cudaMemcpyToSymbolAsync(nn_s5_tmem, &s5on, sizeof(s5on), 0, cudaMemcpyHostToDevice, stream);
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
    const int grid = a.C < sms ? a.C : sms;
    static const CUtensorMap zero_map{};
    const CUtensorMap* mx = nn_tma_enabled(w) ? nn_x_map(w, a.X, a.N, a.ld) : nullptr;
    const CUtensorMap* mt = mx ? nn_b_map(w, a.theta, a.C, a.Pst) : nullptr;
    const CUtensorMap* mh = (mx && a.cv) ? nn_b_map(w, a.theta_hat, a.hat_stride ? a.C : 1, a.hat_stride) : mt;
    const int use_tma = (mx && mt && mh) ? 1 : 0;
    SGM_LAUNCH(softmax_net_chain_kernel, grid, NN_THREADS, w->smem, stream, a, w->dm, w->G, a.cv ? w->Ghat : nullptr, use_tma ? *mx : zero_map,
               use_tma ? *mt : zero_map, use_tma ? *mh : zero_map, use_tma);
    *launches = 1;
    return (int)cudaGetLastError();
}

int softmax_net_fullgrad(SoftmaxNetWork* w, const ChainArgs& a, int64_t row0, int64_t row1, int M, const float* theta_hat,
                         double* gsum, cudaStream_t stream) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
    const int64_t rows = row1 - row0;
    int gx = (int)((sms + M - 1) / M);
    if (gx < 1) gx = 1;
    int64_t rows_per = (rows + gx - 1) / gx;
    if (rows_per < NN_RB) rows_per = NN_RB;
    rows_per = ((rows_per + NN_RB - 1) / NN_RB) * NN_RB;
    gx = (int)((rows + rows_per - 1) / rows_per);
    if (gx < 1) gx = 1;
    const size_t need = (size_t)M * gx * a.Pst * sizeof(float);
    if (need > w->part_cap) {
        if (w->part) cudaFree(w->part);
        w->part = nullptr; w->part_cap = 0;
        cudaError_t e = cudaMalloc(&w->part, need);
        if (e != cudaSuccess) return (int)e;
        w->part_cap = need;
    }
    dim3 grid((unsigned)gx, (unsigned)M);
    static const CUtensorMap zero_map{};
    const CUtensorMap* mx = nn_tma_enabled(w) ? nn_x_map(w, a.X, a.N, a.ld) : nullptr;
    const CUtensorMap* mh = mx ? nn_b_map(w, theta_hat, M, a.Pst) : nullptr;
    const int use_tma = (mx && mh) ? 1 : 0;
    SGM_LAUNCH(softmax_net_fullgrad_kernel, grid, NN_THREADS, w->smem, stream, a, w->dm, row0, row1, (int)rows_per, theta_hat, w->part,
               use_tma ? *mx : zero_map, use_tma ? *mh : zero_map, use_tma);
    const int total = M * a.Pst;
    SGM_LAUNCH(softmax_net_part_reduce, (total + 255) / 256, 256, 0, stream, w->part, gx, a.Pst, gsum, total);
    return (int)cudaGetLastError();
}

int softmax_net_fullgrad_finish(SoftmaxNetWork* w, const ChainArgs& a, int M, const double* gsum, const float* theta_hat,
                                float* g_hat, float* theta, cudaStream_t stream) {
    SGM_LAUNCH(softmax_net_finish_kernel, M, NN_THREADS, 0, stream, a, w->dm, gsum, theta_hat, g_hat, theta);
    return (int)cudaGetLastError();
}

int softmax_net_logpost(SoftmaxNetWork* w, const ChainArgs& a, uint32_t word2, const float* theta, double* out, cudaStream_t stream) {
    static const CUtensorMap zero_map{};
    const CUtensorMap* mx = nn_tma_enabled(w) ? nn_x_map(w, a.X, a.N, a.ld) : nullptr;
    const CUtensorMap* mt = mx ? nn_b_map(w, theta, a.C, a.Pst) : nullptr;
    const int use_tma = (mx && mt) ? 1 : 0;
    SGM_LAUNCH(softmax_net_logpost_kernel, a.C, NN_THREADS, w->smem, stream, a, w->dm, word2, theta, w->G, out, use_tma ? *mx : zero_map,
               use_tma ? *mt : zero_map, use_tma);
    return (int)cudaGetLastError();
}

}  // namespace sgm
