// Softmax-net family (softmax_net.cu): the forward / backward stages S1..S5 of one minibatch row block and the gradient evaluation
// built from them -- tcgen05 helpers, S1 (X B) register-staged and TMA-fed, S2-S4 SIMT, S5 (X^T dZ1) on tcgen05 (and its
// tensor-memory-side variant, compiled only with -DNN_ENABLE_S5_TMEM).
#pragma once
#include "softmax_net_common.cuh"

namespace sgm {
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
// -DNN_ENABLE_S5_TMEM (SGM_BUILD_DEFS=-DNN_ENABLE_S5_TMEM python __graft_entry__.py --force), then selected by sgm_set_option "nn_stage5" = 1.
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

}  // namespace sgm
