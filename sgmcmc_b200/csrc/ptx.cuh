// Every inline-PTX statement of the sampler kernels, one wrapper per instruction: shared-space loads / stores, cp.async
// (LDGSTS), mbarrier, TMA (cp.async.bulk, cp.async.bulk.tensor tile and tile::gather4) and tcgen05 (TMEM allocation, UMMA,
// tensor-memory loads / stores).  sm_100a only.
//
// The kernels contain no `asm` of their own, so the same sources also compile for the CPU emulation the GPU-less test container
// uses (tests/emu, -DSGM_EMU: each wrapper then calls the emulated engine).  That build is test infrastructure; the product
// library is the nvcc build of these files and nothing else.
#pragma once
#include <cstdint>
#include <cuda.h>   // CUtensorMap
#include <cuda_runtime.h>

// kernel launch and dynamic shared memory, spelled so that both builds read the same call sites
#ifdef SGM_EMU
#define SGM_LAUNCH(kern, grid, block, smem, stream, ...) \
    ::emu::launch(#kern, (const void*)(kern), dim3(grid), dim3(block), (size_t)(smem), [=]() { (kern)(__VA_ARGS__); })
#define SGM_DYN_SMEM(type, name, align) type* name = reinterpret_cast<type*>(::emu::dyn_smem(align))
#define SGM_CLUSTER_OCCUPANCY(err, n_out, kern, grid_, block_, smem_, cluster_) do { *(n_out) = 2; (err) = cudaSuccess; } while (0)
#define SGM_LAUNCH_CLUSTER(err, kern, grid, block, smem, stream, cluster, ...)                                                           \
    do { ::emu::launch_cluster(#kern, (const void*)(kern), dim3(grid), dim3(block), (size_t)(smem), (unsigned)(cluster), [=]() { (kern)(__VA_ARGS__); }); (err) = cudaSuccess; } while (0)
#else
#define SGM_LAUNCH(kern, grid, block, smem, stream, ...) (kern)<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SGM_DYN_SMEM(type, name, align) extern __shared__ __align__(align) type name[]
// how many clusters of `cluster_` CTAs of this kernel the device can hold at once (0: it cannot co-schedule one)
#define SGM_CLUSTER_OCCUPANCY(err, n_out, kern, grid_, block_, smem_, cluster_)                                                          \
    do {                                                                                                                                 \
        cudaLaunchConfig_t cfg_ = {};                                                                                                    \
        cfg_.gridDim = dim3(grid_); cfg_.blockDim = dim3(block_); cfg_.dynamicSmemBytes = (size_t)(smem_);                               \
        cudaLaunchAttribute attr_[1];                                                                                                    \
        attr_[0].id = cudaLaunchAttributeClusterDimension;                                                                               \
        attr_[0].val.clusterDim.x = (unsigned)(cluster_); attr_[0].val.clusterDim.y = 1; attr_[0].val.clusterDim.z = 1;                  \
        cfg_.attrs = attr_; cfg_.numAttrs = 1;                                                                                           \
        (err) = cudaOccupancyMaxActiveClusters((n_out), (kern), &cfg_);                                                                  \
    } while (0)
// thread-block clusters of `cluster` CTAs along x; `err` receives the launch status
#define SGM_LAUNCH_CLUSTER(err, kern, grid_, block_, smem_, stream_, cluster_, ...)                                                      \
    do {                                                                                                                                 \
        cudaLaunchConfig_t cfg_ = {};                                                                                                    \
        cfg_.gridDim = dim3(grid_); cfg_.blockDim = dim3(block_); cfg_.dynamicSmemBytes = (size_t)(smem_); cfg_.stream = (stream_);      \
        cudaLaunchAttribute attr_[1];                                                                                                    \
        attr_[0].id = cudaLaunchAttributeClusterDimension;                                                                               \
        attr_[0].val.clusterDim.x = (unsigned)(cluster_); attr_[0].val.clusterDim.y = 1; attr_[0].val.clusterDim.z = 1;                  \
        cfg_.attrs = attr_; cfg_.numAttrs = 1;                                                                                           \
        (err) = cudaLaunchKernelEx(&cfg_, (kern), __VA_ARGS__);                                                                          \
    } while (0)
#endif

#ifdef SGM_EMU
#define SGM_NOINLINE __attribute__((noinline))
#define SGM_EMU_NOTE(name) do { if (threadIdx.x == 0) ::emu::note(name); } while (0)   // "this path ran" counters for the emulated tests
#else
#define SGM_NOINLINE __noinline__
#define SGM_EMU_NOTE(name) do { } while (0)
#endif

namespace sgm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#ifndef SGM_EMU
// ---- shared-space accessors: generic pointers into shared memory cost an address-window conversion per access and keep ptxas
// from batching the loads (profiles/r01_c3_staged_v4.txt)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- cp.async (LDGSTS), 16 bytes
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {   // one arrival + the transaction bytes to wait for
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA, UMMA operand fetch)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA
// 1-D bulk copy global -> shared::cta, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 1-D bulk store shared::cta -> global, tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }   // sources may be reused
template <int N> __device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }             // writes are complete
// this CTA's generic-proxy accesses to GLOBAL memory <-> its async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// tile::gather4: FOUR rows of a 2-D tensor (box = cols x 1) per instruction, landing contiguously (SASS: UTMALDG.2D.GATHER4)
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}
__device__ __forceinline__ void tma_box2(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_box3(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---- tcgen05 (cta_group::1)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot) {   // whole warp; the base address lands in shared memory at `slot`
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(NCOLS));
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS)); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[128 x N] (+)= A[128 x 8] B[N x 8]^T, TF32 operands, float32 accumulators in TMEM; both operands from shared memory
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// A operand read from tensor memory (lane = row, one TF32 element per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// one arrival on `bar` once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
// 16 consecutive columns of this thread's TMEM lane; the caller issues tmem_wait_ld() before using the values
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- thread-block clusters: rank, the cluster-wide barrier (split arrive / wait, release / acquire at cluster scope) and loads
// from a peer CTA's shared memory (distributed shared memory)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {   // the same shared-memory offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float ld_dsmem32(uint32_t addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

#else   // ---- SGM_EMU: the same interface on the emulated engines (tests/emu/include/emu_async.h) ---------------------------------

__device__ __forceinline__ float4 lds128(uint32_t addr) { float4 v; std::memcpy(&v, ::emu::smem_ptr(addr, 16, 16), 16); return v; }
__device__ __forceinline__ float lds32(uint32_t addr) { float v; std::memcpy(&v, ::emu::smem_ptr(addr, 4, 4), 4); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; std::memcpy(&v, ::emu::smem_ptr(addr, 4, 4), 4); return v; }
__device__ __forceinline__ void sts32(uint32_t addr, float v) { std::memcpy(::emu::smem_ptr(addr, 4, 4), &v, 4); }
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v) { std::memcpy(::emu::smem_ptr(addr, 16, 16), &v, 16); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { ::emu::cp_async16(dst, src); }
__device__ __forceinline__ void cp_async_commit() { ::emu::cp_async_commit(); }
__device__ __forceinline__ void cp_async_wait_all() { ::emu::cp_async_wait_group(0); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { ::emu::mbar_init(bar, count); }
__device__ __forceinline__ void fence_mbar_init() {}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { ::emu::mbar_arrive_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { ::emu::mbar_arrive(bar); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { ::emu::mbar_wait(bar, parity); }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) { ::emu::bulk_g2s(dst, src, bytes, bar); }
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) { ::emu::bulk_s2g(dst, src, bytes); }
__device__ __forceinline__ void bulk_commit_group() { ::emu::bulk_commit_group(); }
template <int N> __device__ __forceinline__ void bulk_wait_group_read() { ::emu::bulk_wait_group(N, true); }
template <int N> __device__ __forceinline__ void bulk_wait_group() { ::emu::bulk_wait_group(N, false); }
__device__ __forceinline__ void fence_proxy_async_global() {}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    ::emu::tma_gather4(dst, map, col, r0, r1, r2, r3, bar);
}
__device__ __forceinline__ void tma_box2(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) { const int c[2] = {c0, c1}; ::emu::tma_tile(dst, map, c, 2, bar); }
__device__ __forceinline__ void tma_box3(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) { const int c[3] = {c0, c1, c2}; ::emu::tma_tile(dst, map, c, 3, bar); }
template <int NCOLS> __device__ __forceinline__ void tmem_alloc(uint32_t slot) { ::emu::tmem_alloc(slot, NCOLS); }
__device__ __forceinline__ void tmem_relinquish() { ::emu::tmem_relinquish(); }
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) { ::emu::tmem_dealloc(taddr, NCOLS); }
__device__ __forceinline__ void tc_fence_before() {}
__device__ __forceinline__ void tc_fence_after() {}
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) { ::emu::umma_tf32_ss(tmem_d, adesc, bdesc, idesc, accumulate); }
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) { ::emu::umma_tf32_ts(tmem_d, tmem_a, bdesc, idesc, accumulate); }
__device__ __forceinline__ void umma_commit(uint32_t bar) { ::emu::umma_commit(bar); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) { uint32_t r[16]; std::memcpy(r, v, 64); ::emu::tmem_st_32x32b(taddr, r, 16); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) { ::emu::tmem_ld_32x32b(taddr, r, 16); }
__device__ __forceinline__ void tmem_wait_ld() {}
__device__ __forceinline__ void tmem_wait_st() {}
__device__ __forceinline__ bool elect_one() { return ::emu::elect_one(); }
__device__ __forceinline__ uint32_t cluster_ctarank() { return ::emu::cluster_ctarank(); }
__device__ __forceinline__ uint32_t cluster_nctarank() { return ::emu::cluster_nctarank(); }
__device__ __forceinline__ void cluster_arrive() { ::emu::cluster_arrive(); }
__device__ __forceinline__ void cluster_wait() { ::emu::cluster_wait(); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { return ::emu::mapa(addr, rank); }
__device__ __forceinline__ float ld_dsmem32(uint32_t addr) { float v; std::memcpy(&v, ::emu::dsmem_ptr(addr, 4, 4), 4); return v; }
#endif

}  // namespace sgm
