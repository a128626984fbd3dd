// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): the asynchronous engines the kernels drive through inline PTX --
// cp.async (LDGSTS), mbarrier, TMA (cp.async.bulk / cp.async.bulk.tensor tile + tile::gather4) and tcgen05 (TMEM, UMMA).
// The kernel sources keep every PTX statement inside a small wrapper function; under -DSGM_EMU the wrapper calls one of these.
//
// Timing model, chosen by SGM_EMU_ASYNC (default "lazy"; tests run both):
//   lazy   an asynchronous operation takes effect at the LAST moment the memory model allows -- when a thread observes its
//          completion (cp.async.wait_group, an mbarrier test that the operation's completion feeds);
//   eager  it takes effect at issue.
// A read that is missing its wait sees stale data under "lazy"; a buffer that is refilled before its readers are done is
// clobbered under "eager".
#pragma once
#include <cstdint>
#include "cuda.h"

namespace emu {

// what the emulated cuTensorMapEncodeTiled stores inside CUtensorMap::opaque (validated like the driver validates it)
struct TensorMapRec {
    uint64_t magic;
    const unsigned char* base;
    uint8_t rank, elem_bytes, swizzle, pad_;
    uint64_t dim[5];
    uint64_t stride[5];   // bytes; stride[0] = elem_bytes
    uint32_t box[5];
};
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* gdim, const cuuint64_t* gstride,
                      const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion l2,
                      CUtensorMapFloatOOBfill oob);

// ---- cp.async (per-thread groups)
void cp_async16(uint32_t dst, const void* src);
void cp_async_commit();
void cp_async_wait_group(int allowed_pending);

// ---- mbarrier (64-bit object in shared memory, addressed by its shared-window address)
void mbar_init(uint32_t bar, uint32_t count);
void mbar_arrive(uint32_t bar);
void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes);
void mbar_expect_tx(uint32_t bar, uint32_t bytes);
bool mbar_try_wait(uint32_t bar, uint32_t parity);
void mbar_wait(uint32_t bar, uint32_t parity);   // spins (yielding) until the phase with this parity has completed

// ---- TMA
void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar);
void bulk_s2g(void* dst, uint32_t src, uint32_t bytes);          // cp.async.bulk.global.shared::cta.bulk_group
void bulk_commit_group();
void bulk_wait_group(int allowed_pending, bool read_only);
void tma_tile(uint32_t dst, const CUtensorMap* map, const int* coord, int ncoord, uint32_t bar);
void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar);
inline void fence_proxy_async() {}
inline void fence_mbarrier_init() {}

// ---- tcgen05 (cta_group::1)
void tmem_alloc(uint32_t slot_smem, uint32_t ncols);   // warp-wide; writes the base address to shared memory
void tmem_relinquish();
void tmem_dealloc(uint32_t taddr, uint32_t ncols);
void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate);
void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate);
void umma_commit(uint32_t bar);
void tmem_ld_32x32b(uint32_t taddr, uint32_t* regs, int n);   // n consecutive columns of this thread's TMEM lane
void tmem_st_32x32b(uint32_t taddr, const uint32_t* regs, int n);
inline void tmem_wait_ld() {}
inline void tmem_wait_st() {}
inline void tcgen05_fence() {}
bool elect_one();

}  // namespace emu
