// TEST INFRASTRUCTURE ONLY: deliberately broken kernels, to show that the emulation REPORTS the failure classes it exists to
// catch (on a device these are Xid faults, hangs or silently wrong numbers).  Each entry point runs one kernel and returns the
// emulator's message ("" when nothing was detected); tests/test_emu_selftest.py asserts on the messages.
#define SGM_EMU 1
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../sgmcmc_b200/csrc/ptx.cuh"

using namespace sgm;
extern "C" void emu_reset_error();
extern "C" const char* emu_last_message();

namespace {

// ---- 1. a consumer that reads a TMA destination without waiting for the mbarrier
__global__ void k_missing_wait(const float* src, float* out, int wait) {
    SGM_DYN_SMEM(unsigned char, smem, 128);
    const uint32_t buf = smem_u32(smem), bar = buf + 1024;
    float* tile = reinterpret_cast<float*>(smem);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tile[i] = -1.f;
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(bar, 1024); bulk_g2s(buf, src, 1024, bar); }
    if (wait) mbar_wait(bar, 0);
    out[threadIdx.x] = tile[threadIdx.x];
}

// ---- 2. two warps that wait for each other
__global__ void k_deadlock() {
    SGM_DYN_SMEM(unsigned char, smem, 16);
    const uint32_t bar = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x < 32) { mbar_wait(bar, 0); if (threadIdx.x == 0) mbar_arrive(bar + 8); }
    else { mbar_wait(bar + 8, 0); if (threadIdx.x == 32) mbar_arrive(bar); }
}

// ---- 3. global read one element past the allocation
__global__ void k_oob_global(const float* src, float* out, int n) { out[threadIdx.x] = src[n + threadIdx.x]; }

// ---- 4. shared-memory store past the dynamic allocation
__global__ void k_oob_shared(int n_floats) {
    SGM_DYN_SMEM(unsigned char, smem, 16);
    sts32(smem_u32(smem) + 4u * (uint32_t)(n_floats + threadIdx.x), 1.f);
}

// ---- 5. TMA tile with an innermost coordinate that is not 16-byte aligned (the fault of round 1's 8-GPU run)
__global__ void k_tma_unaligned(const __grid_constant__ CUtensorMap map, int c0) {
    SGM_DYN_SMEM(unsigned char, smem, 16);
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u, bar = base + 128 * 128;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1); fence_mbar_init();
        mbar_expect_tx(bar, 128 * 128);
        tma_box2(base, &map, c0, 0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
}

// ---- 6. __syncthreads that only half of the CTA reaches while the other half waits on something else
__global__ void k_divergent_barrier() {
    SGM_DYN_SMEM(unsigned char, smem, 16);
    const uint32_t bar = smem_u32(smem);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x < 32) __syncthreads();
    else mbar_wait(bar, 0);
}

// ---- 7. a stage that is refilled by TMA while a slow reader still uses it (visible in the EAGER timing model)
__global__ void k_refill_race(const float* src, float* out) {
    SGM_DYN_SMEM(unsigned char, smem, 128);
    const uint32_t buf = smem_u32(smem), bar = buf + 1024;
    float* tile = reinterpret_cast<float*>(smem);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); mbar_expect_tx(bar, 1024); bulk_g2s(buf, src, 1024, bar); }
    __syncthreads();
    mbar_wait(bar, 0);
    const float first = tile[threadIdx.x];
    // BUG: no barrier between the readers and the refill
    if (threadIdx.x == 0) { mbar_expect_tx(bar, 1024); bulk_g2s(buf, src + 256, 1024, bar); }
    __syncwarp();
    out[threadIdx.x] = first + tile[(threadIdx.x + 1) % 64];   // expects the FIRST tile's neighbour
    mbar_wait(bar, 1);
}

// ---- 8. thread-block clusters: every CTA publishes a value in its shared memory and sums its peers' values through distributed
// shared memory.  mode 0: correct (barrier before the reads, barrier before exit); 1: no barrier before the reads (a peer's slot is
// still poison); 2: no barrier before exit (a CTA that is done leaves while a peer has yet to read its slot)
__global__ void k_cluster_exchange(float* out, int mode) {
    SGM_DYN_SMEM(float, smem, 16);
    const uint32_t slot = smem_u32(smem);
    const uint32_t S = cluster_nctarank(), rank = cluster_ctarank();
    if (threadIdx.x == 0) smem[0] = (float)(rank + 1);
    if (mode != 1) { cluster_arrive(); cluster_wait(); } else __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (uint32_t p = 0; p < S; ++p) s += ld_dsmem32(mapa(slot, p));
        out[blockIdx.x] = s;
    }
    if (mode != 2) { cluster_arrive(); cluster_wait(); }
}

const char* finish() {
    static char buf[1024];
    cudaError_t e = cudaDeviceSynchronize();
    (void)e;
    snprintf(buf, sizeof buf, "%s", emu_last_message());
    return buf;
}

}  // namespace

extern "C" {

const char* selftest_missing_wait(int wait, float* first_value) {
    emu_reset_error();
    float *src, *out;
    cudaMalloc(&src, 1024); cudaMalloc(&out, 64 * 4);
    float h[256]; for (int i = 0; i < 256; ++i) h[i] = (float)i;
    cudaMemcpy(src, h, 1024, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_missing_wait, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048);
    SGM_LAUNCH(k_missing_wait, 1, 64, 2048, 0, src, out, wait);
    float o[64]; cudaMemcpy(o, out, 256, cudaMemcpyDeviceToHost);
    *first_value = o[5];
    const char* m = finish();
    cudaFree(src); cudaFree(out);
    return m;
}
const char* selftest_deadlock() { emu_reset_error(); SGM_LAUNCH(k_deadlock, 1, 64, 64, 0); return finish(); }
const char* selftest_oob_global() {
    emu_reset_error();
    float *src, *out; cudaMalloc(&src, 100 * 4); cudaMalloc(&out, 32 * 4);
    SGM_LAUNCH(k_oob_global, 1, 32, 0, 0, src, out, 100);
    const char* m = finish();
    return m;   // the context is gone, like after a device fault: no frees
}
const char* selftest_oob_shared() { emu_reset_error(); SGM_LAUNCH(k_oob_shared, 1, 32, 256, 0, 64); return finish(); }
const char* selftest_tma_unaligned(int c0) {
    emu_reset_error();
    float* g; cudaMalloc(&g, 256 * 256 * 4);
    cudaMemset(g, 0, 256 * 256 * 4);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault);
    CUtensorMap map;
    const cuuint64_t gdim[2] = {256, 256}, gstride[1] = {256 * 4};
    const cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
    if (((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return "encode failed";
    cudaFuncSetAttribute(k_tma_unaligned, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    SGM_LAUNCH(k_tma_unaligned, 1, 32, 32 * 1024, 0, map, c0);
    const char* m = finish();
    if (!*m) cudaFree(g);
    return m;
}
// out[0] = the sum CTA 0 computed (1 + 2 + ... + cluster size when the protocol is right)
const char* selftest_cluster_exchange(int cluster, int mode, float* sum0) {
    emu_reset_error();
    float* out; cudaMalloc(&out, 16 * 4);
    cudaError_t err;
    SGM_LAUNCH_CLUSTER(err, k_cluster_exchange, 2 * cluster, 64, 64, 0, cluster, out, mode);
    (void)err;
    const char* m = finish();
    float o[16] = {0};
    if (!*m) { cudaMemcpy(o, out, 2 * cluster * 4, cudaMemcpyDeviceToHost); cudaFree(out); }
    *sum0 = o[0];
    return m;
}
const char* selftest_divergent_barrier() { emu_reset_error(); SGM_LAUNCH(k_divergent_barrier, 1, 64, 64, 0); return finish(); }
// returns the number of threads that saw a clobbered neighbour
int selftest_refill_race() {
    emu_reset_error();
    float *src, *out; cudaMalloc(&src, 2048); cudaMalloc(&out, 64 * 4);
    float h[512]; for (int i = 0; i < 512; ++i) h[i] = (float)i;
    cudaMemcpy(src, h, 2048, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_refill_race, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048);
    SGM_LAUNCH(k_refill_race, 1, 64, 2048, 0, src, out);
    float o[64]; cudaMemcpy(o, out, 256, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 64; ++t) if (o[t] != (float)t + (float)((t + 1) % 64)) ++bad;
    cudaFree(src); cudaFree(out);
    return bad;
}
// tensor-map validation: returns 0 when the encoder REJECTS the description
int selftest_encode(int which) {
    emu_reset_error();
    float* g; cudaMalloc(&g, 64 * 64 * 4);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault);
    CUtensorMap map;
    cuuint64_t gdim[2] = {64, 64}, gstride[1] = {64 * 4};
    cuuint32_t box[2] = {32, 8}, estr[2] = {1, 1};
    void* base = g;
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    switch (which) {
        case 0: break;                                    // valid
        case 1: gstride[0] = 63 * 4; break;               // stride not a multiple of 16 bytes
        case 2: base = (char*)g + 4; break;               // base not 16-byte aligned
        case 3: box[1] = 300; break;                      // box dimension > 256
        case 4: gdim[1] = 65; break;                      // tensor larger than the allocation
        case 5: box[0] = 64; break;                       // 256-byte inner box with a 128-byte swizzle
        case 6: box[0] = 3; sw = CU_TENSOR_MAP_SWIZZLE_NONE; break;   // inner box not a multiple of 16 bytes
    }
    const CUresult rc = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFree(g);
    return rc == CUDA_SUCCESS ? 1 : 0;
}
}
