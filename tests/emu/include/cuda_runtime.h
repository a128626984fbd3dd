// TEST INFRASTRUCTURE ONLY -- never part of the product library.
//
// A functional CPU emulation of the CUDA subset that sgmcmc_b200/csrc uses, so that the SAME kernel sources can be compiled with
// g++ (-DSGM_EMU, this directory first on the include path) and checked against the oracle in a container without a GPU
// (tests/emu/README.md).  The product library (sgmcmc_b200/lib/libsgmcmc_b200.so) is built by nvcc only and fails with SGM_ECUDA
// when there is no CUDA device; nothing under sgmcmc_b200/ or bench.py loads the emulated build.
//
// Execution model: one OS thread; the threads of a CTA are ucontext fibers scheduled round-robin, CTAs run one after another.
// Every blocking primitive (__syncthreads, warp shuffles, mbarrier waits) yields; if a whole pass over the CTA makes no progress
// the launch is reported as a DEADLOCK instead of hanging.  Asynchronous engines (cp.async, TMA, tcgen05.mma) are modelled
// "adversarially lazy": their effects become visible at the LAST moment the PTX memory model allows (when a thread observes the
// completion through wait_group / an mbarrier phase), so a missing wait reads stale data here too.
#pragma once
#ifndef SGM_EMU
#error "tests/emu/include is only for the emulated test build (-DSGM_EMU)"
#endif
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---- qualifiers ----------------------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static   // CTAs run one after another on one OS thread: one static copy is the CTA's copy

// ---- vector types --------------------------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

// ---- emulator core (tests/emu/emu.cpp) -----------------------------------------------------------------------------------------
namespace emu {
extern uint3 threadIdx_, blockIdx_;
extern dim3 blockDim_, gridDim_;

void launch(const char* name, const void* fn, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
// thread-block clusters of cluster_x CTAs along x (cudaLaunchKernelEx + cudaLaunchAttributeClusterDimension on the device)
void launch_cluster(const char* name, const void* fn, dim3 grid, dim3 block, size_t smem, unsigned cluster_x, const std::function<void()>& body);
unsigned cluster_ctarank();
unsigned cluster_nctarank();
void cluster_arrive();             // barrier.cluster.arrive.release
void cluster_wait();               // barrier.cluster.wait.acquire
uint32_t mapa(uint32_t addr, uint32_t rank);                           // mapa.shared::cluster
unsigned char* dsmem_ptr(uint32_t addr, size_t bytes, size_t align);   // checked shared::cluster address -> host pointer
void yield();                      // back to the scheduler (the calling fiber stays runnable)
void progress();                   // "something changed": resets the deadlock detector
[[noreturn]] void fault(const char* fmt, ...);   // device-side fault (illegal address / instruction): aborts the launch, sticky error
void syncthreads();
void syncwarp(unsigned mask);
uint64_t shfl(unsigned mask, uint64_t v, int src_lane);   // generic exchange: returns the value lane src_lane deposited
int lane_id();
int warp_id();
int linear_tid();
// dynamic shared memory of the running CTA
unsigned char* dyn_smem(size_t align);
uint32_t smem_window(const void* p);           // generic pointer into the CTA's dynamic shared memory -> shared-window address
unsigned char* smem_ptr(uint32_t addr, size_t bytes, size_t align);   // checked shared-window address -> host pointer
// global memory registry (cudaMalloc'ed ranges, guard pages at both ends)
void check_global(const void* p, size_t bytes, size_t align, const char* what);
long long clock();
void note(const char* name);   // counts executions of a named code path (SGM_EMU_NOTE)
}  // namespace emu

#define threadIdx (::emu::threadIdx_)
#define blockIdx (::emu::blockIdx_)
#define blockDim (::emu::blockDim_)
#define gridDim (::emu::gridDim_)

// ---- device intrinsics ---------------------------------------------------------------------------------------------------------
static inline void __syncthreads() { ::emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { ::emu::syncwarp(mask); }
template <class T> static inline T __ldg(const T* p) { return *p; }

namespace emu {
template <class T> static inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle payload"); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = ::emu::lane_id();
    const int s = (lane & ~(width - 1)) | (src & (width - 1));
    return ::emu::from_bits<T>(::emu::shfl(mask, ::emu::to_bits(v), s));
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    const int lane = ::emu::lane_id();
    int s = lane ^ lanemask;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return ::emu::from_bits<T>(::emu::shfl(mask, ::emu::to_bits(v), s));
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = ::emu::lane_id();
    int s = lane + (int)delta;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return ::emu::from_bits<T>(::emu::shfl(mask, ::emu::to_bits(v), s));
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = ::emu::lane_id();
    int s = lane - (int)delta;
    if (s < (lane & ~(width - 1))) s = lane;
    return ::emu::from_bits<T>(::emu::shfl(mask, ::emu::to_bits(v), s));
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if ((mask >> l) & 1u) r |= (::emu::shfl(mask, (uint64_t)(pred != 0), l) ? 1u : 0u) << l;
    return r;
}

// single OS thread: plain read-modify-write is atomic
template <class T> static inline T atomicAdd(T* p, T v) { const T old = *p; *p = old + v; return old; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
static inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { const unsigned o = *p; if (o == cmp) *p = v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
// glibc declares __logf & co. as its own internal aliases: map the CUDA fast intrinsics by macro
#define __logf(x) logf(x)
#define __expf(x) expf(x)
#define __cosf(x) cosf(x)
#define __sinf(x) sinf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline long long clock64() { return ::emu::clock(); }
static inline size_t __cvta_generic_to_shared(const void* p) { return ::emu::smem_window(p); }
using std::isfinite;
using std::isnan;

// ---- runtime API subset ----------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum : int {
    cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorLaunchFailure = 719, cudaErrorIllegalAddress = 700,
    cudaErrorIllegalInstruction = 715, cudaErrorLaunchTimeout = 702, cudaErrorNoDevice = 100
};
struct EmuStream; struct EmuEvent;
typedef EmuStream* cudaStream_t;
typedef EmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEnableDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };

cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int dev);
cudaError_t cudaMallocImpl(void** p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMallocImpl((void**)p, bytes); }
cudaError_t cudaFree(void* p);
cudaError_t cudaHostAllocImpl(void** p, size_t bytes, unsigned flags);
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned flags) { return cudaHostAllocImpl((void**)p, bytes, flags); }
cudaError_t cudaFreeHost(void* p);
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void* dst, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamCreate(cudaStream_t* s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaFuncSetAttributeImpl(const void* fn, cudaFuncAttribute a, int v);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F fn, cudaFuncAttribute a, int v) { return cudaFuncSetAttributeImpl((const void*)fn, a, v); }
cudaError_t cudaOccupancyImpl(int* n, const void* fn, int threads, size_t smem);
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F fn, int threads, size_t smem) {
    return cudaOccupancyImpl(n, (const void*)fn, threads, smem);
}
cudaError_t cudaGetDriverEntryPoint(const char* sym, void** fn, unsigned long long flags, cudaDriverEntryPointQueryResult* q = nullptr);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
// __device__ variables are plain globals here
template <class T> static inline cudaError_t cudaMemcpyToSymbolAsync(T& sym, const void* src, size_t bytes, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice, cudaStream_t = nullptr) {
    std::memcpy((char*)&sym + off, src, bytes); return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t bytes, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) {
    std::memcpy((char*)&sym + off, src, bytes); return cudaSuccess;
}

#include "emu_async.h"
