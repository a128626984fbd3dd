// Probe: D[128 x N] = A[128 x K] * B[N x K]^T on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM), operands staged by ordinary threads in the no-swizzle K-major core-matrix layout, with the 3xTF32 split
// (hi*hi + hi*lo + lo*hi) that keeps float32-level accuracy.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tcgen05_probe tools/tcgen05_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 112, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: 16-byte chunk (kg, row) at kg * (rows * 16) + row * 16; LBO = rows * 16, SBO = 128
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t rows) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((rows * 16u) >> 4 & 0x3FFF) << 16;   // leading byte offset (between the two 16-byte K chunks)
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;          // stride byte offset (between 8-row core matrices)
    d |= (uint64_t)1 << 46;                               // descriptor version 1 (Blackwell)
    return d;
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, int split, int reps, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Ahi = reinterpret_cast<float*>(smem);        // [K/4][M][4]
    float* Alo = Ahi + M * K;
    float* Bhi = Alo + M * K;                            // [K/4][N][4]
    float* Blo = Bhi + N * K;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        const float x = A[i];
        const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        Ahi[(k / 4) * (M * 4) + r * 4 + (k % 4)] = hi;
        Alo[(k / 4) * (M * 4) + r * 4 + (k % 4)] = x - hi;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        const float x = B[i];
        const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        Bhi[(k / 4) * (N * 4) + r * 4 + (k % 4)] = hi;
        Blo[(k / 4) * (N * 4) + r * 4 + (k % 4)] = x - hi;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        int first = 1;
        const long long t_start = clock64();
        for (int rep = 0; rep < reps; ++rep)
        for (int k8 = 0; k8 < K / 8; ++k8) {
            const uint32_t aoff = (uint32_t)(2 * k8) * (M * 16), boff = (uint32_t)(2 * k8) * (N * 16);
            const uint64_t ahi = make_desc(smem_u32(Ahi) + aoff, M), alo = make_desc(smem_u32(Alo) + aoff, M);
            const uint64_t bhi = make_desc(smem_u32(Bhi) + boff, N), blo = make_desc(smem_u32(Blo) + boff, N);
            const uint64_t as[3] = {ahi, ahi, alo}, bs[3] = {bhi, blo, bhi};
            for (int t = 0; t < (split ? 3 : 1); ++t) {
                const uint32_t accum = first ? 0u : 1u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr), "l"(as[t]), "l"(bs[t]), "r"(idesc), "r"(accum)
                    : "memory");
                first = 0;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t_issued = clock64();
        asm volatile("{\n\t.reg .pred p;\n\tW2_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2_%=;\n\tbra W2_%=;\n\tD2_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        const long long t_done = clock64();
        if (cycles) { cycles[0] = t_issued - t_start; cycles[1] = t_done - t_start; }
    }
    // everybody waits for the MMAs
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // TMEM -> registers: warp w reads lanes 32w..32w+31 (= rows), 16 columns at a time
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t addr = taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(32 * warp + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(128));
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(1);
    for (auto& x : A) x = (float)rand() / RAND_MAX - 0.5f;
    for (auto& x : B) x = (float)rand() / RAND_MAX - 0.5f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int split = 0; split < 2; ++split) {
        cudaMemset(dD, 0, D.size() * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, split, 1, nullptr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0;
                for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
                maxerr = fmax(maxerr, fabs(r - D[m * N + n])); maxref = fmax(maxref, fabs(r));
            }
        printf("split=%d  max|err|=%.3e  max|ref|=%.3e  rel=%.3e\n", split, maxerr, maxref, maxerr / maxref);
    }
    long long* dc; cudaMalloc(&dc, 16);
    for (int reps : {1, 8, 64}) {
        probe<<<1, 128, smem>>>(dA, dB, dD, 1, reps, dc);
        cudaDeviceSynchronize();
        long long hc[2]; cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
        const int mmas = reps * (K / 8) * 3;
        printf("reps=%d: %d MMAs (128x%dx8 tf32): issue %lld cycles (%.1f / MMA), done %lld cycles (%.1f / MMA)\n", reps, mmas, N, hc[0], (double)hc[0] / mmas, hc[1], (double)hc[1] / mmas);
    }
    return 0;
}
