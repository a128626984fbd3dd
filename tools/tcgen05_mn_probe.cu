// Probe for the NEXT softmax-net step: can a TMA tile [K rows][32 floats] (128-byte swizzle), i.e. the
// tiles S1 already loads with gather4, be used directly as an MN-MAJOR B operand of tcgen05.mma kind::tf32?
//   D[128 x N] = A[128 x K] * Bm[K x N]   with Bm stored row-major (N contiguous) in N-blocks of 32 floats:
//   shared layout of block nb: row k at nb * (K * 128) + k * 128, 16-byte chunk c of the row at position c ^ (k & 7)   (what TMA SW128 writes)
// Canonical MN-major SW128 (cute/atom/mma_traits_sm100.hpp): ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO)), T = 4 floats:
//   one atom = 32 MN elements (128 B) x 8 K rows (128 B apart) = 1024 B; LBO = stride between MN atoms, SBO = stride between K atoms.
// The probe tries both assignments of (LBO, SBO) and prints the error of each against a float64 product (single TF32 pass:
// expect ~1e-3 relative when the layout is right, O(1) when it is wrong).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tcgen05_mn_probe tools/tcgen05_mn_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* Bm, float* D, int variant) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    float* As = reinterpret_cast<float*>(sm);                  // K-major SW128: row m at m * 128, chunk c at c ^ (m & 7)   (K = 32 floats = one span)
    float* Bs = reinterpret_cast<float*>(sm + M * 128);        // N / 32 blocks of [K rows][128 B]
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128) {
        const int m = i / K, k = i % K;
        As[m * 32 + (((k >> 2) ^ (m & 7)) << 2) + (k & 3)] = A[i];
    }
    for (int i = tid; i < K * N; i += 128) {
        const int k = i / N, n = i % N, nb = n >> 5, c = (n & 31) >> 2;
        Bs[nb * (K * 32) + k * 32 + ((c ^ (k & 7)) << 2) + (n & 3)] = Bm[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base;
    if (tid == 0) {
        // kind::tf32, f32 accumulate, A K-major (bit 15 = 0), B MN-major (bit 16 = 1)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t blockB = K * 128;   // bytes between the 32-wide N blocks
        for (int k8 = 0; k8 < K / 8; ++k8) {
            const uint64_t ad = desc_sw128(smem_u32(As) + 32u * k8, 16, 1024);                 // K-major: +32 B per K step, SBO = 1024
            const uint32_t bstart = smem_u32(Bs) + 1024u * k8;                                   // MN-major: +8 rows per K step
            const uint64_t bd = variant == 0 ? desc_sw128(bstart, blockB, 1024) : desc_sw128(bstart, 1024, blockB);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(taddr), "l"(ad), "l"(bd), "r"(idesc), "r"(k8 > 0 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(64));
}

int main() {
    std::vector<float> A(M * K), B(K * N), D(M * N);
    srand(2);
    for (auto& x : A) x = (float)rand() / RAND_MAX - 0.5f;
    for (auto& x : B) x = (float)rand() / RAND_MAX - 0.5f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)M * 128 + (size_t)(N / 32) * K * 128 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dD, 0, D.size() * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error: %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0;
                for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[k * N + n];
                maxerr = fmax(maxerr, fabs(r - D[m * N + n])); maxref = fmax(maxref, fabs(r));
            }
        printf("variant %d (%s): max|err| = %.3e, max|ref| = %.3e, rel = %.3e  -> %s\n", variant,
               variant == 0 ? "LBO = N-block stride, SBO = 1024" : "LBO = 1024, SBO = N-block stride", maxerr, maxref, maxerr / maxref,
               maxerr / maxref < 5e-3 ? "layout accepted" : "WRONG");
    }
    return 0;
}
