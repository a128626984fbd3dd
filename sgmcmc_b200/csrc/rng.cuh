// Counter-based RNG for the sampler kernels (device + host-callable).
//
// Replaces, with stateless streams, the two RNGs of the reference:
//   * R's sample(N, n) without replacement        (R/update.R:171)
//   * tf$random_normal in getNoise / momentum init (R/dynamics.R:47,84,135-142)
// Mirrored line for line by oracle/rng.py (indices bit-exact, normals to float32
// rounding of log/sin/cos).
#pragma once
#include <cstdint>

namespace sgm {

enum RngTag : uint32_t {
    TAG_NOISE = 1, TAG_INDEX = 2, TAG_MOMENTUM = 3, TAG_OPT_INDEX = 4, TAG_CHECK_INDEX = 5, TAG_DATA = 6
};

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011).
__host__ __device__ __forceinline__ U4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}

#if defined(__CUDACC__) || defined(SGM_EMU)
// Two uint32 words -> two N(0,1) float32 (Box-Muller on 24-bit uniforms).
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = ((float)(a >> 8) + 1.0f) * 5.9604644775390625e-8f;  // (0,1]
    const float u2 = (float)(b >> 8) * 5.9604644775390625e-8f;           // [0,1)
    // fast intrinsics: |error| <~ 1e-6 in a unit-variance variate (the noise is then scaled by sqrt(2 eps))
    const float r = sqrtf(-2.0f * __logf(u1));
    const float ang = 6.283185307179586f * u2 - (u2 >= 0.5f ? 6.283185307179586f : 0.f);   // in [-pi, pi)
    z0 = r * __cosf(ang); z1 = r * __sinf(ang);
}

// N(0,1) for flat element e of stream (chain, word2, tag): Philox block e/4, lane e%4.
__device__ __forceinline__ float normal_at(uint32_t e, uint32_t chain, uint32_t word2, uint32_t tag,
                                           uint32_t k0, uint32_t k1) {
    const U4 r = philox4x32(e >> 2, chain, word2, tag, k0, k1);
    float z0, z1;
    if (e & 2u) box_muller(r.z, r.w, z0, z1); else box_muller(r.x, r.y, z0, z1);
    return (e & 1u) ? z1 : z0;
}
__device__ __forceinline__ void normal4_at(uint32_t blk, uint32_t chain, uint32_t word2, uint32_t tag,
                                           uint32_t k0, uint32_t k1, float z[4]) {
    const U4 r = philox4x32(blk, chain, word2, tag, k0, k1);
    box_muller(r.x, r.y, z[0], z[1]);
    box_muller(r.z, r.w, z[2], z[3]);
}
#endif

// Keyed bijection on [0, N): 4 rounds of (odd multiply + key) / xorshift on
// ceil(log2 N) bits, cycle-walked into range.  perm(0..n-1) are n distinct rows.
struct Perm {
    uint32_t key[4];
    uint32_t mask, shift, N;
};

__host__ __device__ __forceinline__ uint32_t perm_bits(uint32_t N) {
    uint32_t b = 1;
    while ((1ull << b) < (uint64_t)N) ++b;
    return b;
}

__host__ __device__ __forceinline__ Perm make_perm(uint32_t N, uint32_t bits, uint32_t chain, uint32_t word2,
                                                   uint32_t tag, uint32_t k0, uint32_t k1) {
    const U4 k = philox4x32(0u, chain, word2, tag, k0, k1);
    Perm p;
    p.key[0] = k.x; p.key[1] = k.y; p.key[2] = k.z; p.key[3] = k.w;
    p.mask = (bits >= 32) ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    p.shift = bits / 2 > 0 ? bits / 2 : 1;
    p.N = N;
    return p;
}

__host__ __device__ __forceinline__ uint32_t perm_at(const Perm& p, uint32_t j) {
    uint32_t x = j;
    do {
        x = (x * (0x9E3779B1u | 1u) + p.key[0]) & p.mask; x ^= x >> p.shift;
        x = (x * (0x85EBCA6Bu | 1u) + p.key[1]) & p.mask; x ^= x >> p.shift;
        x = (x * (0xC2B2AE35u | 1u) + p.key[2]) & p.mask; x ^= x >> p.shift;
        x = (x * (0x27D4EB2Fu | 1u) + p.key[3]) & p.mask; x ^= x >> p.shift;
    } while (x >= p.N);
    return x;
}

}  // namespace sgm
