"""Counter-based RNG restated in numpy (integer work is bit-exact with the device).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws minibatches with R's ``sample(N, n)`` (without replacement,
R/update.R:171) and noise with TensorFlow's stateful ``tf$random_normal``
(R/dynamics.R:135-142).  Neither stream is reproducible outside R/TF
(SURVEY.md Appendix A.7), so the B200 library defines its own stateless streams:

* Philox4x32-10 (Salmon et al. 2011; same round function as cuRAND/TF) keyed by
  the user seed, counter = (block, chain, step, tag);
* minibatch indices = the first ``n`` images of a keyed bijection on [0, N)
  (multiply-add / xorshift rounds on ceil(log2 N) bits with cycle walking):
  distinct by construction, i.e. sampling WITHOUT replacement like ``sample``;
* N(0,1) noise by Box-Muller on two 24-bit uniforms.

Everything here mirrors ``sgmcmc_b200/csrc/rng.cuh`` line for line; indices must
match bit-exactly, normals to float32 rounding of log/sin/cos.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

# stream tags (counter word 3)
TAG_NOISE = 1      # dynamics noise; word 2 = step * 64 + substream
TAG_INDEX = 2      # minibatch permutation keys
TAG_MOMENTUM = 3   # initial momentum draw
TAG_OPT_INDEX = 4  # mode-finder minibatches
TAG_CHECK_INDEX = 5  # divergence-check minibatches
TAG_DATA = 6       # synthetic dataset

PERM_MULT = (0x9E3779B1, 0x85EBCA6B, 0xC2B2AE35, 0x27D4EB2F)

_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def seed_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF


def box_muller(x0, x1):
    """Two uint32 words -> two float32 standard normals (same formula as rng.cuh)."""
    u1 = ((x0 >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)
    u2 = (x1 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    r = np.sqrt(np.float32(-2.0) * np.log(u1))
    ang = (np.float32(2.0) * u2).astype(np.float64) * np.pi  # sincospif(2*u2)
    return (r * np.cos(ang).astype(np.float32)).astype(np.float32), \
           (r * np.sin(ang).astype(np.float32)).astype(np.float32)


def normals(seed: int, chain: int, word2: int, tag: int, count: int) -> np.ndarray:
    """``count`` float32 normals; element e comes from Philox block e//4, lane e%4."""
    k0, k1 = seed_key(seed)
    nblk = (count + 3) // 4
    blk = np.arange(nblk, dtype=np.uint32)
    x0, x1, x2, x3 = philox4x32(blk, np.uint32(chain), np.uint32(word2 & 0xFFFFFFFF),
                                np.uint32(tag), k0, k1)
    z0, z1 = box_muller(x0, x1)
    z2, z3 = box_muller(x2, x3)
    out = np.stack([z0, z1, z2, z3], axis=1).reshape(-1)
    return out[:count].astype(np.float32)


def perm_bits(N: int) -> int:
    b = 1
    while (1 << b) < N:
        b += 1
    return b


def minibatch_indices(seed: int, chain: int, word2: int, tag: int, N: int, n: int) -> np.ndarray:
    """First n images of the keyed bijection on [0, N): n distinct 0-based rows."""
    assert 0 < n <= N < (1 << 31)
    k0, k1 = seed_key(seed)
    keys = philox4x32(np.uint32(0), np.uint32(chain), np.uint32(word2 & 0xFFFFFFFF),
                      np.uint32(tag), k0, k1)
    keys = [int(k) for k in keys]
    b = perm_bits(N)
    mask = (1 << b) - 1
    s = max(1, b // 2)
    x = np.arange(n, dtype=np.uint64)
    todo = np.ones(n, dtype=bool)
    while todo.any():
        y = x[todo]
        for r in range(4):
            y = (y * np.uint64(PERM_MULT[r] | 1) + np.uint64(keys[r])) & np.uint64(mask)
            y ^= y >> np.uint64(s)
        x[todo] = y
        todo = x >= N
    return x.astype(np.int32)
