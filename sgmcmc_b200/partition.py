"""How the path shards over GPUs -- pure host logic.

* independent chains: contiguous chain ranges per rank, no communication;
* control-variate full-data gradient: contiguous row ranges per rank, partial sums
  all-reduced (NCCL on the device; the same formula as csrc/api.cu full_gradient()).
"""
from __future__ import annotations


def chain_range(rank: int, world: int, n_chains: int):
    """Chains [lo, hi) owned by ``rank`` (sizes differ by at most one)."""
    return n_chains * rank // world, n_chains * (rank + 1) // world


def row_range(rank: int, world: int, n_rows: int):
    """Dataset rows [lo, hi) whose gradient ``rank`` sums (matches api.cu: N*rank/world)."""
    return n_rows * rank // world, n_rows * (rank + 1) // world
