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


def aligned_k_walk(row0: int, rows: int, block: int, align: int = 32):
    """Pass 2 of the many-mode control-variate gradient (csrc/api.cu launch_logistic_fullgrad_tc): the K walk over this rank's
    data rows starts at the ``align``-row aligned global row below ``row0`` (a TMA box start must be 16-byte aligned), so the
    first row of every block sits at column ``shift`` of the transposed residual buffer and the columns before it are zero.
    Returns [(k_start_global, shift, rows_in_block)] per row block."""
    shift = row0 % align
    out = []
    b0 = row0
    while b0 < row0 + rows:
        n = min(block, row0 + rows - b0)
        out.append((b0 - shift, shift, n))
        b0 += block
    return out


def shard_of_row(row: int, shard_rows: int):
    """Dataset sharded over ranks in equal shards (csrc/common.cuh row_ptr, sgm_attach_peer_shards): global row -> (owning rank,
    local row).  The chains sample global rows; rows of other ranks are loaded through the peer mapping."""
    return row // shard_rows, row % shard_rows
