"""CPU suite, world_size 2 over gloo: the host-side sharding logic of the N>1 path.

Chains are partitioned with no communication; the control-variate full-data gradient is
summed over per-rank row ranges and all-reduced.  Here the per-rank partial sums come from
the oracle (no GPU), the reduction from torch.distributed/gloo -- the same plumbing bench.py
uses with NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import samplers as OS
from sgmcmc_b200.partition import aligned_k_walk, chain_range, row_range, shard_of_row
from tests.cases import make_case


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = make_case("logistic5")
    lo, hi = row_range(rank, world, case["N"])
    shard = {k: v[lo:hi] for k, v in case["data"].items()}
    part = case["om"].grad_loglik(case["params"], shard, np.float64)
    t = torch.from_numpy(OS.flat(part).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    # max-over-ranks timing reduction used by bench.py
    ms = torch.tensor([10.0 + rank])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        out["sum"] = t.numpy().copy()
        out["ms"] = float(ms)
    dist.destroy_process_group()


def test_row_sharded_full_gradient_allreduce_equals_single_rank():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    case = make_case("logistic5")
    full = OS.flat(case["om"].grad_loglik(case["params"], case["data"], np.float64))
    np.testing.assert_allclose(out["sum"], full, rtol=1e-12, atol=1e-10)
    assert out["ms"] == 11.0


def test_partitions_cover_everything_once():
    for world in (1, 2, 3, 4, 8):
        for total in (1, 7, 128, 1024, 581012):
            chains = [chain_range(r, world, total) for r in range(world)]
            rows = [row_range(r, world, total) for r in range(world)]
            for parts in (chains, rows):
                assert parts[0][0] == 0 and parts[-1][1] == total
                assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
                sizes = [b - a for a, b in parts]
                assert max(sizes) - min(sizes) <= 1


def _worker_modes(rank, world, port, out):
    """One mode per chain, chains split over ranks, rows split over ranks (csrc/api.cu fullgrad_gather_modes / full_gradient):
    all-gather of the modes, every rank sums ITS rows at ALL modes, reduce-scatter hands each rank its own chains' totals."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = make_case("logistic5")
    chains = 3
    rs = np.random.RandomState(100 + rank)
    modes = [OS.flat({k: np.asarray(v) + 0.1 * rs.randn(*np.shape(v)) for k, v in case["params"].items()}) for _ in range(chains)]
    mine = torch.from_numpy(np.stack(modes))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    all_modes = torch.cat(gathered).numpy()                      # [world * chains, P], rank-major like ncclAllGather
    lo, hi = row_range(rank, world, case["N"])
    shard = {k: v[lo:hi] for k, v in case["data"].items()}
    part = np.stack([OS.flat(case["om"].grad_loglik(OS.unflat(m, case["params"]), shard, np.float64)) for m in all_modes])
    # gloo has no reduce_scatter: all_reduce + this rank's slice is the same sum (the device path calls ncclReduceScatter)
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    own = t[rank * chains:(rank + 1) * chains]
    full = np.stack([OS.flat(case["om"].grad_loglik(OS.unflat(m, case["params"]), case["data"], np.float64)) for m in modes])
    out[rank] = float(np.abs(own.numpy() - full).max() / max(1.0, np.abs(full).max()))
    dist.destroy_process_group()


def test_row_sharded_full_gradient_with_one_mode_per_chain():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_modes, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world and all(e < 1e-12 for e in out.values()), dict(out)


def test_pass2_k_walk_is_aligned_and_covers_the_shard():
    for n_rows in (3878, 581012, 12500000):
        for world in (1, 2, 4, 8):
            for rank in range(world):
                lo, hi = row_range(rank, world, n_rows)
                blocks = aligned_k_walk(lo, hi - lo, block=524288)
                covered = 0
                for k0, shift, n in blocks:
                    assert k0 % 32 == 0 and 0 <= shift < 32      # TMA box starts: 16-byte aligned
                    assert k0 + shift == lo + covered            # first data row of the block
                    covered += n
                assert covered == hi - lo


def test_global_rows_map_onto_equal_shards():
    for world, per in ((2, 2000), (8, 12500000)):
        rows = np.array([0, 1, per - 1, per, world * per - 1])
        for r in rows:
            rank, local = shard_of_row(int(r), per)
            assert 0 <= rank < world and 0 <= local < per and rank * per + local == r
        assert world * per < 2 ** 32          # the kernels carry global rows as uint32
