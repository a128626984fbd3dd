"""Multi-rank paths as real multi-process jobs on the CPU: one process per "rank", each with the CPU-emulated build of the kernel
sources (tests/emu), collectives through tests/emu/fake_nccl.c (POSIX shared memory), peer-mapped shards through the emulated CUDA
IPC (shared-memory backed allocations), rendezvous over gloo.  The same check functions run under torchrun + NCCL on GPUs
(tools/check_sharded_modes.py, tools/check_peer_shards.py, tests/test_multi_gpu.py)."""
import os
import socket

import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, which, out):
    import torch.distributed as dist
    from tests.emu import backend, build
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SGM_EMU_NCCL=build.build_fake_nccl(), SGM_EMU_SHM_IPC="1", SGM_EMU_SMS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import multirank_checks as M
    bad = []
    try:
        with backend.emulated() as lib:
            for name, kwargs in which:
                bad += getattr(M, name)(rank, world, **kwargs)
            msg = lib.emu_last_message()
            if msg:
                bad.append(msg.decode())
    except Exception as e:   # noqa: BLE001
        import traceback
        bad.append(f"rank {rank}: {type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}")
    out[rank] = bad
    dist.barrier()
    dist.destroy_process_group()


def _run(world, which):
    from tests.emu import backend, build
    backend.library()          # build once in the parent, not concurrently in the children
    build.build_fake_nccl()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), which, out), nprocs=world, join=True)
    problems = [m for r in range(world) for m in out.get(r, [f"rank {r} did not report"])]
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_pass_with_one_mode_per_chain(world):
    _run(world, [("sharded_modes", dict(rows=3878 if world == 2 else 3001, chains=8))])


def test_row_sharded_pass_with_a_shared_mode_every_family():
    _run(2, [("shared_mode_allreduce", dict(case_name=c)) for c in ("gauss2d", "mixture", "logistic54", "logistic130", "softmaxnet_small")])


def test_dataset_sharded_over_ranks_and_peer_mapped():
    _run(2, [("peer_shards", {})])


@pytest.mark.parametrize("world", [2, 3])
def test_mode_finder_sharded_by_rows(world):
    """SURVEY.md 8e row 3 (the north star's "mode-finding pass, sharded by data rows with an NCCL allreduce of the d-vector")."""
    _run(world, [("sharded_mode_finder", dict(case_name="logistic54" if world == 2 else "gauss2d"))])
