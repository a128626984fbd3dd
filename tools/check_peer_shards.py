"""Run under torchrun with 2 ranks (one GPU each): a dataset sharded over the ranks and peer-mapped through CUDA IPC
(sgm_dataset_ipc_handle / sgm_attach_peer_shards) must give what one process holding all rows gives -- deterministic sampling on
global rows bit for bit, the control-variate pass (shared mode, and one mode per rank) within 1e-5, and a clean detach.  The checks
live in tests/multirank_checks.py and also run on CPU (tests/test_emu_multirank.py, emulated build).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_peer_shards.py
First GPU execution pending: gpurun was closed when this path was written (round 1) and throughout round 2.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tests import multirank_checks as M
    bad = M.peer_shards(rank, world, device=local)
    for m in bad:
        print(m, flush=True)
    ok = torch.tensor([0 if bad else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    print(f"rank {rank}: sharded dataset {'ok' if not bad else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(ok) == 1 else 1)


if __name__ == "__main__":
    main()
