"""Multi-GPU plumbing above the C ABI: one process per GPU, `torch.distributed` only carries the few bytes the
library needs exchanged (the NCCL unique id, CUDA IPC handles); the data path collectives are NCCL calls inside the library.

    comm_init(sg, rank, world)            row-sharded control-variate pass (sgm_comm_init)
    attach_peer_shards(sg, rank, world)   dataset sharded over the ranks, peers mapped through CUDA IPC (sgm_attach_peer_shards)
    detach_peer_shards(sg)

The process group may be NCCL (GPU boxes) or gloo (the CPU-emulated test build): byte tensors live where the backend wants them.
"""
from __future__ import annotations

import ctypes as C

from . import _lib


def _bytes_tensor(raw: bytes):
    import torch
    import torch.distributed as dist
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    return t.cuda() if dist.get_backend() == "nccl" else t


_job_id = None   # the unique id of this job's communicator: exchanged once, every sampler of the process then shares one ncclComm


def comm_init(sg, rank: int, world: int):
    """Gives the handle the job's NCCL communicator (created by the first call: id broadcast from rank 0; later calls reuse it)."""
    global _job_id
    import torch.distributed as dist
    lib = _lib.load()
    if _job_id is None or _job_id[1] != (rank, world):
        raw = (C.c_char * _lib.NCCL_ID_BYTES)()
        if rank == 0:
            _lib.check(lib.sgm_comm_unique_id(raw, _lib.NCCL_ID_BYTES))
        idbuf = _bytes_tensor(raw.raw)
        dist.broadcast(idbuf, 0)
        _job_id = (idbuf.cpu().numpy().tobytes(), (rank, world))
    _lib.check(lib.sgm_comm_init(sg.handle, _job_id[0], _lib.NCCL_ID_BYTES, rank, world))


def attach_peer_shards(sg, rank: int, world: int):
    """Every rank keeps its own N rows; the peers' shards are mapped so that the chains sample from world x N rows."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    mine = (C.c_char * _lib.IPC_HANDLE_BYTES)()
    _lib.check(lib.sgm_dataset_ipc_handle(sg.handle, mine, _lib.IPC_HANDLE_BYTES))
    t = _bytes_tensor(mine.raw)
    allh = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allh, t)
    _lib.check(lib.sgm_attach_peer_shards(sg.handle, b"".join(x.cpu().numpy().tobytes() for x in allh), world, rank))


def detach_peer_shards(sg):
    """Unmaps the peers' shards; every rank calls it, then a barrier, then the owners may free their datasets."""
    import torch.distributed as dist
    _lib.check(_lib.load().sgm_detach_peer_shards(sg.handle))
    dist.barrier()
