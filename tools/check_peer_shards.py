"""Run under torchrun with 2 ranks (one GPU each): a dataset sharded over the ranks and peer-mapped through CUDA IPC
(sgm_dataset_ipc_handle / sgm_attach_peer_shards) must give what one process holding all rows gives.

  * sampling: a deterministic step with GLOBAL row indices (half of them remote) -- the parameters must agree with the
    single-process sampler bit for bit (same arithmetic, only the gather source differs);
  * control variates: mode finder over global rows + full gradient = each rank's shard, all-reduced -- theta_hat identical,
    g_hat within 1e-5 of the single-process one.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_peer_shards.py
NOT YET RUN ON A GPU BOX (written after gpurun was closed for round 1).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def attach(sg, rank, world):
    from sgmcmc_b200 import _lib
    lib = _lib.load()
    mine = (C.c_char * _lib.IPC_HANDLE_BYTES)()
    _lib.check(lib.sgm_dataset_ipc_handle(sg.handle, mine, _lib.IPC_HANDLE_BYTES))
    t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).clone().cuda()
    allh = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allh, t)
    blob = b"".join(x.cpu().numpy().tobytes() for x in allh)
    _lib.check(lib.sgm_attach_peer_shards(sg.handle, blob, world, rank))


def comm(sg, rank, world):
    from sgmcmc_b200 import _lib
    lib = _lib.load()
    idbuf = torch.zeros(_lib.NCCL_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_char * _lib.NCCL_ID_BYTES)()
        _lib.check(lib.sgm_comm_unique_id(raw, _lib.NCCL_ID_BYTES))
        idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    idbuf = idbuf.cuda()
    dist.broadcast(idbuf, 0)
    _lib.check(lib.sgm_comm_init(sg.handle, idbuf.cpu().numpy().tobytes(), _lib.NCCL_ID_BYTES, rank, world))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sgmcmc_b200 as S
    from sgmcmc_b200 import validation as V
    from tests.cases import make_case
    case = make_case("logistic130")          # ld = 132 floats > 128: the long-row kernel, the one that follows shard pointers
    N, n = case["N"], case["n"]
    assert N % world == 0
    per = N // world
    shard = {k: v[rank * per:(rank + 1) * per] for k, v in case["data"].items()}
    chains = 3
    kw = dict(logPrior=case["logprior"], minibatchSize=n, seed=7, device=local, nChains=chains)
    ok = True

    # ---- plain SGLD, deterministic step with global indices
    sh = S.sgldSetup(case["loglik"], shard, case["params"], case["eps"], **kw)
    attach(sh, rank, world)
    V.initSessDeterministic(sh, None, None)
    ref = S.sgldSetup(case["loglik"], case["data"], case["params"], case["eps"], **kw)
    V.initSessDeterministic(ref, None, None)
    rs = np.random.RandomState(3)
    for _ in range(3):
        idx = np.stack([rs.choice(N, n, replace=False)[None] for _ in range(chains)]).astype(np.int32)
        noise = rs.randn(chains, V.noiseLen(sh)).astype(np.float32)
        V.stepDeterministic(sh, idx, noise)
        V.stepDeterministic(ref, idx, noise)
    a, b = V.flatParams(sh), V.flatParams(ref)
    same = np.array_equal(a, b)
    print(f"rank {rank}: sharded vs whole dataset, 3 deterministic SGLD steps: {'bit-identical' if same else 'DIFFERENT, max |d| = %.3e' % np.abs(a - b).max()}", flush=True)
    ok &= same

    # ---- SGLD-CV with one shared mode: device RNG (same streams on both sides), full gradient = shards all-reduced
    kwcv = dict(kw, sharedMode=True, nItersOpt=40, verbose=False)
    shcv = S.sgldcvSetup(case["loglik"], shard, case["params"], case["eps"], case["opt_eps"], **kwcv)
    attach(shcv, rank, world)
    comm(shcv, rank, world)
    S.initSess(shcv, verbose=False)
    refcv = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"], **kwcv)
    S.initSess(refcv, verbose=False)
    th, g = V.getMode(shcv)
    th_r, g_r = V.getMode(refcv)
    same = np.array_equal(th, th_r)
    err = np.abs(g - g_r).max() / max(1.0, np.abs(g_r).max())
    print(f"rank {rank}: control variates: modes {'identical' if same else 'DIFFERENT'}, g_hat rel err {err:.2e}", flush=True)
    ok &= same and err < 1e-5

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    from sgmcmc_b200 import _lib
    for o in (sh, shcv):   # peers unmap, barrier, then the owners free their shards
        _lib.check(_lib.load().sgm_detach_peer_shards(o.handle))
    dist.barrier()
    del sh, shcv
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
