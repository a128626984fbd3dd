"""Multi-rank checks shared by the GPU scripts (tools/check_*.py under torchrun, NCCL) and the CPU-emulated multi-process test
(tests/test_emu_multirank.py: gloo + the emulated library + tests/emu/fake_nccl.c).  Each function runs on EVERY rank of an
initialised torch.distributed process group and returns a list of failure messages (empty = ok).

SURVEY.md 8e rows: the row-sharded control-variate pass (one shared mode: all-reduce; one mode per chain: all-gather of the modes,
contraction against all of them, reduce-scatter), the dataset sharded over the ranks and peer-mapped (the scale-out config), the
row-sharded mode finder.  The reference is single-process: the yardstick is always ONE process holding all rows."""
from __future__ import annotations

import numpy as np


def sharded_modes(rank: int, world: int, device: int = 0, case_name: str = "logistic54", rows: int = 3878, chains: int = 24, tc: bool = True):
    """One mode per chain, rows split over the ranks (rank 1 starts at an odd row): must equal one process over all rows."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import distributed as D
    from sgmcmc_b200 import validation as V
    from tests.cases import make_case
    case = make_case(case_name)
    case["data"] = {k: v[:rows] for k, v in case["data"].items()}
    case["N"] = rows
    kw = dict(logPrior=case["logprior"], minibatchSize=case["n"], seed=7, sharedMode=False, device=device, nItersOpt=30, verbose=False)
    bad = []
    sg = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"], nChains=chains, chainOffset=rank * chains, **kw)
    if tc:
        V.setOption(sg, "fullgrad", 1)   # tensor-core contractions whatever the mode count
    D.comm_init(sg, rank, world)
    S.initSess(sg, verbose=False)
    th, g = V.getMode(sg)
    ref = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"], nChains=chains, chainOffset=rank * chains, **kw)
    if tc:
        V.setOption(ref, "fullgrad", 1)
    S.initSess(ref, verbose=False)
    th_r, g_r = V.getMode(ref)
    if not np.array_equal(th, th_r):
        bad.append(f"rank {rank}: modes differ from the single-process run")
    err = np.abs(g - g_r).max() / max(np.abs(g_r).max(), 1.0)
    if not err < 1e-5:
        bad.append(f"rank {rank}: sharded vs single-process g_hat rel err {err:.2e}")
    return bad


def shared_mode_allreduce(rank: int, world: int, device: int = 0, case_name: str = "gauss2d"):
    """One shared mode, rows split over the ranks, partial sums all-reduced (every family)."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import distributed as D
    from sgmcmc_b200 import validation as V
    from tests.cases import make_case
    case = make_case(case_name)
    kw = dict(logPrior=case["logprior"], minibatchSize=case["n"], seed=11, sharedMode=True, device=device, nItersOpt=20, verbose=False, nChains=2, chainOffset=2 * rank)
    sg = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"], **kw)
    D.comm_init(sg, rank, world)
    S.initSess(sg, verbose=False)
    th, g = V.getMode(sg)
    ref = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"], **dict(kw, chainOffset=0))
    S.initSess(ref, verbose=False)
    th_r, g_r = V.getMode(ref)
    bad = []
    if not np.array_equal(th, th_r):
        bad.append(f"rank {rank}: {case_name}: the shared mode differs between the sharded and the single-process run")
    err = np.abs(g - g_r).max() / max(np.abs(g_r).max(), 1.0)
    if not err < 1e-5:
        bad.append(f"rank {rank}: {case_name}: all-reduced g_hat rel err {err:.2e}")
    return bad


def peer_shards(rank: int, world: int, device: int = 0, case_name: str = "logistic130"):
    """Dataset sharded over the ranks and peer-mapped (option (i)): sampling with GLOBAL rows must be
    bit-identical to one process holding all rows; the control-variate pass sums the rank's shard and reduces; one chain per rank
    WITHOUT the shared-mode flag means one mode per rank (the residual cache must then stay off); detach -> init -> run works on the
    local rows again."""
    import torch.distributed as dist
    import sgmcmc_b200 as S
    from sgmcmc_b200 import distributed as D
    from sgmcmc_b200 import validation as V
    from tests.cases import make_case
    case = make_case(case_name)          # ld = 132 floats > 128: the long-row kernel, the one that follows shard pointers
    N, n = case["N"], case["n"]
    assert N % world == 0
    per = N // world
    shard = {k: v[rank * per:(rank + 1) * per] for k, v in case["data"].items()}
    chains = 3
    kw = dict(logPrior=case["logprior"], minibatchSize=n, seed=7, device=device, nChains=chains)
    bad = []

    # ---- plain SGLD, deterministic steps with global row numbers (most of them remote)
    sh = S.sgldSetup(case["loglik"], shard, case["params"], case["eps"], **kw)
    D.attach_peer_shards(sh, rank, world)
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
    if not np.array_equal(a, b):
        bad.append(f"rank {rank}: sharded vs whole dataset, deterministic SGLD steps differ (max |d| = {np.abs(a - b).max():.3e})")
    # estLogPost on global rows, N / n is the global ratio (R/setup.R:116-118)
    lp_idx = np.stack([rs.choice(N, n, replace=False) for _ in range(chains)]).astype(np.int32)
    la, lb = V.estLogPostDeterministic(sh, lp_idx), V.estLogPostDeterministic(ref, lp_idx)
    if not np.allclose(la, lb, rtol=1e-6, atol=1e-6):
        bad.append(f"rank {rank}: estLogPost over the sharded dataset differs: {la} vs {lb}")

    # ---- SGLD-CV, ONE shared mode: device RNG (same streams on both sides), full gradient = shards all-reduced, residual cached in the rows
    kwcv = dict(kw, sharedMode=True, nItersOpt=40, verbose=False)
    shcv = S.sgldcvSetup(case["loglik"], shard, case["params"], case["eps"], case["opt_eps"], **kwcv)
    D.attach_peer_shards(shcv, rank, world)
    D.comm_init(shcv, rank, world)
    sess = S.initSess(shcv, verbose=False)
    refcv = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"], **kwcv)
    rsess = S.initSess(refcv, verbose=False)
    th, g = V.getMode(shcv)
    th_r, g_r = V.getMode(refcv)
    err = np.abs(g - g_r).max() / max(1.0, np.abs(g_r).max())
    if not np.array_equal(th, th_r):
        bad.append(f"rank {rank}: control variates over the sharded dataset: modes differ")
    if not err < 1e-5:
        bad.append(f"rank {rank}: control variates over the sharded dataset: g_hat rel err {err:.2e}")
    for _ in range(3):
        S.sgmcmcStep(shcv, sess)
        S.sgmcmcStep(refcv, rsess)
    a, b = V.flatParams(shcv), V.flatParams(refcv)
    scale = np.abs(b).max()
    if not np.abs(a - b).max() <= 2e-5 * scale:
        bad.append(f"rank {rank}: SGLD-CV steps over the sharded dataset drift from the single-process chain: {np.abs(a - b).max() / scale:.2e}")

    # ---- one chain per rank, NO shared-mode flag: every rank has its own mode (chain_offset differs); the cached residual of a peer's
    # rows would belong to the peer's mode, so the cache must be off and the result must equal the single-process run of THIS chain
    kw1 = dict(kw, nChains=1, chainOffset=rank, sharedMode=False, nItersOpt=25, verbose=False)
    one = S.sgldcvSetup(case["loglik"], shard, case["params"], case["eps"], case["opt_eps"], **kw1)
    D.attach_peer_shards(one, rank, world)
    D.comm_init(one, rank, world)
    osess = S.initSess(one, verbose=False)
    ref1 = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"], **kw1)
    r1sess = S.initSess(ref1, verbose=False)
    t1, g1 = V.getMode(one)
    t1r, g1r = V.getMode(ref1)
    if not np.array_equal(t1, t1r):
        bad.append(f"rank {rank}: one chain per rank: the rank's own mode differs from the single-process run")
    e1 = np.abs(g1 - g1r).max() / max(1.0, np.abs(g1r).max())
    if not e1 < 1e-5:
        bad.append(f"rank {rank}: one chain per rank: g_hat rel err {e1:.2e} (partial sums taken at different modes?)")
    for _ in range(2):
        S.sgmcmcStep(one, osess)
        S.sgmcmcStep(ref1, r1sess)
    a, b = V.flatParams(one), V.flatParams(ref1)
    if not np.abs(a - b).max() <= 2e-5 * np.abs(b).max():
        bad.append(f"rank {rank}: one chain per rank: sampling over the sharded dataset is biased by a peer's cached residual")

    # ---- detach: back on the local rows (N, the permutation width and N / n are the local ones again)
    for o in (sh, shcv, one):
        D.detach_peer_shards(o)
    loc = S.sgldSetup(case["loglik"], shard, case["params"], case["eps"], **kw)
    V.initSessDeterministic(sh, None, None)
    V.initSessDeterministic(loc, None, None)
    idx = np.stack([rs.choice(per, n, replace=False)[None] for _ in range(chains)]).astype(np.int32)
    noise = rs.randn(chains, V.noiseLen(sh)).astype(np.float32)
    V.stepDeterministic(sh, idx, noise)
    V.stepDeterministic(loc, idx, noise)
    if not np.array_equal(V.flatParams(sh), V.flatParams(loc)):
        bad.append(f"rank {rank}: after detach the handle does not behave like one created on the local shard")
    S.initSess(sh, verbose=False)   # device RNG over the LOCAL row space must not read past the shard
    from sgmcmc_b200 import _lib
    _lib.check(_lib.load().sgm_run(sh.handle, 5, 0))
    dist.barrier()
    return bad


def sharded_mode_finder(rank: int, world: int, device: int = 0, case_name: str = "logistic54"):
    """The *cv mode finder sharded by data rows (north star / SURVEY.md 8e row 3): with the SAME minibatch rows per SGD step
    (validation mode) the sharded finder -- per-rank partial sums of the rows each rank owns, all-reduced every step -- must land on
    the mode the single-process finder finds; with device RNG (stratified minibatches) it must move towards the mode too."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import distributed as D
    from sgmcmc_b200 import validation as V
    from tests.cases import make_case
    case = make_case(case_name)
    N, n, n_opt = case["N"], case["n"], 60
    kw = dict(logPrior=case["logprior"], minibatchSize=n, seed=5, sharedMode=True, device=device, nItersOpt=n_opt, verbose=False, nChains=2)
    bad = []
    rs = np.random.RandomState(17)
    opt_idx = np.stack([rs.choice(N, n, replace=False) for _ in range(n_opt)]).astype(np.int32)[None]   # [1 mode][n_opt][n]
    opt_idx[0, 7] = np.sort(rs.choice(np.arange(N // world, N), n, replace=False))                         # one minibatch with no row on rank 0
    layouts = ("replicated", "sharded") if case_name.startswith("logistic") else ("replicated",)   # peer-mapped shards: logistic family
    for layout in layouts:
        if layout == "replicated":
            data = case["data"]
        else:
            per = N // world
            data = {k: v[rank * per:(rank + 1) * per] for k, v in case["data"].items()}
        sg = S.sgldcvSetup(case["loglik"], data, case["params"], case["eps"], case["opt_eps"] * 20, **kw)
        if layout == "sharded":
            D.attach_peer_shards(sg, rank, world)
        D.comm_init(sg, rank, world)
        V.setOption(sg, "sharded_mode_finder", 1)
        V.initSessDeterministic(sg, None, opt_idx)
        th, g = V.getMode(sg)
        ref = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"] * 20, **kw)
        V.initSessDeterministic(ref, None, opt_idx)
        th_r, g_r = V.getMode(ref)
        moved = np.abs(th_r[0] - np.concatenate([np.ravel(v, order="F") for v in case["params"].values()])).max()
        err = np.abs(th - th_r).max()
        if not err <= 1e-5 * max(np.abs(th_r).max(), moved):
            bad.append(f"rank {rank}: {layout}: row-sharded mode finder vs single process: |d theta_hat| = {err:.3e} (the finder moved {moved:.3e})")
        eg = np.abs(g - g_r).max() / max(1.0, np.abs(g_r).max())
        if not eg < 1e-4:
            bad.append(f"rank {rank}: {layout}: full gradient at the sharded finder's mode differs: {eg:.2e}")
        # device RNG: stratified minibatches (n / world rows from every rank's own rows)
        sg2 = S.sgldcvSetup(case["loglik"], data, case["params"], case["eps"], case["opt_eps"] * 20, **dict(kw, nItersOpt=200))
        if layout == "sharded":
            D.attach_peer_shards(sg2, rank, world)
        D.comm_init(sg2, rank, world)
        V.setOption(sg2, "sharded_mode_finder", 1)
        S.initSess(sg2, verbose=False)
        th2, g2 = V.getMode(sg2)
        ref2 = S.sgldcvSetup(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"] * 20, **dict(kw, nItersOpt=200))
        S.initSess(ref2, verbose=False)
        th2r, g2r = V.getMode(ref2)
        # different minibatches, same target: the full-data gradient norm at the two modes must be of the same size (both descended)
        n2, n2r = np.linalg.norm(g2[0]), np.linalg.norm(g2r[0])
        if not (np.isfinite(th2).all() and n2 <= 1.5 * n2r + 1e-3):
            bad.append(f"rank {rank}: {layout}: device-RNG sharded finder did not descend like the single-process one (|g| {n2:.3e} vs {n2r:.3e})")
        if layout == "sharded":
            for o in (sg, sg2):
                D.detach_peer_shards(o)
    return bad
