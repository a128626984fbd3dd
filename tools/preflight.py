#!/usr/bin/env python
"""Single-GPU pre-flight: every kernel path once, at tiny shapes, each in its own process, under compute-sanitizer.

    python tools/preflight.py [--no-sanitizer] [--only NAME] [--timeout SECONDS]

RULE (README "GPU order"): no kernel path's first execution may happen at more than one GPU.  The order on a fresh box is
    1. python -m pytest tests -q -m "not gpu"     (CPU: oracle, launch plans at full sizes, the emulated kernel sources)
    2. python tools/preflight.py                  (ONE GPU, this file)
    3. python -m pytest tests -q -m gpu           (ONE GPU)
    4. python bench.py                            (ONE GPU)
    5. 2 GPUs, then 4, then 8.
Round 1 skipped steps 2-4 for a new TMA path, went straight to 8 GPUs with a row shard that started at an odd row, and lost its
GPU access to the Xid 13s that followed.  Shard starts are therefore exercised here on one GPU (sgm_set_option "fullgrad_row0" /
"fullgrad_row1": what rank r of a row-sharded job computes, no communicator needed).

Every check runs in a child process with a timeout (a hang is killed, never waited for) and stops at the first failure: nothing
after a fault is launched.  Exit code 0 = all paths clean.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = """
import sys, numpy as np
sys.path.insert(0, %r)
import sgmcmc_b200 as S
from sgmcmc_b200 import validation as V
from oracle import samplers as OS
from tests.cases import device_setup, make_case
from tests import test_parity_gpu as T, test_fullgrad_tc_gpu as F
""" % ROOT

# name -> python snippet (appended to PRELUDE).  Ordered from the simplest kernels to the tensor-core pipelines.
CHECKS = [
    ("gauss / mixture chain kernels", "T._deterministic_steps('gauss2d', 'sgld', True, False); T._deterministic_steps('mixture', 'sghmc', True, True)"),
    ("logistic generic kernel (cp.async-free SIMT)", """
case = make_case('logistic54'); dev = device_setup(case, 'sgnht', True, n_chains=3, n_iters_opt=5); V.setOption(dev, 'logistic_kernel', 1)
S.initSess(dev, verbose=False); S.sgmcmcStep(dev, None); assert np.isfinite(V.flatParams(dev)).all()"""),
    ("logistic short-row kernel: cp.async gather", """
case = make_case('logistic54'); dev = device_setup(case, 'sgld', True, n_chains=3, n_iters_opt=5); V.setOption(dev, 'gather', 1)
S.initSess(dev, verbose=False); S.sgmcmcStep(dev, None); assert np.isfinite(V.flatParams(dev)).all()"""),
    ("logistic short-row kernel: TMA bulk copy per row", """
case = make_case('logistic54'); dev = device_setup(case, 'sgld', True, n_chains=3, n_iters_opt=5); V.setOption(dev, 'gather', 2)
S.initSess(dev, verbose=False); S.sgmcmcStep(dev, None); assert np.isfinite(V.flatParams(dev)).all()"""),
    ("logistic short-row kernel: TMA tile::gather4 (default), R = 1 and R = 2", "T._deterministic_steps('logistic54', 'sgld', True, True, options=T.ONE_CTA_FINDER); T._deterministic_steps('logistic54_n500', 'sghmc', True, False, options=T.ONE_CTA_FINDER)"),
    ("logistic long rows d = 1024: chain kernel <32, 9> + streaming full gradient", "T._deterministic_steps('logistic1024_n1024_N3000', 'sgnht', True, True, options=T.ONE_CTA_FINDER)"),
    ("mode finder on a thread-block cluster per mode (cluster launch, barrier.cluster, DSMEM loads): short and long rows",
     "T._deterministic_steps('logistic54', 'sgld', True, True, options={'mode_finder_split': 8}); T._deterministic_steps('logistic54_n101', 'sgld', True, False, options={'mode_finder_split': 4}); "
     "T._deterministic_steps('logistic1024_n1024_N3000', 'sgld', True, True, options={'mode_finder_split': 8})"),
    ("sampling on a thread-block cluster per chain (few chains): all samplers, Gauss-Seidel passes, control variates, long rows",
     "T._deterministic_steps('logistic54', 'sgld', True, True, options={'chain_split': 8}); T._deterministic_steps('logistic54_n101', 'sghmc', False, False, options={'chain_split': 4}); "
     "T._deterministic_steps('logistic130', 'sgnht', True, False, options={'chain_split': 2}); T._deterministic_steps('logistic1024_n1024_N3000', 'sgnht', True, True, options={'chain_split': 8}, n_opt=5)"),
    ("mode finder with one CTA per mode (short-row and generic kernels in their SGD mode)",
     "T._deterministic_steps('logistic54', 'sgld', True, True, options={'mode_finder_split': 1, 'chain_split': 1}); T._deterministic_steps('logistic1024_n1024_N3000', 'sgld', True, True, options={'mode_finder_split': 1, 'chain_split': 1})"),
    ("full gradient over row ranges that start anywhere: streaming kernel", "F.test_full_gradient_over_a_row_range_that_starts_anywhere(None, 'stream')"),
    ("full gradient over row ranges that start anywhere: SIMT grid", "F.test_full_gradient_over_a_row_range_that_starts_anywhere(None, 'generic')"),
    ("tcgen05 many-mode full gradient (cvg_kernel)", "F.test_many_mode_full_gradient_matches_oracle(None, 'logistic54', 20)"),
    ("tcgen05 many-mode full gradient: row tail, several K splits", "F.test_many_mode_full_gradient_row_tail_and_blocks(None)"),
    ("tcgen05 many-mode full gradient over row ranges that start anywhere (the round-1 fault)", "F.test_full_gradient_over_a_row_range_that_starts_anywhere(None, 'tc')"),
    ("tcgen05 many-mode full gradient at d = 1024 (4 TMEM segments per item)", "F.many_mode_check(make_case('logistic1024_n200_N515'), 128, n_opt=3)"),
    ("softmax net, reduced width", "T._deterministic_steps('softmaxnet_small', 'sghmc', True, False)"),
    ("softmax net 784-100-10: TMA-fed forward GEMM, tcgen05 backward GEMM, streamed update", "T._deterministic_steps('softmaxnet', 'sghmc', False, False)"),
    ("softmax net: register-staged forward GEMM and update", """
case = make_case('softmaxnet_small'); dev = device_setup(case, 'sgld', True, n_chains=2, n_iters_opt=3); V.setOption(dev, 'nn_stage1', 1); V.setOption(dev, 'nn_update', 1)
S.initSess(dev, verbose=False); S.sgmcmcStep(dev, None); assert np.isfinite(V.flatParams(dev)).all()"""),
    ("Gaussian families: dataset staged in shared memory once per CTA (TMA bulk copies)", "T.test_small_models_shared_memory_dataset_equals_global_gather(None, 'gauss2d', 'sgld'); T.test_small_models_shared_memory_dataset_equals_global_gather(None, 'mixture', 'sghmc')"),
    ("dataset upload from C-order / column-major / strided host matrices", "T.test_dataset_layouts_upload_identically(None, 'logistic54')"),
    ("sgm_sample: chain streamed to the host on a second stream while later blocks run", """
case = make_case('logistic54')
ch = S.sgldcv(case['loglik'], case['data'], case['params'], case['eps'], case['opt_eps'], logPrior=case['logprior'], minibatchSize=case['n'], nIters=230, nItersOpt=120, verbose=False, seed=3, nChains=340)
assert all(np.isfinite(np.asarray(v)).all() for v in ch.values())"""),
    ("mode finder: launches of 100 iterations with checkOptDivergence between them", """
case = make_case('logistic54'); dev = device_setup(case, 'sgld', True, n_chains=2, n_iters_opt=250); S.initSess(dev, verbose=False); assert np.isfinite(V.flatParams(dev)).all()"""),
    ("estLogPost / held-out log-likelihood / chain copy-out", "T.test_est_logpost_matches_oracle(None, 'logistic54'); T.test_heldout_loglik_matches_oracle(None, 'mixture'); T.test_chain_storage_layout_and_shapes(None)"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-sanitizer", action="store_true", help="run with CUDA_LAUNCH_BLOCKING=1 instead of compute-sanitizer")
    ap.add_argument("--only", default=None, help="substring of the check names to run")
    ap.add_argument("--timeout", type=int, default=600, help="seconds per check")
    args = ap.parse_args()
    import torch  # noqa: F401  (pages the image in; also the quickest way to see whether a device is visible)
    if not torch.cuda.is_available():
        print("preflight: no CUDA device visible -- nothing was run (the CPU-side checks are `pytest -m 'not gpu'`)")
        return 2
    if torch.cuda.device_count() > 1 and "CUDA_VISIBLE_DEVICES" not in os.environ:
        print("preflight: more than one GPU visible; pinning to device 0 (pre-flight is a ONE-GPU step)")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0], CUDA_LAUNCH_BLOCKING="1")
    sanitizer = None if args.no_sanitizer else (shutil.which("compute-sanitizer") or ("/usr/local/cuda/bin/compute-sanitizer" if os.path.exists("/usr/local/cuda/bin/compute-sanitizer") else None))
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()
    failed = 0
    for name, body in CHECKS:
        if args.only and args.only not in name:
            continue
        cmd = [sys.executable, "-c", PRELUDE + body + "\nprint('PREFLIGHT_OK')"]
        if sanitizer:
            cmd = [sanitizer, "--tool", "memcheck", "--error-exitcode", "86", "--launch-timeout", "0", "--print-limit", "5"] + cmd
        t0 = time.time()
        try:
            out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=args.timeout)
            ok = out.returncode == 0 and "PREFLIGHT_OK" in out.stdout
            tail = (out.stdout[-1500:] + out.stderr[-2500:]) if not ok else ""
        except subprocess.TimeoutExpired:
            ok, tail = False, f"TIMEOUT after {args.timeout} s (the child was killed)"
        print(f"[{'ok' if ok else 'FAIL'}] {name}  ({time.time() - t0:.1f} s{', memcheck' if sanitizer else ''})", flush=True)
        if not ok:
            print(tail)
            failed += 1
            break   # nothing is launched after a fault
    print("preflight:", "all paths clean" if not failed else "STOPPED at the first failure -- do not go to more GPUs")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
