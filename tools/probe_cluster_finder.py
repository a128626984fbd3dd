#!/usr/bin/env python
"""Probe of the cluster kernel (csrc/logistic_split.cuh: a thread-block cluster per chain / per mode) in a process of its own.

    python tools/probe_cluster_finder.py                    # exit 0: the cluster kernel runs on this device and agrees with the one-CTA kernels
    python tools/probe_cluster_finder.py --part many_mode   # exit 0: the tensor-core many-mode full gradient (second pass rewritten in round 2)

The cluster kernel is the default dispatch for few chains and few modes and has never executed on hardware (round 2 had no GPU access).
A device fault is sticky for the process that raised it, so bench.py and the GPU test session ask THIS script first, in a child
process: when it does not exit 0 they pin "mode_finder_split" = 1 and "chain_split" = 1 (the kernels measured in round 1) for
everything else they do, and only the cluster kernel's own tests / bench entries report the failure.  Prints one JSON line.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def many_mode() -> int:
    from tests import test_fullgrad_tc_gpu as F
    from tests.cases import make_case
    worst = max(F.many_mode_check(make_case("logistic54"), 20, n_opt=3), F.many_mode_check(make_case("logistic1024_n200_N515"), 128, n_opt=3))
    ok = bool(worst < 1e-5)
    print(json.dumps({"ok": ok, "part": "many_mode", "worst_scaled_err": float(worst)}))
    return 0 if ok else 1


def main() -> int:
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"ok": False, "why": "no CUDA device"}))
        return 2
    if "--part" in sys.argv and sys.argv[sys.argv.index("--part") + 1] == "many_mode":
        return many_mode()
    import sgmcmc_b200 as S
    from sgmcmc_b200 import validation as V
    from tests.cases import device_setup, make_case
    out = {"ok": True, "part": "cluster", "checks": []}
    for case_name, split in (("logistic54", 8), ("logistic54_n101", 4), ("logistic130", 2), ("logistic1024_n1024_N3000", 8)):
        case = make_case(case_name)
        modes = []
        for s in (1, split):
            dev = device_setup(case, "sgld", True, n_chains=2, n_iters_opt=60, sharedMode=True, pin_finder=False)
            V.setOption(dev, "mode_finder_split", s)
            S.initSess(dev, verbose=False)
            modes.append(V.getMode(dev)[0])
            del dev
        err = float(np.abs(modes[0] - modes[1]).max() / max(np.abs(modes[0]).max(), 1e-30))
        out["checks"].append({"case": case_name, "split": split, "rel_err_vs_one_cta": err})
        if not (err <= 1e-5):
            out["ok"] = False
            out["why"] = f"{case_name}: cluster of {split} CTAs differs from one CTA per mode by {err:.2e}"
            break
    # sampling on a cluster per chain: 60 stored iterations of every sampler against the one-CTA kernels (same rows, same noise streams)
    from sgmcmc_b200 import _lib
    for case_name, algo, cv, split in (("logistic54", "sgld", True, 8), ("logistic54_n101", "sghmc", False, 4), ("logistic130", "sgnht", True, 2),
                                       ("logistic1024_n1024_N3000", "sgld", True, 8)):
        if not out["ok"]:
            break
        case = make_case(case_name)
        chains = []
        for s in (1, split):
            dev = device_setup(case, algo, cv, n_chains=2, n_iters_opt=20, pin_finder=False)
            V.setOption(dev, "mode_finder_split", 1)
            V.setOption(dev, "chain_split", s)
            S.initSess(dev, verbose=False)
            _lib.check(_lib.load().sgm_run(dev.handle, 60, 1))
            ch = S.getChains(dev)
            chains.append(np.concatenate([np.asarray(v).reshape(2, 60, -1) for v in ch.values()], axis=2))
            del dev
        err = float(np.abs(chains[0] - chains[1]).max() / max(np.abs(chains[0]).max(), 1e-30))
        out["checks"].append({"case": case_name, "sampler": algo, "split": split, "rel_err_vs_one_cta": err})
        if not (err <= 1e-5):
            out["ok"] = False
            out["why"] = f"{case_name} {algo}: chains on clusters of {split} CTAs differ from one CTA per chain by {err:.2e}"
    print(json.dumps(out))
    return 0 if out["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
