"""TEST INFRASTRUCTURE: random shapes of all four families (dimensions, minibatch n, rows N, sampler, plain / control variates / shared
mode) through the CPU-emulated build against the oracle -- the deterministic-step parity body of tests/test_parity_gpu.py, the streamed
device-RNG run, the tensor-core many-mode full gradient -- at shapes nobody wrote down (d % 4 in {0, 1, 2, 3}, n not a multiple of 4 or 32, n == N, tiny N, row strides either
side of the short-row / long-row dispatch boundary).  Every failure class of the emulation (deadlock, out-of-range access, misaligned
TMA operand) is an exception here.

    python -m tests.fuzz_emu [--cases 40] [--seed 1]          # prints one line per case; exit code 1 on the first failure

tests/test_emu_fuzz.py runs a small fixed-seed batch of it inside the CPU suite."""
from __future__ import annotations

import argparse
import sys
import traceback

import numpy as np


def _rows(rs, n):
    return int(max(n, rs.choice([n, n + 1, n + rs.randint(2, 50), rs.randint(n + 1, n + 1500)])))


def random_case(rs):
    """(kind, case name, sampler, mode).  kind: steps = deterministic-step parity; finder = the same with the mode finder on a thread-block
    cluster of `sampler` CTAs per mode; stream = sgm_sample against sgm_run + sgm_copy_chain
    with the device RNG; cvg = the tensor-core many-mode full gradient."""
    algo = str(rs.choice(["sgld", "sghmc", "sgnht"]))
    mode = str(rs.choice(["plain", "cv", "cv_shared"]))
    fam = rs.choice(["logistic"] * 5 + ["gauss", "mixture", "net", "cvg", "stream", "finder", "finder"])
    if fam == "gauss":
        d, n = int(rs.randint(1, 9)), int(rs.choice([1, 2, 7, 32, 33, 100, 257]))
        return "steps", f"gauss{d}d{'_full' if rs.rand() < 0.5 else ''}_n{n}_N{_rows(rs, n)}", algo, mode
    if fam == "mixture":
        d, n = int(rs.randint(1, 9)), int(rs.choice([1, 3, 31, 32, 64, 100, 130]))
        return "steps", f"mixture{d}d_n{n}_N{_rows(rs, n)}", algo, mode
    if fam == "net":
        din, dh = 4 * int(rs.randint(1, 40)), 4 * int(rs.randint(1, 17))
        do = int(rs.choice([1, 2, 3, 4, 5, 7, 8, 10, 16]))
        n = int(rs.choice([1, 2, 5, 40, 100, 127, 128, 129, 200]))
        return "steps", f"softmaxnet_{din}_{dh}_{do}_n{n}_N{_rows(rs, n)}", algo, mode
    if fam == "cvg":
        d = int(rs.choice([rs.randint(1, 130), rs.randint(130, 400), 1024 + rs.randint(-3, 4)]))
        n = int(rs.randint(1, 60))
        N = int(rs.choice([rs.randint(n + 1, 300), rs.randint(300, 2500)]))
        return "cvg", f"logistic{d}_n{n}_N{N}", "sgld", str(int(rs.choice([16, 17, 31, 64, 129, 150])))
    kind = rs.choice(["short", "short", "long", "edge"])
    if kind == "short":      # rows <= 512 B: logistic_staged_kernel (d <= 123)
        d = int(rs.randint(1, 124))
        n = int(rs.choice([rs.randint(1, 40), rs.randint(40, 300), rs.randint(256, 700)]))
    elif kind == "long":     # long rows: logistic_chain_kernel<G, V>
        d = int(rs.choice([rs.randint(124, 300), rs.randint(300, 700), 1024 + rs.randint(-3, 4)]))
        n = int(rs.randint(1, 130))
    else:
        d = int(rs.choice([1, 2, 3, 4, 5, 7, 123, 124, 127, 128]))
        n = int(rs.choice([1, 2, 3, 31, 32, 33, 255, 256, 257]))
    name = f"logistic{d}_n{n}_N{_rows(rs, n)}"
    if fam == "finder":      # mode finder and sampling steps on a thread-block cluster of a pinned size (csrc/logistic_split.cuh)
        return "finder", name, str(int(rs.choice([2, 4, 8]))), str(rs.choice(["cv", "cv_shared"]))
    if fam == "stream":
        return "stream", name, algo, mode
    return "steps", name, algo, mode


def run_case(kind, name, algo, mode):
    from tests import test_parity_gpu as T
    if kind == "steps":
        T._deterministic_steps(name, algo, mode != "plain", shared=(mode == "cv_shared"), mass_scaled=True)
    elif kind == "finder":
        k = int(algo)
        T._deterministic_steps(name, ("sgld", "sghmc", "sgnht")[k % 3], True, shared=(mode == "cv_shared"), mass_scaled=True,
                               options={"mode_finder_split": k, "chain_split": k}, n_opt=9)
    elif kind == "cvg":
        from tests import test_fullgrad_tc_gpu as F
        from tests.cases import make_case
        F.many_mode_check(make_case(name), int(mode), n_opt=3)
    else:
        import sgmcmc_b200 as S
        from sgmcmc_b200 import _lib, api as A
        from tests.cases import device_setup, make_case
        case = make_case(name)
        cv, shared, iters, chains = mode != "plain", mode == "cv_shared", 57, 3
        one = device_setup(case, algo, cv, n_chains=chains, seed=77, n_iters_opt=7, sharedMode=shared)
        S.initSess(one, verbose=False)
        _lib.check(_lib.load().sgm_run(one.handle, iters, 1))
        want = {k: np.array(v) for k, v in S.getChains(one).items()}
        two = device_setup(case, algo, cv, n_chains=chains, seed=77, n_iters_opt=7, sharedMode=shared)
        got = A.runSGMCMC(two, case["params"], {"nIters": iters, "verbose": False})
        for k in want:
            assert np.array_equal(np.asarray(got[k]), want[k]), f"{k}: streamed chain differs from run + copy"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    from tests.emu import backend
    rs = np.random.RandomState(args.seed)
    bad = 0
    with backend.emulated() as lib:
        for i in range(args.cases):
            kind, name, algo, mode = random_case(rs)
            try:
                lib.emu_reset_error()
                run_case(kind, name, algo, mode)
                msg = lib.emu_last_message()
                if msg:
                    raise RuntimeError(msg.decode())
                print(f"[ok]   {kind} {name} {algo} {mode}", flush=True)
            except Exception as e:   # noqa: BLE001
                bad += 1
                print(f"[FAIL] {kind} {name} {algo} {mode}: {type(e).__name__}: {str(e)[:500]}", flush=True)
                traceback.print_exc(limit=3)
                break
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
