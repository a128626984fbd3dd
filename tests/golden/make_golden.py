"""Generates tests/golden/golden_v1.npz from the CPU oracle.

The reference (R + TensorFlow) cannot run in this image and ships no golden vectors
for this path, so these fixtures pin OUR oracle: float64 closed form
outputs on seeded inputs.  Regenerate with
    python tests/golden/make_golden.py
Any change in the stored numbers is an oracle behaviour change and must be explained.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import samplers as OS  # noqa: E402
from tests.cases import ALGOS, make_case, oracle_chain  # noqa: E402

CASES = ["gauss1d", "gauss2d", "gauss3d_fullcov", "mixture", "logistic5", "logistic54", "softmaxnet_small"]
STEPS, N_OPT, L = 3, 12, 3


def inputs(case, algo, seed):
    rs = np.random.RandomState(seed)
    P = sum(int(np.size(v)) for v in case["params"].values())
    n_sub = L if algo == "sghmc" else 1
    nl = (L + 1) * P if algo == "sghmc" else P
    mom = rs.randn(P).astype(np.float32)
    opt_idx = np.stack([rs.choice(case["N"], case["n"], replace=False) for _ in range(N_OPT)]).astype(np.int32)
    idx = np.stack([np.stack([rs.choice(case["N"], case["n"], replace=False) for _ in range(n_sub)]) for _ in range(STEPS)]).astype(np.int32)
    noise = rs.randn(STEPS, nl).astype(np.float32)
    return mom, opt_idx, idx, noise


def run(case, algo, cv, dtype, graph=False, seed=101):
    mom, opt_idx, idx, noise = inputs(case, algo, seed)
    ch = oracle_chain(case, algo, cv, dtype, L=L, graph=graph)
    if algo != "sgld":
        ch.init_momentum(mom)
    if cv:
        ch.find_mode(opt_idx)
    for s in range(STEPS):
        ch.step(idx[s], noise[s])
    return ch


def main():
    out = {}
    for cn in CASES:
        case = make_case(cn)
        for algo in ALGOS:
            for cv in (False, True):
                ch = run(case, algo, cv, np.float64)
                key = f"{cn}/{algo}{'cv' if cv else ''}"
                out[key + "/theta"] = OS.flat(ch.theta)
                out[key + "/mom"] = OS.flat(ch.mom)
                out[key + "/thermo"] = np.array([ch.thermo[k] for k in ch.theta], dtype=np.float64)
                if cv:
                    out[key + "/theta_hat"] = OS.flat(ch.theta_hat)
                    out[key + "/g_hat"] = OS.flat(ch.g_hat)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
