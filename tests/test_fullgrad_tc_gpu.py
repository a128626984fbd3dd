"""-m gpu: control-variate full-data gradient at MANY modes on the tensor cores (csrc/logistic_cv_gemm.cuh, 3xTF32
tcgen05 contractions) against the float64 oracle and against the SIMT row-chunk x mode kernel, through the C ABI.

Reference path: calcFullGrads (R/controlVariates.R:105-111) once per independent chain.  Tolerance: 1e-5 relative to the
per-datum gradient mass (the total is near zero at a mode), the same scale tests/test_parity_gpu.py uses for g_hat.
"""
import numpy as np
import pytest

from oracle import samplers as OS
from tests.cases import device_setup, make_case

pytestmark = pytest.mark.gpu


FULLGRAD_OPTION = {"tc": 1, "generic": 2}   # sgm_set_option(h, "fullgrad", ...)


def _mode_and_grad(case, chains, path, n_opt=30):
    from sgmcmc_b200 import validation as V
    import sgmcmc_b200 as S
    dev = device_setup(case, "sgld", True, n_chains=chains, n_iters_opt=n_opt)
    V.setOption(dev, "fullgrad", FULLGRAD_OPTION[path])
    S.initSess(dev, verbose=False)
    if path == "tc":
        assert V.lastFullGradFlops(dev) > 0, "the tensor-core path did not run"
    return V.getMode(dev)


@pytest.mark.parametrize("case_name,chains", [("logistic54", 20), ("logistic54", 150), ("logistic130", 40), ("logistic5", 16)])
def test_many_mode_full_gradient_matches_oracle(lib, case_name, chains):
    case = make_case(case_name)
    th_tc, g_tc = _mode_and_grad(case, chains, "tc")
    th_sm, g_sm = _mode_and_grad(case, chains, "generic")
    # same seed, same device RNG streams: the modes are identical, so the two gradient paths see the same input
    assert np.array_equal(th_tc, th_sm)
    assert len({tuple(t) for t in th_tc}) == chains, "every chain finds its own mode"
    post = OS.Posterior(case["om"], case["N"], case["n"], np.float64, use_prior=True)
    worst = 0.0
    for c in range(chains):
        th = OS.unflat(th_tc[c], case["params"])
        ref = OS.flat(post.full_grad(th, case["data"]))
        gl = OS.flat(case["om"].grad_loglik(th, case["data"], np.float64))
        # mass of the per-datum summands: sum_i |y_i - s_i| |x_ik| >= |sum|; use the data scale N * mean|x| as in test_parity_gpu
        scale = max(np.abs(ref).max(), np.abs(gl).max(), 1.0) * 10
        err = np.abs(g_tc[c] - ref).max()
        worst = max(worst, err / scale)
        assert err <= 1e-5 * scale + 1e-7, f"chain {c}: tensor-core g_hat err {err:.3e} vs scale {scale:.3e}"
        err_sm = np.abs(g_sm[c] - ref).max()
        assert err <= err_sm + 1e-5 * scale, f"chain {c}: tensor-core path ({err:.2e}) worse than the SIMT path ({err_sm:.2e})"
    print(f"{case_name} x {chains} modes: worst scaled error {worst:.2e}")


def test_many_mode_full_gradient_row_tail_and_blocks(lib):
    """N not a multiple of the 128-row tile or of the 256-row K segment; more than one K split."""
    case = make_case("logistic54")
    keep = 3877
    case["data"] = {k: v[:keep] for k, v in case["data"].items()}
    case["N"] = keep
    th_tc, g_tc = _mode_and_grad(case, 24, "tc", n_opt=10)
    post = OS.Posterior(case["om"], case["N"], case["n"], np.float64, use_prior=True)
    for c in range(24):
        th = OS.unflat(th_tc[c], case["params"])
        ref = OS.flat(post.full_grad(th, case["data"]))
        gl = OS.flat(case["om"].grad_loglik(th, case["data"], np.float64))
        scale = max(np.abs(ref).max(), np.abs(gl).max(), 1.0) * 10
        assert np.abs(g_tc[c] - ref).max() <= 1e-5 * scale + 1e-7


@pytest.mark.parametrize("case_name,chains", [("logistic1024_n200_N4099", 256)])
def test_many_mode_full_gradient_at_the_benchmarked_width(lib, case_name, chains):
    """d = 1024 (the c5 shape bench.py times): pass 1 walks 32 K stages = 4 TMEM segments per item (buffer alternation, register
    accumulation across segments), 2 mode tiles, a row tail (4099 = 32 x 128 + 3) and 8 feature tiles in pass 2."""
    many_mode_check(make_case(case_name), chains, n_opt=4)


def many_mode_check(case, chains, n_opt=4):
    th_tc, g_tc = _mode_and_grad(case, chains, "tc", n_opt=n_opt)
    post = OS.Posterior(case["om"], case["N"], case["n"], np.float64, use_prior=True)
    worst = 0.0
    for c in range(chains):
        th = OS.unflat(th_tc[c], case["params"])
        ref = OS.flat(post.full_grad(th, case["data"]))
        gl = OS.flat(case["om"].grad_loglik(th, case["data"], np.float64))
        scale = max(np.abs(ref).max(), np.abs(gl).max(), 1.0) * 10
        err = np.abs(g_tc[c] - ref).max()
        worst = max(worst, err / scale)
        assert err <= 1e-5 * scale + 1e-7, f"chain {c}: tensor-core g_hat err {err:.3e} vs scale {scale:.3e}"
    return worst


ROW_RANGES = [(1, 700), (37, 37 + 1153), (2001, 4000), (333, 334), (3, 131), (1939, 3878)]


@pytest.mark.parametrize("case_name,chains", [("logistic128_n20_N300", 17), ("logistic256_n20_N300", 16), ("logistic127_n20_N333", 33), ("logistic129_n20_N259", 130),
                                              ("logistic3_n5_N100", 16), ("logistic124_n9_N131", 20)])
def test_many_mode_bias_component_either_way(lib, case_name, chains):
    """The bias component of the gradient, sum_i R[i][m]: carried by pass 2 as a constant-1 row of its last feature tile when that tile has a
    free row (d % 128 != 0: d = 3, 124, 127, 129), summed from R^T by cvg_rowsum_kernel when the tiles are full (d = 128, 256)."""
    assert many_mode_check(make_case(case_name), chains, n_opt=3) < 1e-5


@pytest.mark.parametrize("path", ["tc", "generic", "stream"])
def test_full_gradient_over_a_row_range_that_starts_anywhere(lib, path):
    """What rank r of a row-sharded job computes: the sum over rows [row0, row1) only, row0 odd / not a multiple of 4 or 32 (the
    tensor-core pass walks K from the aligned row below and masks the rows before the range: TMA box starts must be 16-byte aligned).
    Selected through sgm_set_option("fullgrad_row0" / "fullgrad_row1"), no communicator needed: a single-GPU pre-flight of the
    shard starts that faulted at 8 GPUs in round 1."""
    from sgmcmc_b200 import validation as V
    import sgmcmc_b200 as S
    case = make_case("logistic54")
    chains = 1 if path == "stream" else 20
    post = OS.Posterior(case["om"], case["N"], case["n"], np.float64, use_prior=True)
    for row0, row1 in ROW_RANGES:
        dev = device_setup(case, "sgld", True, n_chains=chains, n_iters_opt=3)
        if path != "stream":
            V.setOption(dev, "fullgrad", FULLGRAD_OPTION[path])
        V.setOption(dev, "fullgrad_row0", row0)
        V.setOption(dev, "fullgrad_row1", row1)
        S.initSess(dev, verbose=False)
        th_hat, g_hat = V.getMode(dev)
        sub = {k: np.asarray(v)[row0:row1] for k, v in case["data"].items()}
        for c in range(chains):
            th = OS.unflat(th_hat[c], case["params"])
            ref = OS.flat(post.full_grad(th, sub))
            gl = OS.flat(case["om"].grad_loglik(th, sub, np.float64))
            scale = max(np.abs(ref).max(), np.abs(gl).max(), 1.0) * 10
            err = np.abs(g_hat[c] - ref).max()
            assert err <= 1e-5 * scale + 1e-7, f"rows [{row0}, {row1}), chain {c}, {path}: err {err:.3e} vs scale {scale:.3e}"
