"""CPU suite: the oracle against itself (two independent restatements), its committed
golden vectors, the reference's statistical acceptance bounds, and the RNG spec."""
import os

import numpy as np
import pytest

from oracle import models as OM
from oracle import rng
from oracle import samplers as OS
from oracle import tfgraph as TG
from tests.cases import ALGOS, make_case
from tests.golden import make_golden as MG

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = rng.philox4x32(*[np.uint32(c) for c in ctr], *key)
        assert tuple(int(g) for g in got) == want


@pytest.mark.parametrize("N,n", [(1, 1), (2, 2), (7, 7), (100, 10), (10 ** 4, 100), (581012, 500), (2 ** 20 + 3, 1024)])
def test_minibatch_is_without_replacement_and_in_range(N, n):
    """sample(N, n) semantics (R/update.R:171): n distinct rows of [0, N)."""
    for step in range(5):
        idx = rng.minibatch_indices(13, 3, step, rng.TAG_INDEX, N, n)
        assert idx.shape == (n,) and idx.min() >= 0 and idx.max() < N
        assert len(set(idx.tolist())) == n


def test_minibatch_inclusion_is_uniform():
    """Every row is included with probability n/N; pairs of positions are uncorrelated."""
    N, n, reps = 500, 50, 4000
    counts = np.zeros(N)
    first_two = np.zeros((reps, 2))
    for r in range(reps):
        idx = rng.minibatch_indices(99, 0, r, rng.TAG_INDEX, N, n)
        counts[idx] += 1
        first_two[r] = idx[:2]
    expect = reps * n / N
    chi2 = ((counts - expect) ** 2 / (expect * (1 - n / N))).sum()
    assert abs(chi2 - N) < 6 * np.sqrt(2 * N), chi2            # chi-square with ~N dof
    assert abs(np.corrcoef(first_two.T)[0, 1]) < 0.06
    # different chains / steps give different minibatches
    a = rng.minibatch_indices(99, 0, 0, rng.TAG_INDEX, N, n)
    b = rng.minibatch_indices(99, 1, 0, rng.TAG_INDEX, N, n)
    assert not np.array_equal(a, b)


def test_normals_moments():
    z = rng.normals(5, 2, 17, rng.TAG_NOISE, 400000).astype(np.float64)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3) < 0.06
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 0.01


@pytest.mark.parametrize("case_name", ["gauss1d", "gauss2d", "gauss3d_fullcov", "gauss2d_flat", "mixture",
                                       "logistic54", "softmaxnet_small"])
def test_closed_form_matches_graph_replica(case_name):
    """Closed-form gradients == autograd of the op-by-op TensorFlow-graph replica."""
    case = make_case(case_name)
    rs = np.random.RandomState(1)
    use_prior = case["logprior"] is not None
    p64 = OS.Posterior(case["om"], case["N"], case["n"], np.float64, use_prior)
    p32 = OS.Posterior(case["om"], case["N"], case["n"], np.float32, use_prior)
    gp = TG.GraphPosterior(case["om"], case["N"], case["n"], use_prior)
    batch = OS.take(case["data"], rs.choice(case["N"], case["n"], replace=False))
    g64, g32, gg = p64.grad(case["params"], batch), p32.grad(case["params"], batch), gp.grad(case["params"], batch)
    for k in g64:
        s = np.abs(g64[k]).max() + 1e-12
        assert np.abs(gg[k] - g64[k]).max() <= 2e-5 * s, k
        assert np.abs(g32[k] - g64[k]).max() <= 2e-5 * s, k
    v64, vg = p64.est_logpost(case["params"], batch), gp.est_logpost(case["params"], batch)
    assert abs(v64 - vg) <= 1e-5 * abs(v64)
    f64, fg = p64.full_grad(case["params"], case["data"]), gp.full_grad(case["params"], case["data"])
    for k in f64:
        gl = case["om"].grad_loglik(case["params"], case["data"], np.float64)[k]
        s = max(np.abs(f64[k]).max(), np.abs(gl).max()) + 1e-12
        assert np.abs(fg[k] - f64[k]).max() <= 1e-4 * s, k


@pytest.mark.parametrize("case_name", MG.CASES)
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("cv", [False, True])
def test_oracle_reproduces_golden(case_name, algo, cv):
    """float64 closed form reproduces the committed fixture exactly (to 1e-12); the float32
    graph replica stays within fp32 distance of it."""
    case = make_case(case_name)
    key = f"{case_name}/{algo}{'cv' if cv else ''}"
    ch = MG.run(case, algo, cv, np.float64)
    np.testing.assert_allclose(OS.flat(ch.theta), GOLDEN[key + "/theta"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(OS.flat(ch.mom), GOLDEN[key + "/mom"], rtol=1e-10, atol=1e-14)
    if case_name in ("gauss2d", "mixture", "logistic5"):
        ch32 = MG.run(case, algo, cv, np.float32, graph=True)
        ref = GOLDEN[key + "/theta"]
        tol = 1e-5 if not cv else 1e-3  # the reference-literal CV subtraction cancels in fp32
        assert np.abs(OS.flat(ch32.theta) - ref).max() <= tol * np.abs(ref).max() + 1e-7


@pytest.mark.parametrize("algo,cv,eps", [("sgld", False, 1e-4), ("sgld", True, 1e-4), ("sgnht", True, 1e-4)])
def test_reference_statistical_bounds_1d_gauss(algo, cv, eps):
    """tests/testthat/test1dGauss.R:92-95 on the oracle: chain stays in [-0.5, 0.5], mean^2 <= 0.1."""
    rs = np.random.RandomState(13)
    N, n, nIters, burn = 10 ** 4, 100, 1100, 100
    data = {"X": rs.randn(N)}
    post = OS.Posterior(OM.GaussMean(1, prior_sd=5.0), N, n, np.float64)
    ch = OS.Chain(post, data, {"theta": np.array(rs.randn())}, algo, eps, cv=cv, a=0.01, opt_stepsize=1e-5)
    if algo != "sgld":
        ch.init_momentum(rs.randn(1))
    if cv:
        ch.find_mode([rs.choice(N, n, replace=False) for _ in range(1000)])
    out = []
    for _ in range(nIters):
        ch.step(rs.choice(N, n, replace=False), rs.randn(ch.noise_len()))
        out.append(float(ch.theta["theta"]))
    out = np.array(out[burn:])
    assert out.min() >= -0.5 and out.max() <= 0.5 and (out ** 2).mean() <= 0.1


def test_sghmc_refresh_and_substeps_consume_expected_inputs():
    case = make_case("mixture")
    ch = OS.Chain(OS.Posterior(case["om"], case["N"], case["n"]), case["data"], case["params"], "sghmc", 2e-5, L=5)
    assert ch.n_sub() == 5 and ch.noise_len() == 6 * 4
