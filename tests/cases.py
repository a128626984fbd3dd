"""Shared seeded test cases: (oracle model, device model spec, data, params, tuning)."""
from __future__ import annotations

import numpy as np

from oracle import models as OM

SIGMA3 = np.array([[1, 0.3, 0.5], [0.3, 1, 0.2], [0.5, 0.2, 1.0]])


def make_case(name: str, seed: int = 13):
    """Returns dict(oracle_model, loglik, logprior, data, params, N, n, eps, opt_eps)."""
    import sgmcmc_b200 as S
    rs = np.random.RandomState(seed)
    if name == "gauss1d":  # tests/testthat/test1dGauss.R
        N = 2000
        return dict(om=OM.GaussMean(1, prior_sd=5.0), loglik=S.gaussMeanLogLik(), logprior=S.normalPrior(5.0),
                    data={"X": rs.randn(N)}, params={"theta": np.array(rs.randn())}, N=N, n=50, eps=1e-4, opt_eps=1e-5)
    if name == "gauss2d":  # vignettes/mvGauss.Rmd
        N = 3000
        return dict(om=OM.GaussMean(2, prior_sd=10.0), loglik=S.gaussMeanLogLik(), logprior=S.normalPrior(10.0),
                    data={"X": rs.randn(N, 2)}, params={"theta": np.array([0.3, -0.2])}, N=N, n=100, eps=1e-5, opt_eps=1e-5)
    if name == "gauss3d_fullcov":  # tests/testthat/test3dGauss.R
        N = 2500
        X = rs.multivariate_normal(np.zeros(3), SIGMA3, N)
        return dict(om=OM.GaussMean(3, precision=np.linalg.inv(SIGMA3), prior_sd=5.0), loglik=S.gaussMeanLogLik(SIGMA3),
                    logprior=S.normalPrior(5.0), data={"X": X}, params={"theta": rs.randn(3)}, N=N, n=100, eps=1e-4, opt_eps=1e-5)
    if name == "gauss2d_flat":  # tests/testthat/testOptional.R (logPrior = NULL)
        N = 1500
        return dict(om=OM.GaussMean(2, prior_sd=None), loglik=S.gaussMeanLogLik(), logprior=None,
                    data={"X": rs.randn(N, 2)}, params={"theta": np.array([0.1, 0.1])}, N=N, n=15, eps=1e-4, opt_eps=1e-5)
    if name == "mixture":  # vignettes/gaussMixture.Rmd
        N = 3000
        z = rs.rand(N) < 0.5
        X = rs.randn(N, 2) + np.where(z[:, None], 0.0, 0.1)
        return dict(om=OM.Mixture2(2, 10.0), loglik=S.mixture2LogLik(), logprior=S.normalPrior(10.0),
                    data={"X": X}, params={"theta1": np.array([0.0, 0.0]), "theta2": np.array([0.1, 0.1])},
                    N=N, n=100, eps=2e-5, opt_eps=1e-5)
    import re
    m = re.fullmatch(r"gauss(\d+)d(_full)?_n(\d+)_N(\d+)", name)   # random shapes (tests/fuzz_emu.py): "gauss<d>d[_full]_n<minibatch>_N<rows>"
    if m:
        d, full, n, N = int(m.group(1)), bool(m.group(2)), int(m.group(3)), int(m.group(4))
        if full:
            A = rs.randn(d, d) * 0.3 + np.eye(d)
            Sig = A @ A.T
            X = rs.multivariate_normal(np.zeros(d), Sig, N)
            return dict(om=OM.GaussMean(d, precision=np.linalg.inv(Sig), prior_sd=5.0), loglik=S.gaussMeanLogLik(Sig), logprior=S.normalPrior(5.0),
                        data={"X": X}, params={"theta": rs.randn(d)}, N=N, n=n, eps=1e-4, opt_eps=1e-5)
        return dict(om=OM.GaussMean(d, prior_sd=10.0), loglik=S.gaussMeanLogLik(), logprior=S.normalPrior(10.0),
                    data={"X": rs.randn(N, d)}, params={"theta": rs.randn(d) * 0.3}, N=N, n=n, eps=1e-5, opt_eps=1e-5)
    m = re.fullmatch(r"mixture(\d+)d_n(\d+)_N(\d+)", name)
    if m:
        d, n, N = int(m.group(1)), int(m.group(2)), int(m.group(3))
        z = rs.rand(N) < 0.5
        X = rs.randn(N, d) + np.where(z[:, None], 0.0, 0.1)
        return dict(om=OM.Mixture2(d, 10.0), loglik=S.mixture2LogLik(), logprior=S.normalPrior(10.0),
                    data={"X": X}, params={"theta1": np.zeros(d), "theta2": np.full(d, 0.1)}, N=N, n=n, eps=2e-5, opt_eps=1e-5)
    m = re.fullmatch(r"softmaxnet_(\d+)_(\d+)_(\d+)_n(\d+)_N(\d+)", name)
    if m:
        din, dh, do, n, N = (int(g) for g in m.groups())
        X = rs.rand(N, din)
        Y = np.eye(do)[rs.randint(0, do, N)]
        p = {"A": rs.randn(dh, do), "B": rs.randn(din, dh), "a": rs.randn(do), "b": rs.randn(dh),
             "lambdaA": np.array(1.3), "lambdaB": np.array(0.7), "lambdaa": np.array(2.0), "lambdab": np.array(0.5)}
        return dict(om=OM.SoftmaxNet(din, dh, do), loglik=S.softmaxNetLogLik(), logprior=S.gammaPrecisionPrior(),
                    data={"X": X, "y": Y}, params=p, N=N, n=n, eps=1e-5, opt_eps=1e-6)
    if name.startswith("logistic"):  # vignettes/logisticRegression.Rmd; "logistic<d>[_n<minibatch>][_N<rows>]"
        m = re.fullmatch(r"logistic(\d+)(?:_n(\d+))?(?:_N(\d+))?", name)
        d = int(m.group(1))
        n = int(m.group(2)) if m.group(2) else 200
        N = int(m.group(3)) if m.group(3) else 4000
        X = rs.randn(N, d)
        beta = rs.randn(d) * 0.5 / np.sqrt(d)
        y = (rs.rand(N) < 1 / (1 + np.exp(-(0.2 + X @ beta)))).astype(np.float64)
        return dict(om=OM.Logistic(d), loglik=S.logisticLogLik(), logprior=S.laplacePrior(),
                    data={"X": X, "y": y}, params={"bias": np.array(0.05), "beta": (rs.randn(d, 1) * 0.05)},
                    N=N, n=n, eps=5e-6, opt_eps=5e-6)
    if name == "softmaxnet_small":  # vignettes/nn.Rmd at reduced width
        N, din, dh, do = 600, 20, 8, 4
        X = rs.rand(N, din)
        Y = np.eye(do)[rs.randint(0, do, N)]
        p = {"A": rs.randn(dh, do), "B": rs.randn(din, dh), "a": rs.randn(do), "b": rs.randn(dh),
             "lambdaA": np.array(1.3), "lambdaB": np.array(0.7), "lambdaa": np.array(2.0), "lambdab": np.array(0.5)}
        return dict(om=OM.SoftmaxNet(din, dh, do), loglik=S.softmaxNetLogLik(), logprior=S.gammaPrecisionPrior(),
                    data={"X": X, "y": Y}, params=p, N=N, n=40, eps=1e-5, opt_eps=1e-6)
    if name == "softmaxnet":  # vignettes/nn.Rmd 784-100-10
        N, din, dh, do = 1000, 784, 100, 10
        X = rs.rand(N, din)
        Y = np.eye(do)[rs.randint(0, do, N)]
        p = {"A": rs.randn(dh, do), "B": rs.randn(din, dh), "a": rs.randn(do), "b": rs.randn(dh),
             "lambdaA": np.array(1.3), "lambdaB": np.array(0.7), "lambdaa": np.array(2.0), "lambdab": np.array(0.5)}
        return dict(om=OM.SoftmaxNet(din, dh, do), loglik=S.softmaxNetLogLik(), logprior=S.gammaPrecisionPrior(),
                    data={"X": X, "y": Y}, params=p, N=N, n=100, eps=1e-5, opt_eps=1e-6)
    raise KeyError(name)


ALGOS = ("sgld", "sghmc", "sgnht")

# GPU sessions (tests/conftest.py): set when tools/probe_cluster_finder.py did not exit 0 in its own process; the cluster kernel's own
# tests then fail with the probe's output instead of raising a sticky device fault inside the pytest process.
CLUSTER_PROBE_ERROR = None


def device_setup(case, algo, cv, n_chains=1, seed=7, L=3, n_iters_opt=25, pin_finder=True, **kw):
    """pin_finder (default): logistic samplers keep to one CTA per chain and per mode -- the kernels with a record on hardware -- so that
    the general tests judge those; the cluster kernel (the library's default dispatch for few chains / modes, never run on hardware
    before this round's first GPU session) is exercised by the tests that select it ("chain_split" / "mode_finder_split", or
    pin_finder=False for the default dispatch), which run last (tests/conftest.py)."""
    import sgmcmc_b200 as S
    if pin_finder and "bias" in case["params"]:
        from sgmcmc_b200 import validation as V
        dev = device_setup(case, algo, cv, n_chains=n_chains, seed=seed, L=L, n_iters_opt=n_iters_opt, pin_finder=False, **kw)
        V.setOption(dev, "mode_finder_split", 1)
        V.setOption(dev, "chain_split", 1)
        return dev
    fn = {("sgld", False): S.sgldSetup, ("sghmc", False): S.sghmcSetup, ("sgnht", False): S.sgnhtSetup,
          ("sgld", True): S.sgldcvSetup, ("sghmc", True): S.sghmccvSetup, ("sgnht", True): S.sgnhtcvSetup}[(algo, cv)]
    args = dict(logPrior=case["logprior"], minibatchSize=case["n"], seed=seed, nChains=n_chains, **kw)
    if algo == "sghmc":
        args.update(alpha=0.1, L=L)
    if algo == "sgnht":
        args.update(a=0.01)
    if cv:
        return fn(case["loglik"], case["data"], case["params"], case["eps"], case["opt_eps"], nItersOpt=n_iters_opt, **args)
    return fn(case["loglik"], case["data"], case["params"], case["eps"], **args)


def oracle_chain(case, algo, cv, dtype, L=3, graph=False):
    from oracle import samplers as OS
    from oracle import tfgraph as TG
    post = OS.Posterior(case["om"], case["N"], case["n"], dtype, use_prior=case["logprior"] is not None)
    kw = {}
    if graph:
        gp = TG.GraphPosterior(case["om"], case["N"], case["n"], use_prior=case["logprior"] is not None)
        kw = dict(grad_fn=gp.grad, full_grad_fn=gp.full_grad)
    return OS.Chain(post, case["data"], case["params"], algo, case["eps"], cv=cv, alpha=0.1, a=0.01, L=L,
                    opt_stepsize=case["opt_eps"], **kw)
