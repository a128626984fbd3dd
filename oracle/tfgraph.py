"""Op-by-op float32 replica of the reference's TensorFlow graph (torch CPU autograd).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference differentiates user closures with ``tf$gradients`` (R/setup.R:36-62).
TensorFlow/TFP are absent here, so the closures of the four vignette families are
re-expressed with the same primitive op sequence in torch float32 and
differentiated by torch autograd -- the stand-in for "the reference's TensorFlow
CPU graph" in the parity tests, and the arithmetic timed by the CPU baseline.

Op chains follow TFP's implementations:
  * Normal.log_prob(x)   = -0.5*squared_difference(x/scale, loc/scale)
                           - (0.5*log(2*pi) + log(scale))
  * MultivariateNormalDiag / FullCovariance = affine bijector over a standard
    normal: y = solve(scale, x - loc); sum(Normal(0,1).log_prob(y)) - log|det scale|
  * Mixture.log_prob     = logsumexp_k(component_k.log_prob(x) + log_softmax(logits)_k)
  * Gamma(1,1).log_prob  = (c-1)*log(x) - rate*x + c*log(rate) - lgamma(c)
and the closures in vignettes/{mvGauss,gaussMixture,logisticRegression,nn}.Rmd and
tests/testthat/test{1d,3d}Gauss.R.
"""
from __future__ import annotations

import math

import numpy as np
import torch

HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
F32 = torch.float32


def _normal_log_prob(x, loc, scale):
    scale = torch.as_tensor(scale, dtype=F32)
    loc = torch.as_tensor(loc, dtype=F32)
    d = x / scale - loc / scale
    return -0.5 * (d * d) - (HALF_LOG_2PI + torch.log(scale))


def _mvn_tril_log_prob(x, loc, scale_tril):
    # y = L^{-1}(x - loc); standard-normal log-density summed over event dims; -log|det L|
    r = (x - loc).reshape(-1, scale_tril.shape[0])
    y = torch.linalg.solve_triangular(scale_tril, r.T, upper=False).T
    lp = (-0.5 * (y * y) - HALF_LOG_2PI).sum(dim=1)
    return lp - torch.log(torch.diagonal(scale_tril).abs()).sum()


class TFGaussMean:
    def __init__(self, model):
        self.m = model
        cov = np.linalg.inv(model.precision)
        self.diag = np.allclose(cov, np.diag(np.diag(cov)))
        self.scale_tril = torch.tensor(np.linalg.cholesky(cov), dtype=F32)

    def loglik(self, p, b):
        th = p["theta"].reshape(-1)
        X = b["X"].reshape(-1, self.m.d)
        if self.m.d == 1:  # tests/testthat/test1dGauss.R:19-22
            return _normal_log_prob(X.reshape(-1), th.reshape(()), 1.0).sum()
        return _mvn_tril_log_prob(X, th, self.scale_tril).sum()

    def logprior(self, p):
        if self.m.prior_sd is None:
            return None
        return _normal_log_prob(p["theta"], 0.0, self.m.prior_sd).sum()


class TFMixture2:
    def __init__(self, model):
        self.m = model

    def loglik(self, p, b):
        X = b["X"]
        eye = torch.eye(self.m.d, dtype=F32)
        logits = torch.tensor([0.5, 0.5], dtype=F32)
        logw = torch.log_softmax(logits, dim=0)
        comps = [_mvn_tril_log_prob(X, p[k].reshape(-1), eye) + logw[j]
                 for j, k in enumerate(self.m.names)]
        return torch.logsumexp(torch.stack(comps, dim=0), dim=0).sum()

    def logprior(self, p):
        if self.m.prior_sd is None:
            return None
        L = torch.eye(self.m.d, dtype=F32) * float(self.m.prior_sd)
        zero = torch.zeros(self.m.d, dtype=F32)
        return sum(_mvn_tril_log_prob(p[k].reshape(1, -1), zero, L).sum() for k in self.m.names)


class TFLogistic:
    def __init__(self, model):
        self.m = model

    def loglik(self, p, b):
        z = (p["bias"] + b["X"] @ p["beta"].reshape(-1, 1)).squeeze(-1)
        yhat = 1.0 / (1.0 + torch.exp(-z))
        y = b["y"].reshape(-1)
        return (y * torch.log(yhat) + (1.0 - y) * torch.log(1.0 - yhat)).sum()

    def logprior(self, p):
        if not self.m.laplace_prior:
            return None
        return -(p["beta"].abs().sum() + p["bias"].abs().sum())


class TFSoftmaxNet:
    def __init__(self, model):
        self.m = model

    def loglik(self, p, b):
        beta = torch.softmax(b["X"] @ p["B"] + p["b"], dim=1)
        beta = torch.softmax(beta @ p["A"] + p["a"], dim=1)
        return (b["y"] * torch.log(beta)).sum()

    def logprior(self, p):
        tot = 0.0
        for w, l in self.m._lam.items():
            lam = p[l].reshape(())
            gamma_lp = (1.0 - 1.0) * torch.log(lam) - 1.0 * lam + 1.0 * math.log(1.0) - math.lgamma(1.0)
            tot = tot + _normal_log_prob(p[w], 0.0, torch.rsqrt(lam)).sum() + gamma_lp
        return tot


def replica_for(model):
    from . import models as M
    return {M.GaussMean: TFGaussMean, M.Mixture2: TFMixture2, M.Logistic: TFLogistic,
            M.SoftmaxNet: TFSoftmaxNet}[type(model)](model)


class GraphPosterior:
    """estLogPost graph + tf$gradients of it (R/setup.R:36-42,116-125), float32."""

    def __init__(self, model, N: int, n: int, use_prior: bool = True):
        self.g = replica_for(model)
        self.correction = torch.tensor(N / n, dtype=F32)
        self.use_prior = use_prior

    def _t(self, d, grad):
        return {k: torch.tensor(np.asarray(v, dtype=np.float32), requires_grad=grad) for k, v in d.items()}

    def _post(self, p, b, corrected=True):
        ll = self.g.loglik(p, b)
        lp = self.g.logprior(p) if self.use_prior else None
        val = self.correction * ll if corrected else ll
        return val if lp is None else lp + val

    def est_logpost(self, params, batch):
        with torch.no_grad():
            return float(self._post(self._t(params, False), self._t(batch, False)))

    def _grads(self, params, batch, corrected):
        p = self._t(params, True)
        val = self._post(p, self._t(batch, False), corrected)
        gs = torch.autograd.grad(val, list(p.values()), allow_unused=True)
        return {k: (np.zeros(np.shape(params[k]), np.float32) if g is None else g.numpy().astype(np.float32))
                for k, g in zip(p.keys(), gs)}

    def grad(self, params, batch):
        return self._grads(params, batch, True)

    def full_grad(self, params, data):
        """fullLogPostOpt gradient: no N/n factor (R/controlVariates.R:55-63,96)."""
        return self._grads(params, data, False)
