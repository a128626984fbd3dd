"""Restatement of the reference's sampler dynamics and run loop (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference file:line:
  * estimated log posterior  logPrior + (N/n) logLik      R/setup.R:116-125
  * plain / control-variate gradient estimators            R/setup.R:36-62
  * SGLD / SGHMC / SGNHT update ops                        R/dynamics.R:7-120
  * their execution order inside sgmcmcStep                R/update.R:40-81
  * mode finder + full gradient for *cv                    R/controlVariates.R:30-52,85-111
  * run loop / storage                                     R/update.R:145-163, R/storage.R:42-80

Semantics worth restating:
  - SGLD and SGHMC update parameter tensors one ``Session.run`` at a time, so tensor
    p's gradient is evaluated AFTER tensors 1..p-1 were updated (Gauss-Seidel);
    SGNHT updates all momenta before any parameter, so it is simultaneous.
  - SGHMC refreshes the momentum on every call (variance eta/2) and then runs L
    sub-steps, each on its own minibatch.
  - SGNHT keeps ONE scalar thermostat per parameter tensor, initialised to ``a``.
  - *cv chains start at the mode estimate, not at the user's ``params``.

``dtype`` selects the arithmetic: float64 = ground truth, float32 = what the
reference computes (R/setup.R:99,110,117).  Scalar constants are formed in double
(R numerics) and rounded once to ``dtype`` exactly where the reference's graph would
hold a float32 constant.
"""
from __future__ import annotations

import math
from typing import Callable, Dict

import numpy as np

Params = Dict[str, np.ndarray]


def flat(params: Params) -> np.ndarray:
    """Concatenate tensors in names() order, each in R (column-major) element order."""
    return np.concatenate([np.asarray(v).reshape(-1, order="F") for v in params.values()])


def unflat(vec: np.ndarray, like: Params) -> Params:
    out, o = {}, 0
    for k, v in like.items():
        sz = int(np.size(v))
        out[k] = np.asarray(vec[o:o + sz]).reshape(np.shape(v), order="F")
        o += sz
    return out


def split_noise(noise: np.ndarray, like: Params) -> Params:
    return unflat(np.asarray(noise), like)


class Posterior:
    """Estimated log posterior of R/setup.R:116-125 and its gradients for one model."""

    def __init__(self, model, N: int, n: int, dtype=np.float64, use_prior: bool = True):
        self.model, self.N, self.n, self.dtype = model, int(N), int(n), dtype
        self.use_prior = use_prior
        self.correction = dtype(N / n)  # tf$constant(N / n, dtype = tf$float32)

    def est_logpost(self, params: Params, batch) -> float:
        ll = self.model.loglik(params, batch, self.dtype)
        lp = self.model.logprior(params, self.dtype) if self.use_prior else self.dtype(0)
        return self.dtype(lp + self.correction * ll)

    def grad(self, params: Params, batch) -> Params:
        """d estLogPost / d theta_p for every tensor (R/setup.R:36-42)."""
        gl = self.model.grad_loglik(params, batch, self.dtype)
        if self.use_prior:
            gp = self.model.grad_logprior(params, self.dtype)
            return {k: (gp[k] + self.correction * gl[k]).astype(self.dtype) for k in gl}
        return {k: (self.correction * gl[k]).astype(self.dtype) for k in gl}

    def full_grad(self, params: Params, data) -> Params:
        """d [logPrior + logLik(all data)] / d theta_p, no N/n factor (cV.R:55-63,96)."""
        gl = self.model.grad_loglik(params, data, self.dtype)
        if self.use_prior:
            gp = self.model.grad_logprior(params, self.dtype)
            return {k: (gp[k] + gl[k]).astype(self.dtype) for k in gl}
        return {k: gl[k].astype(self.dtype) for k in gl}


def take(data, idx):
    """dataSelect: slice every dataset entry on its first axis (R/update.R:180-194)."""
    return {k: np.asarray(v)[idx] for k, v in data.items()}


class Chain:
    """One Markov chain driven in the deterministic validation mode.

    ``algo`` in {"sgld","sghmc","sgnht"}; ``cv`` adds the control-variate estimator.
    ``stepsize``/``alpha``/``a`` are per-tensor dicts (convertList, R/setup.R:160-170).
    Callers supply 0-based minibatch rows and standard-normal noise; scale factors
    are applied here exactly as in R/dynamics.R.
    """

    def __init__(self, post: Posterior, data, params: Params, algo: str, stepsize,
                 cv: bool = False, alpha=0.01, a=0.01, L: int = 5, opt_stepsize: float = 0.0,
                 grad_fn: Callable | None = None, full_grad_fn: Callable | None = None):
        self.post, self.data, self.algo, self.cv, self.L = post, data, algo, cv, int(L)
        dt = post.dtype
        self.dtype = dt
        self.theta = {k: np.array(v, dtype=dt) for k, v in params.items()}
        as_dict = lambda c: c if isinstance(c, dict) else {k: c for k in params}
        self.eps = as_dict(stepsize)
        self.alpha = as_dict(alpha)
        self.a = as_dict(a)
        self.opt_stepsize = opt_stepsize
        self.grad_fn = grad_fn or post.grad
        self.full_grad_fn = full_grad_fn or post.full_grad
        self.mom = {k: np.zeros_like(v) for k, v in self.theta.items()}
        # one scalar thermostat per tensor, initialised at a (R/dynamics.R:85)
        self.thermo = {k: dt(self.a[k]) for k in self.theta}
        self.theta_hat = None
        self.g_hat = None

    # ---- initialisation -------------------------------------------------
    def init_momentum(self, noise: np.ndarray):
        """nu0 / u0 = sqrt(stepsize) * N(0,1) (R/dynamics.R:47,84)."""
        z = split_noise(noise.astype(self.dtype), self.theta)
        for k in self.theta:
            self.mom[k] = self.dtype(math.sqrt(self.eps[k])) * z[k]

    def find_mode(self, idx_per_iter):
        """getMode: plain SGD ascent on theta_hat, all tensors at once (cV.R:30-52,87-88),
        then calcFullGrads and theta := theta_hat (cV.R:105-111)."""
        th = {k: v.copy() for k, v in self.theta.items()}
        lr = self.dtype(self.opt_stepsize)
        for idx in idx_per_iter:
            g = self.grad_fn(th, take(self.data, idx))
            th = {k: (th[k] + lr * g[k]).astype(self.dtype) for k in th}
        self.theta_hat = th
        self.g_hat = self.full_grad_fn(th, self.data)
        self.theta = {k: v.copy() for k, v in th.items()}

    # ---- gradient estimator --------------------------------------------
    def grad_estimate(self, batch) -> Params:
        g = self.grad_fn(self.theta, batch)
        if not self.cv:
            return g
        gh = self.grad_fn(self.theta_hat, batch)
        # optGradFull - optGradCurr + gradCurr  (R/setup.R:58)
        return {k: ((self.g_hat[k] - gh[k]) + g[k]).astype(self.dtype) for k in g}

    # ---- one sgmcmcStep ---------------------------------------------------
    def step(self, idx, noise):
        """idx: [n] (sgld, sgnht) or [L, n] (sghmc); noise: flat standard normals,
        [P] for sgld/sgnht, [(L+1) P] for sghmc (refresh block first)."""
        dt = self.dtype
        idx = np.asarray(idx)
        noise = np.asarray(noise, dtype=dt)
        P = sum(int(np.size(v)) for v in self.theta.values())
        if self.algo == "sgld":
            z = split_noise(noise[:P], self.theta)
            batch = take(self.data, idx.reshape(-1))
            for k in self.theta:  # R/update.R:43-45, one run per tensor
                g = self.grad_estimate(batch)[k]
                e = self.eps[k]
                inj = dt(math.sqrt(0.5)) * (dt(math.sqrt(2 * e)) * z[k])
                self.theta[k] = (self.theta[k] + (dt(0.5 * e) * g + inj)).astype(dt)
        elif self.algo == "sghmc":
            idx = idx.reshape(self.L, -1)
            z = split_noise(noise[:P], self.theta)
            for k in self.theta:  # refresh (R/dynamics.R:62), last of the L executions
                self.mom[k] = (dt(0.5) * (dt(math.sqrt(2 * self.eps[k])) * z[k])).astype(dt)
            for l in range(self.L):  # R/update.R:59-65
                batch = take(self.data, idx[l])
                z = split_noise(noise[(l + 1) * P:(l + 2) * P], self.theta)
                for k in self.theta:
                    g = self.grad_estimate(batch)[k]
                    e, al = self.eps[k], self.alpha[k]
                    inj = dt(math.sqrt(al)) * (dt(math.sqrt(2 * e)) * z[k])
                    upd = dt(e) * g - dt(al) * self.mom[k] + inj
                    self.mom[k] = (self.mom[k] + upd).astype(dt)
                    self.theta[k] = (self.theta[k] + self.mom[k]).astype(dt)
        elif self.algo == "sgnht":
            z = split_noise(noise[:P], self.theta)
            batch = take(self.data, idx.reshape(-1))
            g = self.grad_estimate(batch)  # theta is untouched during the u pass
            for k in self.theta:  # R/update.R:72-74
                e, a = self.eps[k], self.a[k]
                inj = dt(math.sqrt(a)) * (dt(math.sqrt(2 * e)) * z[k])
                self.mom[k] = (self.mom[k] + (dt(e) * g[k] - self.mom[k] * self.thermo[k] + inj)).astype(dt)
            for k in self.theta:  # R/update.R:75-77
                self.theta[k] = (self.theta[k] + self.mom[k]).astype(dt)
            for k in self.theta:  # R/update.R:78-80, R/dynamics.R:108-116
                u = self.mom[k]
                self.thermo[k] = dt(self.thermo[k] + ((u * u).sum(dtype=dt) / dt(np.size(u)) - dt(self.eps[k])))
        else:
            raise ValueError(self.algo)

    def noise_len(self) -> int:
        P = sum(int(np.size(v)) for v in self.theta.values())
        return (self.L + 1) * P if self.algo == "sghmc" else P

    def n_sub(self) -> int:
        return self.L if self.algo == "sghmc" else 1
