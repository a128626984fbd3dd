"""Reference-structured CPU baseline ("TF-proxy (torch CPU) -- not the reference binary").

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's R + TensorFlow path cannot run in this image (no R, no TensorFlow,
no network), so the CPU arm of bench.py times this proxy: the
float32 op-by-op graph replica of oracle/tfgraph.py executed with the reference's
structure --
  * host-side ``sample(N, n)`` without replacement and a gather from column-major
    double matrices converted to float32 on every feed (R/update.R:166-194);
  * ONE gradient evaluation per parameter tensor per dynamics op, each standing for
    one ``Session.run`` (R/update.R:43-45, 60-64, 72-80);
  * a per-iteration copy of every parameter into a host double array
    (R/storage.R:54-80);
  * a fresh-minibatch log-posterior evaluation every 100 iterations
    (R/update.R:157-160).
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from . import samplers as OS
from . import tfgraph as TG


class ReferenceStructuredSampler:
    def __init__(self, model, data, params, algo, stepsize, cv=False, alpha=0.01, a=0.01, L=5,
                 opt_stepsize=0.0, n: int = 100, use_prior=True, seed=13):
        self.rs = np.random.RandomState(seed)
        self.data = {k: np.asfortranarray(np.asarray(v, dtype=np.float64)) for k, v in data.items()}
        self.N = next(iter(self.data.values())).shape[0]
        self.n = int(n)
        self.gp = TG.GraphPosterior(model, self.N, self.n, use_prior)
        self.algo, self.cv, self.L = algo, cv, int(L)
        self.names = list(params)
        self.theta = {k: torch.tensor(np.asarray(v, np.float32)) for k, v in params.items()}
        d = lambda c: c if isinstance(c, dict) else {k: c for k in params}
        self.eps, self.alpha, self.a = d(stepsize), d(alpha), d(a)
        self.opt_stepsize = opt_stepsize
        self.mom = {k: math.sqrt(self.eps[k]) * torch.randn_like(v) for k, v in self.theta.items()}
        self.zeta = {k: torch.tensor(float(self.a[k])) for k in params}
        self.theta_hat = None
        self.g_hat = None

    # dataFeed (R/update.R:166-177).  Costed like R, not like a naive numpy transcription: sample(N, n) without replacement is O(n)
    # random draws there (numpy's choice(replace=False) permutes all N rows: 5 ms at N = 581012, more than the rest of the step), and
    # `[` on a column-major matrix walks column by column (numpy's row gather of an F-ordered array is strided in the wrong direction).
    def sample_rows(self):
        n, N = self.n, self.N
        if 2 * n > N:
            return self.rs.permutation(N)[:n]
        sel = np.unique(self.rs.randint(0, N, size=n + n // 8 + 8))
        while sel.size < n:
            sel = np.unique(np.concatenate([sel, self.rs.randint(0, N, size=n)]))
        self.rs.shuffle(sel)
        return sel[:n]

    def feed(self):
        sel = self.sample_rows()
        out = {}
        for k, v in self.data.items():
            g = v[sel] if v.ndim < 2 else np.take(v.T, sel, axis=-1).T    # v.T is C-contiguous: one contiguous column per output column
            out[k] = torch.from_numpy(np.ascontiguousarray(g, dtype=np.float32))
        return out

    def _grad_one(self, theta, batch, name, corrected=True):
        p = {k: v.clone().requires_grad_(k == name) for k, v in theta.items()}
        val = self.gp._post(p, batch, corrected)
        (g,) = torch.autograd.grad(val, [p[name]], allow_unused=True)
        return torch.zeros_like(p[name]) if g is None else g

    def grad(self, batch, name):
        g = self._grad_one(self.theta, batch, name)
        if self.cv:  # R/setup.R:58
            g = self.g_hat[name] - self._grad_one(self.theta_hat, batch, name) + g
        return g

    def find_mode(self, n_iters_opt):
        th = {k: v.clone() for k, v in self.theta.items()}
        for i in range(n_iters_opt):
            b = self.feed()
            p = {k: v.clone().requires_grad_(True) for k, v in th.items()}
            val = self.gp._post(p, b, True)
            gs = torch.autograd.grad(val, list(p.values()), allow_unused=True)
            th = {k: v + self.opt_stepsize * (g if g is not None else 0) for (k, v), g in zip(th.items(), gs)}
        self.theta_hat = th
        full = {k: torch.from_numpy(v.astype(np.float32)) for k, v in self.data.items()}
        self.g_hat = {k: self._grad_one(th, full, k, corrected=False) for k in self.names}  # one run per tensor
        self.theta = {k: v.clone() for k, v in th.items()}

    def step(self):
        if self.algo == "sgld":
            b = self.feed()
            for k in self.names:
                e = self.eps[k]
                g = self.grad(b, k)
                self.theta[k] = self.theta[k] + 0.5 * e * g + math.sqrt(0.5) * math.sqrt(2 * e) * torch.randn_like(g)
        elif self.algo == "sghmc":
            feeds = []
            for _ in range(self.L):
                feeds.append(self.feed())
                for k in self.names:
                    self.mom[k] = 0.5 * math.sqrt(2 * self.eps[k]) * torch.randn_like(self.theta[k])
            for l in range(self.L):
                for k in self.names:
                    e, al = self.eps[k], self.alpha[k]
                    g = self.grad(feeds[l], k)
                    self.mom[k] = self.mom[k] + e * g - al * self.mom[k] + math.sqrt(al) * math.sqrt(2 * e) * torch.randn_like(g)
                    self.theta[k] = self.theta[k] + self.mom[k]
        else:
            b = self.feed()
            for k in self.names:
                e, a = self.eps[k], self.a[k]
                g = self.grad(b, k)
                self.mom[k] = self.mom[k] + e * g - self.mom[k] * self.zeta[k] + math.sqrt(a) * math.sqrt(2 * e) * torch.randn_like(g)
            for k in self.names:
                self.theta[k] = self.theta[k] + self.mom[k]
            for k in self.names:
                u = self.mom[k]
                self.zeta[k] = self.zeta[k] + (u * u).sum() / u.numel() - self.eps[k]

    def run(self, n_iters, storage=None):
        """runSGMCMC's loop incl. storeState and the divergence check (R/update.R:154-161)."""
        if storage is None:
            storage = {k: np.zeros((n_iters,) + tuple(v.shape)) for k, v in self.theta.items()}
        for i in range(n_iters):
            self.step()
            for k in self.names:
                storage[k][i] = self.theta[k].numpy().astype(np.float64)
            if (i + 1) % 100 == 0:
                with torch.no_grad():
                    v = float(self.gp._post(self.theta, self.feed(), True))
                if math.isnan(v):
                    raise RuntimeError("Chain diverged, try decreasing stepsize")
        return storage


def time_iterations(sampler: ReferenceStructuredSampler, budget_s: float, probe: int = 20):
    """Iterations/s over a bounded sample sized to ~budget_s of CPU work."""
    t0 = time.perf_counter()
    sampler.run(probe)
    per = (time.perf_counter() - t0) / probe
    iters = int(max(probe, min(20000, budget_s / max(per, 1e-9))))
    t0 = time.perf_counter()
    sampler.run(iters)
    dt = time.perf_counter() - t0
    return iters / dt, iters, dt


def time_full_gradient(model, data, params, use_prior=True, repeats=3):
    """calcFullGrads (R/controlVariates.R:105-111): one full-data gradient run per tensor; returns seconds."""
    N = next(iter(data.values())).shape[0]
    gp = TG.GraphPosterior(model, N, N, use_prior)
    th = {k: torch.tensor(np.asarray(v, np.float32)) for k, v in params.items()}
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        full = {k: torch.from_numpy(np.asarray(v).astype(np.float32)) for k, v in data.items()}  # feed conversion
        for name in params:
            p = {k: v.clone().requires_grad_(k == name) for k, v in th.items()}
            val = gp._post(p, full, False)
            torch.autograd.grad(val, [p[name]], allow_unused=True)
        best = min(best, time.perf_counter() - t0)
    return best
