"""Closed-form restatement of the four vignette model families (numpy, any dtype).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each class restates the ``logLik(params, dataset)`` / ``logPrior(params)`` closures
the reference's vignettes and tests define with TensorFlow / TFP expressions, and
their gradients in closed form (the closed forms were checked
against autograd of the op-by-op replica in ``oracle/tfgraph.py``, see
tests/test_oracle.py).  Parameters are dicts of arrays in the reference's
``names(params)`` order (python dicts preserve insertion order); datasets are dicts
sliced on their first axis (R/update.R:180-194).

TFP parameterisation notes: ``Normal(loc, scale)`` takes a standard deviation;
``MultivariateNormalDiag(loc, scale_diag)`` likewise; ``Categorical(c(.5,.5))``'s
first positional argument is logits; ``Gamma(concentration, rate)``.
"""
from __future__ import annotations

import numpy as np

LOG_2PI = float(np.log(2.0 * np.pi))


def _f(x, dtype):
    return np.asarray(x, dtype=dtype)


class GaussMean:
    """Gaussian with unknown mean, known covariance.

    vignettes/mvGauss.Rmd:37-54 (MultivariateNormalDiag(theta, c(1,1)), Normal(0,10)
    prior); tests/testthat/test1dGauss.R:19-27 (scalar, prior sd 5);
    tests/testthat/test3dGauss.R:20-30 (full covariance); testOptional.R:24-27 (flat
    prior).  ``precision`` is Sigma^{-1}; ``prior_sd=None`` is the flat prior of
    R/setup.R:119-123.
    """

    names = ("theta",)
    data_names = ("X",)

    def __init__(self, d, precision=None, prior_sd=None):
        self.d = int(d)
        self.precision = np.eye(self.d) if precision is None else np.asarray(precision, dtype=np.float64)
        self.prior_sd = prior_sd
        sign, logdet = np.linalg.slogdet(self.precision)
        assert sign > 0
        self.logdet_prec = float(logdet)

    def loglik(self, params, batch, dtype=np.float64):
        th = _f(params["theta"], dtype).reshape(-1)
        X = _f(batch["X"], dtype).reshape(-1, self.d)
        P = _f(self.precision, dtype)
        r = X - th
        q = np.einsum("ij,jk,ik->i", r, P, r)
        n = X.shape[0]
        return dtype(-0.5) * q.sum(dtype=dtype) + dtype(n * (0.5 * self.logdet_prec - 0.5 * self.d * LOG_2PI))

    def logprior(self, params, dtype=np.float64):
        if self.prior_sd is None:
            return dtype(0.0)
        th = _f(params["theta"], dtype).reshape(-1)
        sd = dtype(self.prior_sd)
        return (dtype(-0.5) * (th / sd) ** 2 - np.log(sd) - dtype(0.5 * LOG_2PI)).sum(dtype=dtype)

    def grad_loglik(self, params, batch, dtype=np.float64):
        th = _f(params["theta"], dtype).reshape(-1)
        X = _f(batch["X"], dtype).reshape(-1, self.d)
        P = _f(self.precision, dtype)
        s = (X - th).sum(axis=0, dtype=dtype)
        return {"theta": (P @ s).reshape(np.shape(params["theta"]))}

    def grad_logprior(self, params, dtype=np.float64):
        th = _f(params["theta"], dtype)
        if self.prior_sd is None:
            return {"theta": np.zeros_like(th)}
        return {"theta": -th / dtype(self.prior_sd) ** 2}


class Mixture2:
    """Two-component, equal-weight, unit-variance Gaussian mixture.

    vignettes/gaussMixture.Rmd:48-76: Mixture(Categorical(logits=c(.5,.5)),
    MultivariateNormalDiag(theta_k, c(1,1))); prior MultivariateNormalDiag(0, c(10,10))
    on each component mean.
    """

    names = ("theta1", "theta2")
    data_names = ("X",)

    def __init__(self, d=2, prior_sd=10.0):
        self.d = int(d)
        self.prior_sd = prior_sd

    def _logcomp(self, params, batch, dtype):
        X = _f(batch["X"], dtype).reshape(-1, self.d)
        out = []
        for k in self.names:
            r = X - _f(params[k], dtype).reshape(-1)
            out.append(dtype(-0.5) * (r * r).sum(axis=1, dtype=dtype) - dtype(0.5 * self.d * LOG_2PI)
                       + dtype(np.log(0.5)))
        return X, np.stack(out, axis=1)

    def loglik(self, params, batch, dtype=np.float64):
        _, lc = self._logcomp(params, batch, dtype)
        m = lc.max(axis=1)
        return (m + np.log(np.exp(lc - m[:, None]).sum(axis=1, dtype=dtype))).sum(dtype=dtype)

    def logprior(self, params, dtype=np.float64):
        if self.prior_sd is None:
            return dtype(0.0)
        sd = dtype(self.prior_sd)
        tot = dtype(0.0)
        for k in self.names:
            th = _f(params[k], dtype).reshape(-1)
            tot = tot + (dtype(-0.5) * (th / sd) ** 2 - np.log(sd) - dtype(0.5 * LOG_2PI)).sum(dtype=dtype)
        return tot

    def grad_loglik(self, params, batch, dtype=np.float64):
        X, lc = self._logcomp(params, batch, dtype)
        m = lc.max(axis=1, keepdims=True)
        w = np.exp(lc - m)
        resp = w / w.sum(axis=1, keepdims=True, dtype=dtype)
        out = {}
        for j, k in enumerate(self.names):
            th = _f(params[k], dtype).reshape(-1)
            out[k] = (resp[:, j:j + 1] * (X - th)).sum(axis=0, dtype=dtype).reshape(np.shape(params[k]))
        return out

    def grad_logprior(self, params, dtype=np.float64):
        out = {}
        for k in self.names:
            th = _f(params[k], dtype)
            out[k] = np.zeros_like(th) if self.prior_sd is None else -th / dtype(self.prior_sd) ** 2
        return out


class Logistic:
    """Logistic regression with Laplace(0,1) prior.

    vignettes/logisticRegression.Rmd:45-57.  ``bias`` is rank 0, ``beta`` is [d,1].
    The gradient uses the closed form y - sigmoid(z); the reference's naive
    ``log(1 - 1/(1+exp(-z)))`` overflows for |z| >~ 17 in float32 (SURVEY.md
    Appendix C) -- parity is only asserted on unsaturated inputs.
    """

    names = ("bias", "beta")
    data_names = ("X", "y")

    def __init__(self, d, laplace_prior=True):
        self.d = int(d)
        self.laplace_prior = laplace_prior

    def _z(self, params, batch, dtype):
        X = _f(batch["X"], dtype)
        beta = _f(params["beta"], dtype).reshape(-1)
        return X, X @ beta + _f(params["bias"], dtype).reshape(())

    def loglik(self, params, batch, dtype=np.float64):
        _, z = self._z(params, batch, dtype)
        y = _f(batch["y"], dtype).reshape(-1)
        # y*log(s) + (1-y)*log(1-s) in the overflow-free form
        return (y * z - np.logaddexp(dtype(0.0), z)).sum(dtype=dtype)

    def logprior(self, params, dtype=np.float64):
        if not self.laplace_prior:
            return dtype(0.0)
        return -(np.abs(_f(params["beta"], dtype)).sum(dtype=dtype) + np.abs(_f(params["bias"], dtype)).sum(dtype=dtype))

    def grad_loglik(self, params, batch, dtype=np.float64):
        X, z = self._z(params, batch, dtype)
        y = _f(batch["y"], dtype).reshape(-1)
        r = y - dtype(1.0) / (dtype(1.0) + np.exp(-z))
        return {"bias": r.sum(dtype=dtype).reshape(np.shape(params["bias"])),
                "beta": (X.T @ r).reshape(np.shape(params["beta"]))}

    def grad_logprior(self, params, dtype=np.float64):
        if not self.laplace_prior:
            return {k: np.zeros_like(_f(params[k], dtype)) for k in self.names}
        return {k: -np.sign(_f(params[k], dtype)) for k in self.names}


class SoftmaxNet:
    """784-100-10 softmax-softmax network with Gamma-precision Normal priors.

    vignettes/nn.Rmd:58-79: H = softmax(X B + b), Pr = softmax(H A + a),
    logLik = sum(y * log Pr); every element of W in {A,B,a,b} ~ Normal(0,
    rsqrt(lambda_W)), lambda_W ~ Gamma(1, 1).  Parameter order is the vignette's
    (nn.Rmd:43-53): A, B, a, b, lambdaA, lambdaB, lambdaa, lambdab.
    """

    names = ("A", "B", "a", "b", "lambdaA", "lambdaB", "lambdaa", "lambdab")
    data_names = ("X", "y")
    _lam = {"A": "lambdaA", "B": "lambdaB", "a": "lambdaa", "b": "lambdab"}

    def __init__(self, d_in=784, d_hidden=100, d_out=10):
        self.d_in, self.d_hidden, self.d_out = d_in, d_hidden, d_out

    @staticmethod
    def _softmax(z):
        z = z - z.max(axis=1, keepdims=True)
        e = np.exp(z)
        return e / e.sum(axis=1, keepdims=True)

    def _forward(self, params, batch, dtype):
        X = _f(batch["X"], dtype)
        H = self._softmax(X @ _f(params["B"], dtype) + _f(params["b"], dtype).reshape(1, -1))
        Pr = self._softmax(H @ _f(params["A"], dtype) + _f(params["a"], dtype).reshape(1, -1))
        return X, H, Pr

    def loglik(self, params, batch, dtype=np.float64):
        _, _, Pr = self._forward(params, batch, dtype)
        return (_f(batch["y"], dtype) * np.log(Pr)).sum(dtype=dtype)

    def logprior(self, params, dtype=np.float64):
        tot = dtype(0.0)
        for w, l in self._lam.items():
            W = _f(params[w], dtype)
            lam = _f(params[l], dtype).reshape(())
            tot = tot + dtype(W.size) * (dtype(0.5) * np.log(lam) - dtype(0.5 * LOG_2PI)) \
                - dtype(0.5) * lam * (W * W).sum(dtype=dtype) - lam
        return tot

    def grad_loglik(self, params, batch, dtype=np.float64):
        X, H, Pr = self._forward(params, batch, dtype)
        Y = _f(batch["y"], dtype)
        A = _f(params["A"], dtype)
        dZ2 = Y - Pr * Y.sum(axis=1, keepdims=True)
        dH = dZ2 @ A.T
        dZ1 = H * (dH - (dH * H).sum(axis=1, keepdims=True))
        out = {"A": H.T @ dZ2, "B": X.T @ dZ1,
               "a": dZ2.sum(axis=0).reshape(np.shape(params["a"])),
               "b": dZ1.sum(axis=0).reshape(np.shape(params["b"]))}
        for l in self._lam.values():
            out[l] = np.zeros(np.shape(params[l]), dtype=dtype)
        return out

    def grad_logprior(self, params, dtype=np.float64):
        out = {}
        for w, l in self._lam.items():
            W = _f(params[w], dtype)
            lam = _f(params[l], dtype).reshape(())
            out[w] = -lam * W
            out[l] = (dtype(W.size) / (dtype(2.0) * lam) - dtype(0.5) * (W * W).sum(dtype=dtype)
                      - dtype(1.0)).reshape(np.shape(params[l]))
        return out
