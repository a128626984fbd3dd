"""Model-spec objects that stand where the reference takes ``logLik`` / ``logPrior`` closures.

The reference differentiates arbitrary TensorFlow closures (R/setup.R:36-62); on the
device that is impossible, so the four vignette families are tagged objects routed to
hand-written gradient kernels.  Anything else -- including an
ordinary Python callable -- is rejected with a clear error; there is no fallback.
"""
from __future__ import annotations

import numpy as np

from . import _lib


class LogLik:
    model_id = 0
    #: dataset entry names in the order the C ABI expects them
    data_names: tuple = ()


class LogPrior:
    pass


class gaussMeanLogLik(LogLik):
    """Gaussian with unknown mean ``params$theta`` and known covariance ``Sigma`` (default I).

    vignettes/mvGauss.Rmd:37-46; tests/testthat/test1dGauss.R:19-22 (scalar);
    tests/testthat/test3dGauss.R:20-24 (full covariance)."""
    model_id = _lib.MODEL_GAUSS_MEAN
    data_names = ("X",)

    def __init__(self, Sigma=None):
        self.Sigma = None if Sigma is None else np.atleast_2d(np.asarray(Sigma, dtype=np.float64))

    def precision(self):
        return None if self.Sigma is None else np.linalg.inv(self.Sigma)


class mixture2LogLik(LogLik):
    """0.5 N(theta1, I) + 0.5 N(theta2, I) (vignettes/gaussMixture.Rmd:48-61)."""
    model_id = _lib.MODEL_MIXTURE2
    data_names = ("X",)


class logisticLogLik(LogLik):
    """Bernoulli(sigmoid(bias + X beta)) (vignettes/logisticRegression.Rmd:45-49)."""
    model_id = _lib.MODEL_LOGISTIC
    data_names = ("X", "y")


class softmaxNetLogLik(LogLik):
    """Categorical(softmax(softmax(X B + b) A + a)) (vignettes/nn.Rmd:58-65)."""
    model_id = _lib.MODEL_SOFTMAX_NET
    data_names = ("X", "y")


class normalPrior(LogPrior):
    """Independent Normal(0, sd) on every mean entry (mvGauss.Rmd:49-54,
    gaussMixture.Rmd:66-76, test1dGauss.R:24-27).  ``sd`` is a standard deviation."""

    def __init__(self, sd):
        self.sd = float(sd)


class laplacePrior(LogPrior):
    """Laplace(0, 1) on bias and beta (vignettes/logisticRegression.Rmd:53-57)."""


class gammaPrecisionPrior(LogPrior):
    """W ~ Normal(0, lambda_W^-1/2), lambda_W ~ Gamma(1, 1) (vignettes/nn.Rmd:67-79)."""


_PRIORS = {
    _lib.MODEL_GAUSS_MEAN: normalPrior,
    _lib.MODEL_MIXTURE2: normalPrior,
    _lib.MODEL_LOGISTIC: laplacePrior,
    _lib.MODEL_SOFTMAX_NET: gammaPrecisionPrior,
}

NOT_ROUTED = ("logLik/logPrior must be model-spec objects from sgmcmc_b200.models "
              "(gaussMeanLogLik, mixture2LogLik, logisticLogLik, softmaxNetLogLik and their priors): "
              "arbitrary closures cannot be differentiated on the device and there is no CPU fallback")


def resolve(logLik, logPrior):
    """Validate the (logLik, logPrior) pair; returns (model_id, use_prior, prior_sd, precision)."""
    if not isinstance(logLik, LogLik):
        raise _lib.SgmcmcError(_lib.SGM_ENOMODEL, NOT_ROUTED)
    if logPrior is not None and not isinstance(logPrior, LogPrior):
        raise _lib.SgmcmcError(_lib.SGM_ENOMODEL, NOT_ROUTED)
    if logPrior is not None and not isinstance(logPrior, _PRIORS[logLik.model_id]):
        raise _lib.SgmcmcError(_lib.SGM_ENOMODEL,
                               f"{type(logLik).__name__} is routed with {_PRIORS[logLik.model_id].__name__} "
                               f"(or logPrior=None), not {type(logPrior).__name__}")
    prior_sd = logPrior.sd if isinstance(logPrior, normalPrior) else 0.0
    precision = logLik.precision() if isinstance(logLik, gaussMeanLogLik) else None
    return logLik.model_id, int(logPrior is not None), prior_sd, precision
