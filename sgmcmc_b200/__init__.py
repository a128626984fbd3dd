"""sgmcmc_b200 -- B200-native (sm_100a) implementation of the STOR-i/sgmcmc sampler hot path.

Only the hot path is here: the C-ABI library (csrc/, include/sgmcmc_b200.h) and the
host-side mirror of the reference's R API for that path (api.py).  Importing the API
needs the built library; there is no CPU fallback.
"""
from .api import (getChains, getParams, heldoutLogLik, initSess, loadState, saveState, sghmc, sghmccv, sghmccvSetup, sghmcSetup, sgld,  # noqa: F401
                  sgldcv, sgldcvSetup, sgldSetup, sgmcmcStep, sgnht, sgnhtcv, sgnhtcvSetup, sgnhtSetup)
from .models import (gammaPrecisionPrior, gaussMeanLogLik, laplacePrior, logisticLogLik,  # noqa: F401
                     mixture2LogLik, normalPrior, softmaxNetLogLik)
from ._lib import SgmcmcError  # noqa: F401
