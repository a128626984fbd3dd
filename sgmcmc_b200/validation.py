"""Deterministic validation mode and state introspection (not part of the reference's API).

The reference's RNG streams (R's ``sample`` + TensorFlow's stateful ``random_normal``)
cannot be reproduced outside R/TF, so parity is checked by
feeding the SAME minibatch rows and standard-normal noise to the device kernels and
to the CPU oracle.  These helpers expose the corresponding C-ABI entry points.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import Session, Sgmcmc, _split_flat

_PF = C.POINTER(C.c_float)
_PI = C.POINTER(C.c_int32)
_PD = C.POINTER(C.c_double)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def initSessDeterministic(sgmcmc: Sgmcmc, momentumNoise=None, optIdx=None) -> Session:
    """initSess with caller-supplied momentum normals [nChains, P] and, for *cv, optimiser
    minibatch rows [nChains, nItersOpt, n] (0-based)."""
    mn, oi = _f32(momentumNoise), _i32(optIdx)
    _lib.check(_lib.load().sgm_init_deterministic(sgmcmc.handle, _p(mn, _PF), _p(oi, _PI)))
    return Session(sgmcmc)


def stepDeterministic(sgmcmc: Sgmcmc, idx, noise):
    """One sgmcmcStep on rows idx [nChains, nSub, n] with normals noise [nChains, noiseLen]."""
    ii, nn = _i32(idx), _f32(noise)
    _lib.check(_lib.load().sgm_step_deterministic(sgmcmc.handle, _p(ii, _PI), _p(nn, _PF)))


def noiseLen(sgmcmc: Sgmcmc) -> int:
    return int(_lib.load().sgm_noise_len(sgmcmc.handle))


def gradEstimate(sgmcmc: Sgmcmc, idx) -> np.ndarray:
    """Gradient estimate at the current state on rows idx [nChains, n]: flat [nChains, P]."""
    ii = _i32(idx)
    out = np.empty((sgmcmc.nChains, sgmcmc.P), dtype=np.float64)
    _lib.check(_lib.load().sgm_grad_estimate(sgmcmc.handle, _p(ii, _PI), _p(out, _PD)))
    return out


def flatParams(sgmcmc: Sgmcmc) -> np.ndarray:
    out = np.empty((sgmcmc.nChains, sgmcmc.P), dtype=np.float64)
    _lib.check(_lib.load().sgm_get_params(sgmcmc.handle, _p(out, _PD)))
    return out


def getState(sgmcmc: Sgmcmc):
    """(momentum [nChains, P], thermostat [nChains, nTensors])."""
    mom = np.empty((sgmcmc.nChains, sgmcmc.P), dtype=np.float64)
    th = np.empty((sgmcmc.nChains, len(sgmcmc.paramShapes)), dtype=np.float64)
    _lib.check(_lib.load().sgm_get_state(sgmcmc.handle, _p(mom, _PD), _p(th, _PD)))
    return mom, th


def getMode(sgmcmc: Sgmcmc):
    """*cv: (theta_hat, full-data gradient at it), each flat [nChains, P]."""
    th = np.empty((sgmcmc.nChains, sgmcmc.P), dtype=np.float64)
    g = np.empty((sgmcmc.nChains, sgmcmc.P), dtype=np.float64)
    _lib.check(_lib.load().sgm_get_mode(sgmcmc.handle, _p(th, _PD), _p(g, _PD)))
    return th, g


def estLogPostDeterministic(sgmcmc: Sgmcmc, idx) -> np.ndarray:
    ii = _i32(idx)
    out = np.empty(sgmcmc.nChains, dtype=np.float64)
    _lib.check(_lib.load().sgm_est_logpost_deterministic(sgmcmc.handle, _p(ii, _PI), _p(out, _PD)))
    return out


def copyDataset(sgmcmc: Sgmcmc, entry: int, row0: int, rows: int, cols: int) -> np.ndarray:
    out = np.empty((cols, rows), dtype=np.float64)  # column-major (rows, cols)
    _lib.check(_lib.load().sgm_copy_dataset(sgmcmc.handle, entry, row0, rows, _p(out, _PD)))
    return out.T.copy()


def lastTiming(sgmcmc: Sgmcmc):
    ms, n = C.c_float(), C.c_int32()
    _lib.check(_lib.load().sgm_last_timing(sgmcmc.handle, C.byref(ms), C.byref(n)))
    return ms.value, n.value


def lastFullGradTiming(sgmcmc: Sgmcmc):
    ms, b = C.c_float(), C.c_int64()
    _lib.check(_lib.load().sgm_last_fullgrad_timing(sgmcmc.handle, C.byref(ms), C.byref(b)))
    return ms.value, b.value


def lastFullGradFlops(sgmcmc: Sgmcmc) -> float:
    """4 N d M of the last full-data gradient if it ran as tensor-core contractions (one mode per chain), else 0."""
    f = C.c_double()
    _lib.check(_lib.load().sgm_last_fullgrad_flops(sgmcmc.handle, C.byref(f)))
    return f.value


def benchFullGradient(sgmcmc: Sgmcmc, reps: int = 20):
    """(ms per launch, bytes read) of the control-variate full-data gradient kernel, `reps` launches back to back."""
    ms, b = C.c_float(), C.c_int64()
    _lib.check(_lib.load().sgm_bench_full_gradient(sgmcmc.handle, reps, C.byref(ms), C.byref(b)))
    return ms.value, b.value


def setOption(sgmcmc: Sgmcmc, key: str, value: int):
    """Pins a kernel path of this sampler (include/sgmcmc_b200.h: sgm_set_option); 0 restores the default dispatch."""
    _lib.check(_lib.load().sgm_set_option(sgmcmc.handle, key.encode(), int(value)))


def lastCollectiveTiming(sgmcmc: Sgmcmc):
    """(reduction ms, all-gather ms) of the last control-variate full-data pass's collectives, timed on the device."""
    a, b = C.c_float(), C.c_float()
    _lib.check(_lib.load().sgm_last_collective_timing(sgmcmc.handle, C.byref(a), C.byref(b)))
    return a.value, b.value
