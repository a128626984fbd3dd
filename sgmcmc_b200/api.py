"""Host-side mirror of the reference's exported R API, above the C ABI.

R is absent from this image, so the host layer the north star asks for in R
(r-package/R/*.R, unverified here) is mirrored in Python with the same function
names, argument names / order / defaults and error behaviour (NAMESPACE:3-24):

    sgld, sghmc, sgnht, sgldcv, sghmccv, sgnhtcv,
    sgldSetup, sghmcSetup, sgnhtSetup, sgldcvSetup, sghmccvSetup, sgnhtcvSetup,
    initSess, sgmcmcStep, getParams

``logLik`` / ``logPrior`` take model-spec objects (sgmcmc_b200.models) instead of
TensorFlow closures; ``dataset`` / ``params`` / ``stepsize`` are dicts (R named lists).
Extra keyword arguments (``nChains``, ``device``, ``sharedMode``) default to the
reference's behaviour (one chain).  Everything numeric happens in
libsgmcmc_b200.so; this file only marshals.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Dict

import numpy as np

from . import _lib, models
from ._lib import SgmcmcError


def _message(text: str):
    print(text, file=sys.stderr)


def _as_list(const, params):
    """convertList (R/setup.R:160-170)."""
    if isinstance(const, dict):
        return dict(const)
    return {k: const for k in params}


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Sgmcmc:
    """The reference's sgmcmc S3 list (fields: R/sgld.R:71-87, R/sgldcv.R:68-92)."""

    def __init__(self):
        self.classes = ()
        self.handle = None
        self._keep = []

    def inherits(self, cls):
        return cls in self.classes

    def __del__(self):
        try:
            if self.handle is not None:
                _lib.load().sgm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Session:
    """What initSess returns: the device state behind a handle (replaces tf$Session)."""

    def __init__(self, sgmcmc: Sgmcmc):
        self.sgmcmc = sgmcmc

    @property
    def handle(self):
        return self.sgmcmc.handle


def _host_matrix(x):
    """A dataset entry as the library takes it: float64, contiguous in EITHER order (sgm_array.row_major says which; the packing into
    device rows happens in a kernel).  A float64 numpy array in C or Fortran order is passed as it is -- transposing 581,012 x 54
    doubles on the host (np.asfortranarray) took 0.42 s, more than the rest of a c3 job; anything else is copied once, in C order."""
    v = np.asarray(x, dtype=np.float64)
    if v.ndim == 2 and v.flags.f_contiguous and not v.flags.c_contiguous:
        return v, 0
    return np.ascontiguousarray(v), 1


def _create(logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, algo, cv, *,
            optStepsize=0.0, nItersOpt=0, alpha=0.01, L=5, a=0.01, nChains=1, device=0, sharedMode=False,
            synth=None, chainOffset=0) -> Sgmcmc:
    """createSGMCMC / createSGMCMCCV (R/setup.R:2-28, R/controlVariates.R:2-25)."""
    lib = _lib.load()
    model_id, use_prior, prior_sd, precision = models.resolve(logLik, logPrior)
    obj = Sgmcmc()
    keep = obj._keep
    params = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    stepsize = _as_list(stepsize, params)
    alpha = _as_list(alpha, params)
    a = _as_list(a, params)

    names = list(logLik.data_names)
    arrs = (_lib.SgmArray * len(names))()
    for i, nm in enumerate(names):
        arrs[i].name = nm.encode()
        if synth is None:
            if nm not in dataset:
                raise SgmcmcError(_lib.SGM_EINVAL, f"dataset entry '{nm}' is required by {type(logLik).__name__}")
            v, row_major = _host_matrix(dataset[nm])
            if v.ndim == 0 or v.ndim > 2:
                raise SgmcmcError(_lib.SGM_EINVAL, f"dataset entry '{nm}' must be a vector or a matrix")
            keep.append(v)
            arrs[i].data = _dptr(v)
            arrs[i].ndim = v.ndim
            arrs[i].row_major = row_major
            for j, s in enumerate(v.shape):
                arrs[i].dims[j] = s
    pars = (_lib.SgmParam * len(params))()
    for i, (nm, v) in enumerate(params.items()):
        vf = np.asfortranarray(v).reshape(-1, order="F").copy()
        keep.append(vf)
        pars[i].name = nm.encode()
        pars[i].init = _dptr(vf)
        pars[i].ndim = v.ndim
        for j, s in enumerate(v.shape):
            pars[i].dims[j] = s
        pars[i].stepsize = float(stepsize[nm])
        pars[i].alpha = float(alpha[nm])
        pars[i].a = float(a[nm])
    cfg = _lib.SgmConfig()
    cfg.abi_version = _lib.SGM_ABI_VERSION
    cfg.model, cfg.algo, cfg.cv = model_id, algo, int(cv)
    cfg.n_data, cfg.data = len(names), arrs
    cfg.n_params, cfg.params = len(params), pars
    cfg.minibatch_size = float(minibatchSize)
    cfg.L = int(L)
    cfg.opt_stepsize = float(optStepsize)
    cfg.n_iters_opt = int(nItersOpt)
    cfg.seeded = int(seed is not None)
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF if seed is not None else 0
    cfg.n_chains = int(nChains)
    cfg.chain_offset = int(chainOffset)
    cfg.device = int(device)
    cfg.flags = _lib.FLAG_SHARED_MODE if sharedMode else 0
    cfg.use_prior = use_prior
    cfg.prior_sd = prior_sd
    if precision is not None:
        pf = np.asfortranarray(precision, dtype=np.float64)
        keep.append(pf)
        cfg.precision = _dptr(pf)
    if synth is not None:
        cfg.synth_rows, cfg.synth_dim, cfg.synth_seed = int(synth["rows"]), int(synth["dim"]), int(synth.get("seed", 13))
    h = C.c_void_p()
    _lib.check(lib.sgm_create(C.byref(cfg), C.byref(h)))
    obj.handle = h
    obj._keep = []  # the library copied everything
    obj.N = int(lib.sgm_num_rows(h))
    obj.n = int(lib.sgm_minibatch(h))
    obj.data = dataset
    obj.stepsize = stepsize
    obj.params = params
    obj.paramShapes = {k: v.shape for k, v in params.items()}
    obj.nChains = int(nChains)
    obj.dataNames = tuple(names)
    obj.P = int(lib.sgm_num_scalars(h))
    if algo == _lib.ALGO_SGHMC:
        obj.alpha, obj.L = alpha, int(L)
    if algo == _lib.ALGO_SGNHT:
        obj.a = a
        obj.ranks = {k: v.ndim for k, v in params.items()}
    if cv:
        obj.nItersOpt = int(nItersOpt)
        obj.optStepsize = float(optStepsize)
    return obj


# ---- *Setup (R/sgld.R:114-123, sghmc.R:105-117, sgnht.R:102-115, sgldcv.R:120-129, ...) ----------

def sgldSetup(logLik, dataset, params, stepsize, logPrior=None, minibatchSize=0.01, seed=None, **kw):
    o = _create(logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, _lib.ALGO_SGLD, False, **kw)
    o.classes = ("sgld", "sgmcmc")
    return o


def sghmcSetup(logLik, dataset, params, stepsize, logPrior=None, minibatchSize=0.01, alpha=0.01, L=5, seed=None, **kw):
    o = _create(logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, _lib.ALGO_SGHMC, False,
                alpha=alpha, L=L, **kw)
    o.classes = ("sghmc", "sgmcmc")
    return o


def sgnhtSetup(logLik, dataset, params, stepsize, logPrior=None, minibatchSize=0.01, a=0.01, seed=None, **kw):
    o = _create(logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, _lib.ALGO_SGNHT, False, a=a, **kw)
    o.classes = ("sgnht", "sgmcmc")
    return o


def sgldcvSetup(logLik, dataset, params, stepsize, optStepsize, logPrior=None, minibatchSize=0.01,
                nItersOpt=10 ** 4, verbose=True, seed=None, **kw):
    o = _create(logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, _lib.ALGO_SGLD, True,
                optStepsize=optStepsize, nItersOpt=nItersOpt, **kw)
    o.classes = ("sgld", "sgmcmccv", "sgmcmc")
    return o


def sghmccvSetup(logLik, dataset, params, stepsize, optStepsize, logPrior=None, minibatchSize=0.01,
                 alpha=0.01, L=5, nItersOpt=10 ** 4, verbose=True, seed=None, **kw):
    o = _create(logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, _lib.ALGO_SGHMC, True,
                optStepsize=optStepsize, nItersOpt=nItersOpt, alpha=alpha, L=L, **kw)
    o.classes = ("sghmc", "sgmcmccv", "sgmcmc")
    return o


def sgnhtcvSetup(logLik, dataset, params, stepsize, optStepsize, logPrior=None, minibatchSize=0.01,
                 a=0.01, nItersOpt=10 ** 4, verbose=True, seed=None, **kw):
    o = _create(logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, _lib.ALGO_SGNHT, True,
                optStepsize=optStepsize, nItersOpt=nItersOpt, a=a, **kw)
    o.classes = ("sgnht", "sgmcmccv", "sgmcmc")
    return o


# ---- initSess / sgmcmcStep / getParams (R/update.R:37-140, R/storage.R:37-39) -------------------

def initSess(sgmcmc: Sgmcmc, verbose=True) -> Session:
    lib = _lib.load()
    if sgmcmc.inherits("sgmcmccv"):
        if verbose:
            _message("\nFinding initial MAP estimates...")  # R/controlVariates.R:40

        def report(_user, _phase, it, lp, n):   # checkOptDivergence's line (R/controlVariates.R:121-123), one value per mode
            if verbose:
                v = lp[0] if n == 1 else float(np.mean([lp[i] for i in range(n)]))
                _message(f"Iteration: {it}\t\tLog posterior estimate: {v}")

        cb = _lib.PROGRESS_FN(report)
        _lib.check(lib.sgm_set_progress(sgmcmc.handle, C.cast(cb, C.c_void_p), None))
        try:
            _lib.check(lib.sgm_init(sgmcmc.handle))
        finally:
            lib.sgm_set_progress(sgmcmc.handle, None, None)
        return Session(sgmcmc)
    _lib.check(lib.sgm_init(sgmcmc.handle))
    return Session(sgmcmc)


def saveState(sgmcmc: Sgmcmc, sess: Session) -> bytes:
    """The whole sampler state as an opaque blob (sgm_save_state): parameters, momenta, thermostats, the control-variate mode and its
    gradient, seed and position in the random streams.  The reference leaves persistence to the user of the step-by-step API
    (NEWS.md:19), who can save getParams() only."""
    lib = _lib.load()
    need = C.c_size_t()
    _lib.check(lib.sgm_save_state(sgmcmc.handle, None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    _lib.check(lib.sgm_save_state(sgmcmc.handle, buf, need.value, C.byref(need)))
    return buf.raw


def loadState(sgmcmc: Sgmcmc, blob: bytes) -> Session:
    """Restores saveState()'s blob into a sampler made by the same *Setup call (same dataset): takes the place of initSess -- the mode
    search is not repeated -- and sgmcmcStep / runs continue the saved run bit for bit."""
    buf = C.create_string_buffer(bytes(blob), len(blob))
    _lib.check(_lib.load().sgm_load_state(sgmcmc.handle, buf, len(blob)))
    return Session(sgmcmc)


def sgmcmcStep(sgmcmc: Sgmcmc, sess: Session):
    _lib.check(_lib.load().sgm_step(sgmcmc.handle))


def _unflatten(block: np.ndarray, lead: tuple, shp: tuple) -> np.ndarray:
    """block = lead + (size,) with the tensor's elements in R (column-major) order."""
    nd = len(shp)
    if nd <= 1:
        return block.reshape(lead + tuple(shp))
    arr = block.reshape(lead + tuple(reversed(shp)))
    k = len(lead)
    return arr.transpose(tuple(range(k)) + tuple(range(k + nd - 1, k - 1, -1)))


def _split_flat(sgmcmc: Sgmcmc, flat: np.ndarray, lead: tuple, scalar_as_1=False):
    """flat = lead + (P,) in flat parameter order -> dict of arrays lead + dims."""
    out, o = {}, 0
    for k, shp in sgmcmc.paramShapes.items():
        shp = tuple(shp)
        if scalar_as_1 and not shp:
            shp = (1,)
        sz = int(np.prod(shp)) if shp else 1
        out[k] = _unflatten(flat[..., o:o + sz], lead, shp)
        o += sz
    return out


def getParams(sgmcmc: Sgmcmc, sess: Session) -> Dict[str, np.ndarray]:
    lib = _lib.load()
    buf = np.empty((sgmcmc.nChains, sgmcmc.P), dtype=np.float64)
    _lib.check(lib.sgm_get_params(sgmcmc.handle, _dptr(buf)))
    lead = () if sgmcmc.nChains == 1 else (sgmcmc.nChains,)
    return _split_flat(sgmcmc, buf.reshape(lead + (sgmcmc.P,)), lead)


def estLogPost(sgmcmc: Sgmcmc, sess: Session) -> np.ndarray:
    """sess$run(sgmcmc$estLogPost, fresh minibatch) -- R/update.R:198-199."""
    out = np.empty(sgmcmc.nChains, dtype=np.float64)
    _lib.check(_lib.load().sgm_est_logpost(sgmcmc.handle, _dptr(out)))
    return out


def heldoutLogLik(sgmcmc: Sgmcmc, sess: Session, testset) -> np.ndarray:
    """logLik(current params, testset) per chain, on the device.  Replaces the user-built TensorFlow tensors that
    referenced ``sgld$params`` in the reference's vignettes (nn.Rmd:113-123: logLoss = -logLik / Ntest;
    logisticRegression.Rmd:89-115)."""
    lib = _lib.load()
    names = list(sgmcmc.dataNames)
    arrs = (_lib.SgmArray * len(names))()
    keep = []
    for i, nm in enumerate(names):
        v, row_major = _host_matrix(testset[nm])
        arrs[i].row_major = row_major
        keep.append(v)
        arrs[i].name = nm.encode(); arrs[i].data = _dptr(v); arrs[i].ndim = v.ndim
        for j, s_ in enumerate(v.shape):
            arrs[i].dims[j] = s_
    out = np.empty(sgmcmc.nChains, dtype=np.float64)
    _lib.check(lib.sgm_heldout_loglik(sgmcmc.handle, arrs, len(names), _dptr(out)))
    return out


def checkDivergence(sess: Session, sgmcmc: Sgmcmc, iter_: int, verbose: bool):
    """R/update.R:197-207."""
    est = estLogPost(sgmcmc, sess)
    if np.isnan(est).any():
        raise SgmcmcError(_lib.SGM_EDIVERGED, "Chain diverged, try decreasing stepsize")
    if verbose:
        _message(f"Iteration: {iter_}\t\tLog posterior estimate: {est[0] if est.size == 1 else est.mean()}")


def _pinned_doubles(count: int) -> np.ndarray:
    """A float64 array in page-locked host memory (freed when the array is collected)."""
    import weakref
    lib = _lib.load()
    ptr = lib.sgm_host_alloc(count * 8)
    if not ptr:
        return np.empty(count, dtype=np.float64)
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(count,))
    weakref.finalize(arr, lib.sgm_host_free, ptr)
    return arr


def getChains(sgmcmc: Sgmcmc, reuse_buffer: bool = False) -> Dict[str, np.ndarray]:
    """One copy-out of the device-resident chain as the reference's storage list
    (R/storage.R:42-51): per parameter an array (nIters, dims...); scalars give (nIters, 1).
    The arrays are zero-copy views of the single buffer the library fills in R's column-major layout."""
    lib = _lib.load()
    iters = int(lib.sgm_chain_length(sgmcmc.handle))
    C_, P = sgmcmc.nChains, sgmcmc.P
    # reuse_buffer: streaming consumers (chunked runs) may let successive calls overwrite one page-locked buffer
    # instead of paying a page-locked allocation per call; the default hands out independent arrays
    if reuse_buffer and getattr(sgmcmc, "_chain_buf", None) is not None and sgmcmc._chain_buf.size == C_ * P * iters:
        buf = sgmcmc._chain_buf
    else:
        buf = _pinned_doubles(C_ * P * iters)
        if reuse_buffer:
            sgmcmc._chain_buf = buf
    _lib.check(lib.sgm_copy_chain(sgmcmc.handle, _dptr(buf)))
    return _chain_views(sgmcmc, buf, iters)


def _chain_views(sgmcmc: Sgmcmc, buf: np.ndarray, iters: int) -> Dict[str, np.ndarray]:
    """The library's chain buffer ([chain][flat parameter][iteration] doubles) as the reference's storage list (R/storage.R:42-51):
    per parameter a zero-copy view (nIters, dims...); scalars give (nIters, 1); nChains > 1 adds a leading chain axis."""
    C_, P = sgmcmc.nChains, sgmcmc.P
    out, o = {}, 0
    for k, shp in sgmcmc.paramShapes.items():
        shp = tuple(shp) if len(shp) else (1,)
        sz = int(np.prod(shp))
        strides = [8]
        for dim in (iters,) + shp[:-1]:
            strides.append(strides[-1] * dim)
        view = np.lib.stride_tricks.as_strided(buf[o * iters:], shape=(C_, iters) + shp, strides=(P * iters * 8,) + tuple(strides))
        out[k] = view[0] if C_ == 1 else view
        o += sz
    return out


def runSGMCMC(sgmcmc: Sgmcmc, paramsRaw, options) -> Dict[str, np.ndarray]:
    """R/update.R:145-163.  The loop, the per-iteration storeState and the divergence check every 100 iterations all run on the device
    inside ONE library call (sgm_sample); the chain streams to the host while later iterations are sampled."""
    lib = _lib.load()
    verbose = options["verbose"]
    initSess(sgmcmc, verbose)
    nIters = int(options["nIters"])
    if verbose:
        _message("\nSimulating from SGMCMC algorithm...")

    def report(_user, phase, it, lp, n):   # checkDivergence's progress line (R/update.R:204-206)
        if verbose and phase == 2:
            v = lp[0] if n == 1 else float(np.mean([lp[i] for i in range(n)]))
            _message(f"Iteration: {it}\t\tLog posterior estimate: {v}")

    cb = _lib.PROGRESS_FN(report)
    buf = np.empty(sgmcmc.nChains * sgmcmc.P * nIters, dtype=np.float64)
    _lib.check(lib.sgm_set_progress(sgmcmc.handle, C.cast(cb, C.c_void_p), None))
    try:
        _lib.check(lib.sgm_sample(sgmcmc.handle, nIters, 100, _dptr(buf)))
    finally:
        lib.sgm_set_progress(sgmcmc.handle, None, None)
    return _chain_views(sgmcmc, buf, nIters)


# ---- one-shot samplers (R/sgld.R:50-58, sghmc.R:40-48, sgnht.R:38-46, *cv) ----------------------

def sgld(logLik, dataset, params, stepsize, logPrior=None, minibatchSize=0.01, nIters=10 ** 4,
         verbose=True, seed=None, **kw):
    o = sgldSetup(logLik, dataset, params, stepsize, logPrior, minibatchSize, seed, **kw)
    return runSGMCMC(o, params, {"nIters": nIters, "verbose": verbose})


def sghmc(logLik, dataset, params, stepsize, logPrior=None, minibatchSize=0.01, alpha=0.01, L=5,
          nIters=10 ** 4, verbose=True, seed=None, **kw):
    o = sghmcSetup(logLik, dataset, params, stepsize, logPrior, minibatchSize, alpha, L, seed, **kw)
    return runSGMCMC(o, params, {"nIters": nIters, "verbose": verbose})


def sgnht(logLik, dataset, params, stepsize, logPrior=None, minibatchSize=0.01, a=0.01,
          nIters=10 ** 4, verbose=True, seed=None, **kw):
    o = sgnhtSetup(logLik, dataset, params, stepsize, logPrior, minibatchSize, a, seed, **kw)
    return runSGMCMC(o, params, {"nIters": nIters, "verbose": verbose})


def sgldcv(logLik, dataset, params, stepsize, optStepsize, logPrior=None, minibatchSize=0.01,
           nIters=10 ** 4, nItersOpt=10 ** 4, verbose=True, seed=None, **kw):
    o = sgldcvSetup(logLik, dataset, params, stepsize, optStepsize, logPrior, minibatchSize, nItersOpt, verbose, seed, **kw)
    return runSGMCMC(o, params, {"nIters": nIters, "nItersOpt": nItersOpt, "verbose": verbose})


def sghmccv(logLik, dataset, params, stepsize, optStepsize, logPrior=None, minibatchSize=0.01,
            alpha=0.01, L=5, nIters=10 ** 4, nItersOpt=10 ** 4, verbose=True, seed=None, **kw):
    o = sghmccvSetup(logLik, dataset, params, stepsize, optStepsize, logPrior, minibatchSize, alpha, L,
                     nItersOpt, verbose, seed, **kw)
    return runSGMCMC(o, params, {"nIters": nIters, "nItersOpt": nItersOpt, "verbose": verbose})


def sgnhtcv(logLik, dataset, params, stepsize, optStepsize, logPrior=None, minibatchSize=0.01,
            a=0.01, nIters=10 ** 4, nItersOpt=10 ** 4, verbose=True, seed=None, **kw):
    o = sgnhtcvSetup(logLik, dataset, params, stepsize, optStepsize, logPrior, minibatchSize, a,
                     nItersOpt, verbose, seed, **kw)
    return runSGMCMC(o, params, {"nIters": nIters, "nItersOpt": nItersOpt, "verbose": verbose})
