"""ctypes binding of include/sgmcmc_b200.h (the C-ABI drop-in boundary).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) as
``sgmcmc_b200/lib/libsgmcmc_b200.so``.  There is no fallback: if the shared object
is missing the import of the product API fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

# the one library the product loads: the in-tree nvcc build (no environment override, no fallback)
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsgmcmc_b200.so")

SGM_ABI_VERSION = 1
SGM_OK, SGM_EINVAL, SGM_ECUDA, SGM_ENOMODEL, SGM_EDIVERGED, SGM_ENOMEM, SGM_ESTATE, SGM_ECOMM = 0, -1, -2, -3, -4, -5, -6, -7
MODEL_GAUSS_MEAN, MODEL_MIXTURE2, MODEL_LOGISTIC, MODEL_SOFTMAX_NET = 1, 2, 3, 4
ALGO_SGLD, ALGO_SGHMC, ALGO_SGNHT = 1, 2, 3
FLAG_SHARED_MODE = 1
NCCL_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


class SgmArray(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_double)), ("ndim", C.c_int32), ("row_major", C.c_int32), ("dims", C.c_int64 * 4)]


class SgmParam(C.Structure):
    _fields_ = [("name", C.c_char_p), ("init", C.POINTER(C.c_double)), ("ndim", C.c_int32), ("dims", C.c_int64 * 4),
                ("stepsize", C.c_double), ("alpha", C.c_double), ("a", C.c_double)]


class SgmConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("algo", C.c_int32), ("cv", C.c_int32),
        ("n_data", C.c_int32), ("data", C.POINTER(SgmArray)),
        ("n_params", C.c_int32), ("params", C.POINTER(SgmParam)),
        ("minibatch_size", C.c_double), ("L", C.c_int32), ("opt_stepsize", C.c_double), ("n_iters_opt", C.c_int64),
        ("seeded", C.c_int32), ("seed", C.c_uint64), ("n_chains", C.c_int32), ("chain_offset", C.c_int64), ("device", C.c_int32), ("flags", C.c_uint32),
        ("use_prior", C.c_int32), ("prior_sd", C.c_double), ("precision", C.POINTER(C.c_double)),
        ("synth_rows", C.c_int64), ("synth_dim", C.c_int64), ("synth_seed", C.c_uint64),
    ]


class SgmPlanQuery(C.Structure):
    _fields_ = [("model", C.c_int32), ("cv", C.c_int32), ("N", C.c_int64), ("d", C.c_int32), ("d_hidden", C.c_int32), ("d_out", C.c_int32),
                ("n", C.c_int32), ("n_chains", C.c_int32), ("flags", C.c_uint32), ("world", C.c_int32), ("rank", C.c_int32),
                ("sharded_dataset", C.c_int32), ("row0", C.c_int64), ("row1", C.c_int64), ("sms", C.c_int32)]


# every symbol include/sgmcmc_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_PD = C.POINTER(C.c_double)
_PF = C.POINTER(C.c_float)
_PI = C.POINTER(C.c_int32)
SYMBOLS = {
    "sgm_create": (C.c_int, [C.POINTER(SgmConfig), C.POINTER(_H)]),
    "sgm_init": (C.c_int, [_H]),
    "sgm_init_deterministic": (C.c_int, [_H, _PF, _PI]),
    "sgm_destroy": (None, [_H]),
    "sgm_step": (C.c_int, [_H]),
    "sgm_step_deterministic": (C.c_int, [_H, _PI, _PF]),
    "sgm_run": (C.c_int, [_H, C.c_int64, C.c_int]),
    "sgm_sample": (C.c_int, [_H, C.c_int64, C.c_int32, _PD]),
    "sgm_get_params": (C.c_int, [_H, _PD]),
    "sgm_get_state": (C.c_int, [_H, _PD, _PD]),
    "sgm_get_mode": (C.c_int, [_H, _PD, _PD]),
    "sgm_save_state": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sgm_load_state": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "sgm_chain_length": (C.c_int64, [_H]),
    "sgm_copy_chain": (C.c_int, [_H, _PD]),
    "sgm_reset_chain": (C.c_int, [_H]),
    "sgm_host_alloc": (C.c_void_p, [C.c_size_t]),
    "sgm_host_free": (None, [C.c_void_p]),
    "sgm_reserve_chain": (C.c_int, [_H, C.c_int64]),
    "sgm_est_logpost": (C.c_int, [_H, _PD]),
    "sgm_est_logpost_deterministic": (C.c_int, [_H, _PI, _PD]),
    "sgm_grad_estimate": (C.c_int, [_H, _PI, _PD]),
    "sgm_heldout_loglik": (C.c_int, [_H, C.POINTER(SgmArray), C.c_int32, _PD]),
    "sgm_num_rows": (C.c_int64, [_H]),
    "sgm_minibatch": (C.c_int32, [_H]),
    "sgm_num_scalars": (C.c_int32, [_H]),
    "sgm_noise_len": (C.c_int32, [_H]),
    "sgm_last_timing": (C.c_int, [_H, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "sgm_last_fullgrad_timing": (C.c_int, [_H, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "sgm_last_collective_timing": (C.c_int, [_H, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "sgm_bench_full_gradient": (C.c_int, [_H, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "sgm_last_fullgrad_flops": (C.c_int, [_H, _PD]),
    "sgm_dataset_ipc_handle": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "sgm_attach_peer_shards": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "sgm_detach_peer_shards": (C.c_int, [_H]),
    "sgm_set_progress": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "sgm_debug_plan": (C.c_int, [C.POINTER(SgmPlanQuery), C.c_char_p, C.c_size_t]),
    "sgm_set_option": (C.c_int, [_H, C.c_char_p, C.c_int32]),
    "sgm_debug_phase_cycles": (C.c_int, [_H, C.c_int, C.POINTER(C.c_ulonglong), C.c_int]),
    "sgm_copy_dataset": (C.c_int, [_H, C.c_int32, C.c_int64, C.c_int64, _PD]),
    "sgm_comm_unique_id": (C.c_int, [C.c_void_p, C.c_int32]),
    "sgm_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "sgm_last_error": (C.c_char_p, []),
    "sgm_version": (C.c_char_p, []),
}

PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double), C.c_int32)

_lib = None


class SgmcmcError(RuntimeError):
    """Raised for every non-zero status; mirrors R's stop() (SURVEY.md 8b "Errors")."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def load():
    """dlopen the C-ABI library (once) and type every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(sgmcmc_b200 has no CPU / Python fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int):
    if rc != SGM_OK:
        msg = load().sgm_last_error()
        raise SgmcmcError(rc, msg.decode() if msg else f"sgmcmc_b200 error {rc}")
