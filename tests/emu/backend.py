"""TEST INFRASTRUCTURE ONLY: swaps the ctypes binding used by sgmcmc_b200 for the CPU-emulated build of the same kernel sources
(tests/emu/build.py) for the duration of a test.  The product never does this: sgmcmc_b200/_lib.py knows one path, the nvcc build."""
from __future__ import annotations

import contextlib
import ctypes as C
import gc
import os
import subprocess
import sys

from . import build as emu_build

_emu = {}

# extra builds of the same sources: kernel generations that are compiled out of the product by default
VARIANTS = {"": (), "s5tmem": ("-DNN_ENABLE_S5_TMEM",)}


def library(variant: str = ""):
    """The emulated library, typed like the real one (built on first use; under a minute of g++)."""
    if variant not in _emu:
        from sgmcmc_b200 import _lib
        path = emu_build.build(variant=variant, defs=VARIANTS[variant])
        lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in _lib.SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib.emu_last_message.restype = C.c_char_p
        _emu[variant] = lib
    return _emu[variant]


@contextlib.contextmanager
def emulated(variant: str = ""):
    """`with emulated():` -- every C-ABI call made through sgmcmc_b200 goes to the emulated build."""
    from sgmcmc_b200 import _lib
    lib = library(variant)
    lib.emu_reset_error()
    saved = _lib._lib
    _lib._lib = lib
    try:
        yield lib
    finally:
        gc.collect()   # handles created on the emulated library are destroyed by it
        _lib._lib = saved


def run_isolated(code: str, env: dict | None = None, timeout: int = 600):
    """Runs a python snippet in a fresh interpreter (the emulator reads SGM_EMU_* once per process); returns CompletedProcess."""
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    e = dict(os.environ)
    e.update(env or {})
    e["PYTHONPATH"] = root + os.pathsep + e.get("PYTHONPATH", "")
    return subprocess.run([sys.executable, "-c", code], cwd=root, env=e, capture_output=True, text=True, timeout=timeout)
