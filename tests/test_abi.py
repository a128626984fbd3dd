"""CPU suite: the C-ABI library loads, exports every symbol include/sgmcmc_b200.h declares,
and its host-side logic (validation, error text) behaves like the reference's -- no compute
calls (there is no GPU here and no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sgmcmc_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    from sgmcmc_b200 import _lib
    names = declared_functions()
    assert len(names) >= 25
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (sgm_[a-z0-9_]+)", nm))
    assert set(names) <= exported, set(names) - exported
    assert b"sm_100a" in lib.sgm_version()


def test_config_struct_layout_matches_header(lib):
    """sizeof/offsets computed by gcc from the header equal the ctypes mirror."""
    from sgmcmc_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "sgmcmc_b200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(sgm_config), sizeof(sgm_array), sizeof(sgm_param),
 offsetof(sgm_config, minibatch_size), offsetof(sgm_config, seed), offsetof(sgm_config, precision), offsetof(sgm_config, synth_seed));return 0;}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.SgmConfig), C.sizeof(_lib.SgmArray), C.sizeof(_lib.SgmParam), _lib.SgmConfig.minibatch_size.offset,
            _lib.SgmConfig.seed.offset, _lib.SgmConfig.precision.offset, _lib.SgmConfig.synth_seed.offset]
    assert got == want


def _expect(fn, code, text):
    import sgmcmc_b200 as S
    with pytest.raises(S.SgmcmcError) as e:
        fn()
    assert e.value.code == code and text in str(e.value), (e.value.code, str(e.value))


def test_host_validation_and_reference_error_text(lib):
    import sgmcmc_b200 as S
    from sgmcmc_b200 import _lib
    X = np.random.RandomState(0).randn(50, 2)
    th = {"theta": np.zeros(2)}
    # closures are not routed: no fallback
    _expect(lambda: S.sgldSetup(lambda p, d: 0.0, {"X": X}, th, 1e-4), _lib.SGM_ENOMODEL, "no CPU fallback")
    _expect(lambda: S.sgldSetup(S.gaussMeanLogLik(), {"X": X}, th, 1e-4, logPrior=S.laplacePrior()), _lib.SGM_ENOMODEL, "normalPrior")
    # convertProp's message (R/setup.R:132-133)
    _expect(lambda: S.sgldSetup(S.gaussMeanLogLik(), {"X": X}, th, 1e-4, minibatchSize=0.001), _lib.SGM_EINVAL,
            "minibatchSize proportion (0.001) too small for dataset size (50).")
    _expect(lambda: S.sgldSetup(S.gaussMeanLogLik(), {"X": X}, {"theta": np.zeros(3)}, 1e-4, minibatchSize=10), _lib.SGM_EINVAL, "must have 2 entries")
    _expect(lambda: S.sgldSetup(S.gaussMeanLogLik(), {"Y": X}, th, 1e-4), _lib.SGM_EINVAL, "dataset entry 'X'")
    _expect(lambda: S.sgldSetup(S.logisticLogLik(), {"X": X, "y": np.zeros(49)}, {"bias": 0.0, "beta": np.zeros((2, 1))}, 1e-4, minibatchSize=5),
            _lib.SGM_EINVAL, "disagree")
    _expect(lambda: S.sgldcvSetup(S.gaussMeanLogLik(), {"X": X}, th, 1e-4, 0.0, minibatchSize=5), _lib.SGM_EINVAL, "optStepsize")


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sgmcmc_b200 as S
    from sgmcmc_b200 import _lib
    X = np.random.RandomState(0).randn(50, 2)
    _expect(lambda: S.sgld(S.gaussMeanLogLik(), {"X": X}, {"theta": np.zeros(2)}, 1e-4, minibatchSize=5, nIters=3, verbose=False),
            _lib.SGM_ECUDA, "no CUDA device")


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sgmcmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_every_export_is_mapped_to_the_reference_in_integration_md():
    """INTEGRATION.md section 1 is the drop-in contract: every function the header declares has a row naming what it replaces."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "sgmcmc_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    exports = sorted(set(re.findall(r"\b(sgm_[a-z0-9_]+)\s*\(", header)) - {"sgm_progress_fn"})
    assert len(exports) >= 40
    missing = [f for f in exports if f"`{f}`" not in doc]
    assert not missing, missing
