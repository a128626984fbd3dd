"""The `.Call` shim of the R package (r-package/src/rshim.c) on top of a stand-in R C API (tests/r_stub/): R itself is absent
from the build image, so this is as far as the R-side binding can be exercised here.

  * CPU: the shim compiles against the stub headers, links against libsgmcmc_b200.so and registers its 14 .Call routines;
  * GPU: the driver makes the calls the R files make (create / init / step / run / get_params / copy_chain / est_logpost) and
    must return bit-identical doubles to the same calls through the Python mirror (same library, same seed).
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "r_stub")


def _build(tmp_path, library=None):
    """The shim + stand-in R API + driver, linked against the product library (default) or a given build of the same C ABI."""
    if library is None:
        import __graft_entry__ as g
        g.build()
        library = os.path.join(ROOT, "sgmcmc_b200", "lib", "libsgmcmc_b200.so")
    exe = str(tmp_path / "shim_driver")
    libdir, libname = os.path.dirname(library), os.path.basename(library)
    cmd = ["gcc", "-std=gnu11", "-O1", f"-I{STUB}", f"-I{os.path.join(ROOT, 'include')}", "-o", exe,
           os.path.join(ROOT, "r-package", "src", "rshim.c"), os.path.join(STUB, "rstub.c"), os.path.join(STUB, "shim_driver.c"),
           f"-L{libdir}", f"-l:{libname}", f"-Wl,-rpath,{libdir}", "-lm"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    return exe


def test_shim_compiles_links_and_registers(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "check"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "registered 14 .Call routines" in out.stdout


def _hash01(i, j):
    x = (np.uint64(i) * np.uint64(2654435761) & np.uint64(0xFFFFFFFF)) ^ (np.uint64(j + 1) * np.uint64(2246822519) & np.uint64(0xFFFFFFFF))
    x ^= x >> np.uint64(15)
    x = x * np.uint64(2654435761) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    return float(x & np.uint64(0xFFFFFF)) / 16777216.0


def _shim_vs_mirror(exe):
    """Runs the driver (the calls the R files make) and the same calls through the Python mirror on the CURRENTLY LOADED library."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import _lib
    from sgmcmc_b200 import api as A
    out = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = {l.split()[0]: l.split()[1:] for l in out.stdout.splitlines() if l.strip()}
    got_params = np.array(lines["params"], dtype=np.float64)
    got_last = np.array(lines["last_stored"], dtype=np.float64)
    assert lines["logpost_finite"] == ["1"]
    # the shim printed the reference's progress lines of the verbose run (50 optimiser iterations: none; 230 sampling iterations: two)
    assert [l.split()[1] for l in out.stderr.splitlines() if l.startswith("Iteration:")] == ["100", "200"], out.stderr

    N, d, chains = 2000, 6, 3
    X = np.empty((N, d))
    y = np.empty(N)
    for i in range(N):
        z = 0.1
        for j in range(d):
            v = 2.0 * _hash01(i, j) - 1.0
            X[i, j] = v
            z += v * (0.3 * (j + 1) - 1.0)
        y[i] = 1.0 if _hash01(i, 77) < 1.0 / (1.0 + np.exp(-z)) else 0.0
    mk = lambda: S.sgldcvSetup(S.logisticLogLik(), {"X": X, "y": y}, {"bias": np.array(0.0), "beta": np.zeros((d, 1))}, 1e-4, 1e-4,   # noqa: E731
                               logPrior=S.laplacePrior(), minibatchSize=100, nItersOpt=50, verbose=False, seed=11, nChains=chains)
    sg = mk()
    sess = S.initSess(sg, verbose=False)
    S.sgmcmcStep(sg, sess)
    _lib.check(_lib.load().sgm_run(sg.handle, 20, 1))
    th = np.empty((chains, sg.P))
    _lib.check(_lib.load().sgm_get_params(sg.handle, th.ctypes.data_as(_lib._PD)))
    assert int(lines["chain_doubles"][0]) == 20 * sg.P * chains
    np.testing.assert_array_equal(got_params, th.reshape(-1))
    np.testing.assert_array_equal(got_last, th[0])   # chain 0: the last stored row is the current state
    # runSGMCMC through C_sgm_sample and heldoutLogLik
    sg2 = mk()
    ch = A.runSGMCMC(sg2, None, {"nIters": 230, "verbose": False})
    last = np.concatenate([np.asarray(ch["bias"])[0, -1].reshape(-1), np.asarray(ch["beta"])[0, -1].reshape(-1)])
    assert int(lines["sampled_doubles"][0]) == 230 * sg2.P * chains
    np.testing.assert_array_equal(np.array(lines["sampled_last"], dtype=np.float64), last)
    np.testing.assert_array_equal(np.array(lines["sampled_state"], dtype=np.float64), last)
    ho = S.heldoutLogLik(sg2, None, {"X": X[:300], "y": y[:300]})
    np.testing.assert_array_equal(np.array(lines["heldout"], dtype=np.float64), ho)
    # saveState / loadState through the shim: the restored sampler continued the run bit for bit; the blob is the mirror's size
    assert lines["resume_equal"] == ["1"]
    assert int(lines["state_blob_bytes"][0]) == len(S.saveState(sg2, None))


@pytest.mark.gpu
def test_shim_calls_match_python_mirror(lib, tmp_path):
    """GPU: the driver and the Python mirror on the product library, bit-identical doubles."""
    _shim_vs_mirror(_build(tmp_path))


def test_shim_calls_match_python_mirror_on_the_emulated_build(tmp_path):
    """CPU: the same comparison with both sides on the CPU-emulated build of the kernel sources (tests/emu): the .Call layer -- argument
    marshalling, R-layout outputs, progress lines, error mapping -- is exercised end to end without R and without a GPU."""
    from tests.emu import backend
    with backend.emulated():
        from tests.emu import build as emu_build
        _shim_vs_mirror(_build(tmp_path, emu_build.lib_path()))
