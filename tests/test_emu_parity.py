"""The kernel SOURCES against the oracle in the GPU-less container: csrc/*.cu(h) compiled by g++ against the CPU emulation of
tests/emu (fibers, shuffles, cp.async, mbarrier, TMA, tcgen05 / TMEM), driven through the same C ABI and the same test bodies as
the `-m gpu` suite (tests/test_parity_gpu.py, tests/test_fullgrad_tc_gpu.py).

What this is: a check of the kernels' logic, indexing, synchronisation protocol and host dispatch (launch plans, shared-memory
budgets, tensor maps, TMA coordinates) that runs where no GPU exists -- the failure classes that were Xid faults in round 1 fail
here as test errors.  What it is not: GPU parity.  The bit-level claims are made by the `-m gpu` tests on a B200 only.
"""
import numpy as np
import pytest

from tests import test_fullgrad_tc_gpu as F
from tests import test_parity_gpu as T
from tests.cases import make_case
from tests.emu import backend


@pytest.fixture(autouse=True)
def emulated_library():
    with backend.emulated() as lib:
        yield lib
        msg = lib.emu_last_message()
        assert not msg, msg.decode()


def ran(lib, path: str) -> int:
    """How often the named kernel path executed in the emulated build (SGM_EMU_NOTE in the kernel sources)."""
    import ctypes as C
    lib.emu_note_count.restype = C.c_longlong
    lib.emu_note_count.argtypes = [C.c_char_p]
    return int(lib.emu_note_count(path.encode()))


FAMILIES = ["gauss1d", "gauss2d", "gauss3d_fullcov", "gauss2d_flat", "mixture", "logistic5", "logistic54", "logistic130"]
BENCHED = ["logistic54_n500", "logistic1024_n1024_N3000",   # the template instantiations bench.py times (c3: R = 2; c5: <32, 9>)
           "logistic54_n101"]                               # an odd minibatch size (mbarrier alignment behind the coefficient array)


@pytest.mark.parametrize("case_name", FAMILIES + BENCHED)
@pytest.mark.parametrize("algo", T.ALGOS)
@pytest.mark.parametrize("mode", ["plain", "cv", "cv_shared"])
def test_deterministic_steps_match_oracle(case_name, algo, mode):
    if mode == "cv_shared" and case_name in ("gauss1d", "gauss3d_fullcov", "gauss2d_flat"):
        pytest.skip("shared-mode variant is covered by gauss2d")
    # (12 mode-finder steps instead of the GPU suite's 25: the float32 autograd replica of the oracle is what takes the time here)
    T._deterministic_steps(case_name, algo, mode != "plain", shared=(mode == "cv_shared"), n_opt=None if case_name.startswith("logistic1024") else 12)


@pytest.mark.parametrize("case_name,algo,cv", [("gauss2d", "sgld", False), ("mixture", "sghmc", False), ("logistic54", "sgld", True),
                                               ("logistic130", "sgnht", True), ("gauss3d_fullcov", "sgnht", False), ("logistic5", "sghmc", True)])
def test_device_rng_streams_match_oracle(case_name, algo, cv):
    T.test_device_rng_streams_match_oracle(None, case_name, algo, cv)


@pytest.mark.parametrize("case_name", ["gauss2d", "mixture", "logistic54"])
def test_est_logpost_and_heldout(case_name):
    T.test_est_logpost_matches_oracle(None, case_name)
    T.test_heldout_loglik_matches_oracle(None, case_name)


@pytest.mark.parametrize("case_name,algo", [("gauss2d", "sgld"), ("mixture", "sghmc"), ("gauss3d_fullcov", "sgnht"), ("gauss1d", "sgld")])
def test_small_models_shared_memory_dataset(emulated_library, case_name, algo):
    before = ran(emulated_library, "small_chain:smem")
    T.test_small_models_shared_memory_dataset_equals_global_gather(None, case_name, algo)
    assert ran(emulated_library, "small_chain:smem") >= before + 2, "the default dispatch did not pick the shared-memory variant for a 60-iteration launch"


@pytest.mark.parametrize("case_name,algo,cv,n_iters,n_chains", [("gauss2d", "sgld", False, 250, 3), ("mixture", "sghmc", False, 137, 3), ("logistic54", "sgld", True, 230, 3),
                                                                ("logistic130", "sgnht", True, 101, 3), ("softmaxnet_small", "sghmc", False, 101, 2),
                                                                ("gauss2d", "sgld", False, 233, 340)])
def test_streamed_sampling(case_name, algo, cv, n_iters, n_chains, capsys):
    T.test_streamed_sampling_equals_run_then_copy(None, case_name, algo, cv, n_iters, n_chains, capsys)


@pytest.mark.parametrize("case_name", ["gauss2d", "logistic54", "softmaxnet_small"])
def test_dataset_layouts_upload_identically(case_name):
    T.test_dataset_layouts_upload_identically(None, case_name)


@pytest.mark.parametrize("case_name,algo,cv", [("gauss2d", "sgld", False), ("mixture", "sghmc", False), ("logistic54", "sgnht", True),
                                               ("logistic130", "sgld", True), ("softmaxnet_small", "sghmc", False)])
def test_call_sequences_walk_one_rng_stream(case_name, algo, cv):
    T.test_call_sequences_walk_one_rng_stream(None, case_name, algo, cv)


def test_streamed_sampling_reports_divergence():
    T.test_streamed_sampling_reports_divergence(None)


@pytest.mark.parametrize("case_name", ["gauss2d", "mixture", "logistic54", "logistic130", "softmaxnet_small"])
@pytest.mark.parametrize("algo", T.ALGOS)
def test_last_stored_state_is_the_sampler_state(case_name, algo):
    T.test_last_stored_state_is_the_sampler_state(None, case_name, algo)


@pytest.mark.parametrize("case_name,algo,cv,shared", [("mixture", "sgld", True, True), ("logistic5", "sgld", False, False)])
def test_posterior_moments_match_the_oracle(case_name, algo, cv, shared):
    """Reduced-size run of tests/test_statistics_gpu.py::test_posterior_moments_match_the_oracle (device RNG vs float64 oracle chains)."""
    from tests import test_statistics_gpu as ST
    ST.test_posterior_moments_match_the_oracle(None, case_name, algo, cv, shared, n_dev=32, n_orc=8, burn=200, keep=400)


@pytest.mark.parametrize("case_name,algo,cv,shared", [("gauss2d", "sgld", False, False), ("mixture", "sghmc", True, False), ("logistic54", "sgnht", True, True),
                                                      ("logistic130", "sghmc", True, False), ("softmaxnet_small", "sgnht", False, False)])
def test_saved_state_resumes_the_run(case_name, algo, cv, shared):
    T.test_saved_state_resumes_the_run_bit_for_bit(None, case_name, algo, cv, shared)


@pytest.mark.parametrize("case_name,algo,cv,n_iters,n_chains", [("logistic54", "sgld", True, 230, 3), ("gauss2d", "sghmc", False, 133, 40), ("softmaxnet_small", "sgnht", False, 101, 2)])
def test_sampling_into_a_page_locked_array(emulated_library, case_name, algo, cv, n_iters, n_chains):
    before = ran(emulated_library, "sample:direct_to_host")
    T.test_sampling_into_a_page_locked_array_equals_the_staged_copy(None, case_name, algo, cv, n_iters, n_chains)
    assert ran(emulated_library, "sample:direct_to_host") > before, "the page-locked array was not written by the device directly"


def test_mode_finder_cadence(capsys):
    T.test_mode_finder_reports_and_checks_every_100_iterations(None, capsys)


def test_chain_storage_layout_and_seeds():
    T.test_chain_storage_layout_and_shapes(None)
    T.test_same_seed_reproduces(None)


# ---- 784-100-10 softmax net: TMA-fed S1 (tile::gather4 + 3-D box, 128-byte swizzle), S5 on tcgen05, streamed B update (TMA bulk copies
# and bulk stores), the register-staged fall-back at the reduced width
@pytest.mark.parametrize("case_name,algo,mode", [
    ("softmaxnet_small", "sgld", "plain"), ("softmaxnet_small", "sghmc", "cv"), ("softmaxnet_small", "sgnht", "cv_shared"),
    ("softmaxnet_small", "sgld", "cv"), ("softmaxnet_small", "sghmc", "plain"), ("softmaxnet_small", "sgnht", "plain"),
    ("softmaxnet", "sghmc", "plain"), ("softmaxnet", "sgld", "cv"), ("softmaxnet", "sgnht", "cv_shared")])
def test_softmax_net_steps_match_oracle(emulated_library, case_name, algo, mode):
    before = ran(emulated_library, "nn_stage1_tma")
    T._deterministic_steps(case_name, algo, mode != "plain", shared=(mode == "cv_shared"))
    assert ran(emulated_library, "nn_stage1_tma") > before, "the TMA-fed forward GEMM did not run"


def test_softmax_net_register_staged_paths():
    """sgm_set_option: the register-staged forward GEMM (no TMA) and the register-staged update give the same chain."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import validation as V
    case = make_case("softmaxnet_small")
    states = []
    for opts in ((), (("nn_stage1", 1),), (("nn_update", 1),)):
        dev = T.device_setup(case, "sghmc", True, n_chains=2, n_iters_opt=3)
        for k, v in opts:
            V.setOption(dev, k, v)
        S.initSess(dev, verbose=False)
        S.sgmcmcStep(dev, None)
        states.append(V.flatParams(dev))
    for other in states[1:]:
        assert np.abs(states[0] - other).max() <= 2e-5 * np.abs(states[0]).max()


def test_softmax_net_misc():
    T.test_est_logpost_matches_oracle(None, "softmaxnet_small")
    T.test_heldout_loglik_matches_oracle(None, "softmaxnet_small")
    T.test_device_rng_streams_match_oracle(None, "softmaxnet_small", "sghmc", False)


def test_softmax_net_stage5_on_the_tensor_memory_side():
    """nn_stage5_tmem: written at the end of round 1, compiled out of the product (-DNN_ENABLE_S5_TMEM), never run on a GPU.  Here it
    runs in a second emulated build, selected through sgm_set_option, and must give the oracle's chain."""
    from sgmcmc_b200 import validation as V
    from tests import cases
    orig = T.device_setup

    def with_s5(*a, **k):
        dev = orig(*a, **k)
        V.setOption(dev, "nn_stage5", 1)
        return dev

    with backend.emulated("s5tmem") as lib:
        T.device_setup = with_s5
        try:
            before = ran(lib, "nn_stage5_tmem")
            T._deterministic_steps("softmaxnet", "sghmc", True, shared=False)
            T._deterministic_steps("softmaxnet", "sgld", False, shared=False)
            assert ran(lib, "nn_stage5_tmem") > before
            assert not lib.emu_last_message()
        finally:
            T.device_setup = orig
    del cases


# ---- every gather engine of the short-row kernel and the generic kernel on the same case (sgm_set_option) -----------------------
@pytest.mark.parametrize("option,value,path", [("gather", 1, "logistic_staged:cp.async"), ("gather", 2, "logistic_staged:bulk"), ("gather", 3, "logistic_staged:bulk"),
                                               ("gather", 4, "logistic_staged:gather4"), ("logistic_kernel", 1, "logistic_chain")])
@pytest.mark.parametrize("case_name", ["logistic54", "logistic54_n500"])
def test_logistic_kernel_variants(emulated_library, case_name, option, value, path):
    before = ran(emulated_library, path)
    import sgmcmc_b200 as S
    from sgmcmc_b200 import validation as V
    case = make_case(case_name)
    rs = np.random.RandomState(9)
    C_ = 3
    states = []
    for pinned in (False, True):
        dev = T.device_setup(case, "sghmc", True, n_chains=C_, n_iters_opt=8, sharedMode=False)
        if pinned:
            V.setOption(dev, option, value)
        S.initSess(dev, verbose=False)
        for _ in range(2):
            S.sgmcmcStep(dev, None)
        states.append(V.flatParams(dev))
    scale = np.abs(states[0]).max()
    assert np.abs(states[0] - states[1]).max() <= 1e-5 * scale, "a pinned kernel path disagrees with the default dispatch"
    assert ran(emulated_library, path) > before, f"{path} did not run"
    del rs


# ---- mode finder on a thread-block cluster per mode (logistic_split.cuh): cluster launch, barrier.cluster, DSMEM loads ------
@pytest.mark.parametrize("case_name,split,shared", [("logistic54", 8, True), ("logistic54_n500", 4, True), ("logistic5", 2, True), ("logistic130", 8, True),
                                                    ("logistic54_n101", 4, True), ("logistic1024_n1024_N3000", 8, True), ("logistic54", 2, False), ("logistic130", 4, False),
                                                    ("logistic54_n20", 8, True)])
def test_mode_finder_cluster_split(emulated_library, case_name, split, shared):
    before = ran(emulated_library, "logistic_opt_split")
    T._deterministic_steps(case_name, "sgld", True, shared=shared, options={"mode_finder_split": split}, n_opt=6 if "1024" in case_name else 25)
    assert ran(emulated_library, "logistic_opt_split") > before, "the cluster mode finder did not run"


# every sampler x every mode on the covertype shape; the other shapes (odd minibatch, tiny rows, padded rows, long rows) with the
# sampler and the mode rotating (the GPU suite runs the full cross: tests/test_parity_gpu.py)
_CLUSTER_CASES = [("logistic54", a, m) for a in T.ALGOS for m in ("plain", "cv", "cv_shared")] + [
    ("logistic54_n101", "sgld", "cv"), ("logistic54_n101", "sghmc", "cv_shared"), ("logistic54_n101", "sgnht", "plain"),
    ("logistic5", "sgld", "cv_shared"), ("logistic5", "sghmc", "plain"), ("logistic5", "sgnht", "cv"),
    ("logistic130", "sgld", "plain"), ("logistic130", "sghmc", "cv"), ("logistic130", "sgnht", "cv_shared"),
    ("logistic1024_n1024_N3000", "sgld", "cv"), ("logistic1024_n1024_N3000", "sgnht", "cv_shared")]


@pytest.mark.parametrize("case_name,algo,mode", _CLUSTER_CASES)
def test_cluster_per_chain_steps_match_oracle(emulated_library, case_name, algo, mode):
    before = ran(emulated_library, "logistic_split:sample")
    for split in ((2, 4, 8) if case_name == "logistic54_n101" else (8,)):
        T.test_cluster_per_chain_steps_match_oracle(None, case_name, algo, mode, split)
    assert ran(emulated_library, "logistic_split:sample") > before, "the cluster kernel did not run the sampling steps"


@pytest.mark.parametrize("case_name,algo,cv,shared", [("logistic54", "sgld", True, True), ("logistic54", "sghmc", False, False), ("logistic130", "sgnht", True, False),
                                                      ("logistic5", "sgld", False, False)])
def test_cluster_per_chain_walks_the_same_chain(case_name, algo, cv, shared):
    T.test_cluster_per_chain_walks_the_same_chain(None, case_name, algo, cv, shared)


def test_dataset_upload_in_many_slabs():
    """The upload converts slabs of rows to float32 on host threads into page-locked staging, two slabs in flight (32 MB each on a
    device: the toy datasets here are one slab).  A fresh interpreter with 4 KB slabs: C-order / column-major / strided inputs of every
    family upload identically over dozens of slabs with ragged tails, the held-out set included."""
    code = """
from tests.emu import backend
from tests import test_parity_gpu as T
with backend.emulated() as lib:
    for case in ("gauss2d", "logistic54", "softmaxnet_small", "logistic130"):
        T.test_dataset_layouts_upload_identically(None, case)
    T.test_heldout_loglik_matches_oracle(None, "logistic54")
    T.test_heldout_loglik_matches_oracle(None, "mixture")
    T._deterministic_steps("logistic54_n101", "sgld", True, shared=True)
    assert not lib.emu_last_message()
print("slab upload ok")
"""
    out = backend.run_isolated(code, env={"SGM_EMU_UPLOAD_SLAB": "4096"})
    assert out.returncode == 0 and "slab upload ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_default_dispatch_on_a_device_with_many_sms():
    """The emulated device reports 2 SMs by default, so the library's default dispatch keeps one CTA per chain in most tests here; on a
    B200 (148 SMs) a handful of chains get a cluster each.  A fresh interpreter with a 64-SM emulated device and the samplers left on
    the default dispatch: call sequences, save / restore, page-locked output, seeds, device RNG streams against the oracle."""
    code = """
from tests.emu import backend
from tests import cases
orig = cases.device_setup
def unpinned(*a, **k):
    k.setdefault("pin_finder", False)
    return orig(*a, **k)
cases.device_setup = unpinned
from tests import test_parity_gpu as T
T.device_setup = unpinned
import ctypes as C
with backend.emulated() as lib:
    lib.emu_note_count.restype = C.c_longlong; lib.emu_note_count.argtypes = [C.c_char_p]
    T.test_call_sequences_walk_one_rng_stream(None, "logistic54", "sgnht", True)
    T.test_saved_state_resumes_the_run_bit_for_bit(None, "logistic130", "sghmc", True, False)
    T.test_sampling_into_a_page_locked_array_equals_the_staged_copy(None, "logistic54", "sgld", True, 230, 3)
    T.test_last_stored_state_is_the_sampler_state(None, "logistic54", "sghmc")
    T.test_same_seed_reproduces(None)
    T.test_device_rng_streams_match_oracle(None, "logistic54", "sgld", True)
    T.test_device_rng_streams_match_oracle(None, "logistic130", "sgnht", True)
    assert lib.emu_note_count(b"logistic_split:sample") > 500 and lib.emu_note_count(b"logistic_opt_split") > 5
    assert not lib.emu_last_message()
print("default dispatch ok")
"""
    out = backend.run_isolated(code, env={"SGM_EMU_SMS": "64"})
    assert out.returncode == 0 and "default dispatch ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_mode_finder_cluster_split_equals_one_cta_per_mode(emulated_library):
    """Device-drawn minibatches: the cluster of CTAs draws exactly the rows the one-CTA finder draws (same keyed permutation, positions
    split over the ranks), so both arrive at the same mode up to the order of the float32 sums; 137 steps cross the 100-iteration
    launch boundary of checkOptDivergence."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import validation as V
    case = make_case("logistic54")
    modes = []
    for split in (1, 4):
        dev = T.device_setup(case, "sgld", True, n_chains=2, n_iters_opt=137, sharedMode=True)
        V.setOption(dev, "mode_finder_split", split)
        S.initSess(dev, verbose=False)
        modes.append(V.getMode(dev)[0])
    scale = np.abs(modes[0]).max()
    assert np.abs(modes[0] - modes[1]).max() <= 1e-5 * scale


# ---- tcgen05 many-mode full gradient (cvg_kernel): TMA 128-byte-swizzled tiles, UMMA descriptors, A operand in TMEM -------------
@pytest.mark.parametrize("case_name,chains", [("logistic54", 20), ("logistic5", 16), ("logistic130", 40)])
def test_many_mode_full_gradient(emulated_library, case_name, chains):
    before = ran(emulated_library, "cvg:pass1"), ran(emulated_library, "cvg:pass2")
    F.test_many_mode_full_gradient_matches_oracle(None, case_name, chains)
    assert ran(emulated_library, "cvg:pass1") > before[0] and ran(emulated_library, "cvg:pass2") > before[1]


@pytest.mark.parametrize("case_name,chains", [("logistic128_n20_N300", 17), ("logistic127_n20_N333", 33), ("logistic129_n20_N259", 130), ("logistic3_n5_N100", 16)])
def test_many_mode_bias_component_either_way(case_name, chains):
    F.test_many_mode_bias_component_either_way(None, case_name, chains)


def test_many_mode_full_gradient_row_tail():
    F.test_many_mode_full_gradient_row_tail_and_blocks(None)


@pytest.mark.parametrize("path", ["tc", "generic", "stream"])
def test_full_gradient_row_ranges_that_start_anywhere(path):
    F.test_full_gradient_over_a_row_range_that_starts_anywhere(None, path)


def test_many_mode_full_gradient_wide_rows():
    """d = 1024: pass 1 spans four TMEM segments per item (the c5 shape of bench.py, at a row count the emulation finishes in seconds)."""
    worst = F.many_mode_check(make_case("logistic1024_n200_N515"), 128, n_opt=3)
    assert worst < 1e-5


# ---- other timing models / schedules, each in a fresh interpreter: "eager" asynchronous engines (copies land at issue), warps
# scheduled in reverse and in random order (inter-warp races are schedule dependent; the default order starts every pass at warp 0)
@pytest.mark.parametrize("env,lazy", [({"SGM_EMU_ASYNC": "eager"}, 0), ({"SGM_EMU_SCHED": "reverse"}, 1), ({"SGM_EMU_SCHED": "20261020"}, 1),
                                      ({"SGM_EMU_CLUSTER": "lockstep", "SGM_EMU_ASYNC": "eager"}, 0)])
def test_kernels_under_other_timing_models(env, lazy):
    code = f"""
from tests.emu import backend
from tests import test_parity_gpu as T, test_fullgrad_tc_gpu as F
with backend.emulated() as lib:
    assert lib.emu_async_is_lazy() == {lazy}
    for case, algo in (("logistic54", "sgld"), ("logistic54_n101", "sgnht"), ("logistic130", "sghmc")):
        T._deterministic_steps(case, algo, True, shared=(algo == "sgld"))
    for case, algo, split, shared in (("logistic54", "sgld", 4, False), ("logistic54_n101", "sghmc", 8, True), ("logistic130", "sgnht", 2, False)):   # clusters: CTA order, lockstep / depth-first
        T._deterministic_steps(case, algo, True, shared=shared, options={{"mode_finder_split": split, "chain_split": split}})
    T._deterministic_steps("mixture", "sghmc", True, shared=False)
    T._deterministic_steps("softmaxnet_small", "sghmc", True, shared=False)
    T._deterministic_steps("softmaxnet", "sghmc", False, shared=False)
    T.test_last_stored_state_is_the_sampler_state(None, "softmaxnet_small", "sghmc")
    F.test_many_mode_full_gradient_matches_oracle(None, "logistic54", 20)
    F.test_many_mode_full_gradient_row_tail_and_blocks(None)
    assert not lib.emu_last_message()
print("variant ok")
"""
    out = backend.run_isolated(code, env=dict(env, SGM_EMU_SMS="1"))
    assert out.returncode == 0 and "variant ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
