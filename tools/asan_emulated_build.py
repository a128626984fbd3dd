#!/usr/bin/env python
"""TEST INFRASTRUCTURE: the host side of the C-ABI library (csrc/api.cu: handle lifecycle, upload / copy-out staging, launch plans, the
state blob) and the kernel sources under AddressSanitizer, on the CPU-emulated build (tests/emu).

    python tools/asan_emulated_build.py            # builds tests/emu/_build_asan, runs a cross-section of the test bodies; exit 0 = clean

The kernel sources and api.cu are compiled with -fsanitize=address; the emulator core (emu.cpp: hand-written context switches between
fiber stacks, which ASan's no-return handling cannot follow) is not.  libasan is LD_PRELOADed into a child interpreter.  Not part of
`pytest -m "not gpu"` (one extra minute of g++ and two of running); run it after touching host-side buffer code."""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BODY = r'''
import sys
sys.path.insert(0, %r)
from tests.emu import backend
backend.VARIANTS["asan"] = ()            # built by the parent
from tests import test_parity_gpu as T, test_fullgrad_tc_gpu as F, test_launch_plan as LP, fuzz_emu
with backend.emulated("asan") as lib:
    T._deterministic_steps("logistic54", "sgld", True, shared=True, options={"mode_finder_split": 8})
    T._deterministic_steps("logistic54_n101", "sghmc", True, shared=False)
    T._deterministic_steps("logistic54_n101", "sghmc", True, shared=False, options={"chain_split": 4, "mode_finder_split": 4})
    T._deterministic_steps("logistic130", "sgnht", True, shared=True, options={"chain_split": 8})
    T.test_cluster_per_chain_walks_the_same_chain(None, "logistic54", "sgld", True, True)
    T.test_sampling_into_a_page_locked_array_equals_the_staged_copy(None, "logistic54", "sgld", True, 230, 3)
    T._deterministic_steps("logistic1024_n1024_N3000", "sgnht", True, shared=True, options={"mode_finder_split": 8}, n_opt=4)
    T._deterministic_steps("mixture", "sgnht", True, shared=False)
    T._deterministic_steps("gauss3d_fullcov", "sghmc", True, shared=False)
    T._deterministic_steps("softmaxnet_small", "sghmc", True, shared=False)
    F.test_many_mode_full_gradient_matches_oracle(None, "logistic54", 20)
    F.test_many_mode_full_gradient_row_tail_and_blocks(None)
    F.test_full_gradient_over_a_row_range_that_starts_anywhere(None, "tc")
    F.test_many_mode_bias_component_either_way(None, "logistic128_n20_N300", 17)
    T.test_saved_state_resumes_the_run_bit_for_bit(None, "logistic54", "sgnht", True, True)
    T.test_chain_storage_layout_and_shapes(None)
    T.test_heldout_loglik_matches_oracle(None, "logistic54")
    T.test_dataset_layouts_upload_identically(None, "logistic54")
    T.test_dataset_layouts_upload_identically(None, "softmaxnet_small")
    T.test_call_sequences_walk_one_rng_stream(None, "mixture", "sghmc", False)
    T.test_est_logpost_matches_oracle(None, "gauss2d")
    T.test_same_seed_reproduces(None)
    T.test_streamed_sampling_reports_divergence(None)
    for case, algo, mode in (("logistic54", "sgld", "cv_shared"), ("gauss2d", "sghmc", "plain"), ("softmaxnet_small", "sgnht", "cv")):
        fuzz_emu.run_case("stream", case, algo, mode)
    for name, cfg in LP.CONFIGS.items():
        for world in (1, 3, 8):
            assert LP.plan(lib, world=world, rank=world - 1, **cfg)["violations"] == []
    msg = lib.emu_last_message()
    assert not msg, msg
print("ASAN_RUN_OK")
''' % ROOT


def main() -> int:
    from tests.emu import build as B
    flags = ("-fsanitize=address", "-fno-omit-frame-pointer", "-O1")
    lib = B.build(variant="asan", defs=flags)
    out_dir = os.path.dirname(lib)
    cxx = os.environ.get("CXX", "g++")
    # the emulator core without instrumentation, then relink
    plain = [f for f in B.CXXFLAGS]
    subprocess.check_call([cxx, *plain, "-c", os.path.join(B.HERE, "emu.cpp"), "-o", os.path.join(out_dir, "emu.o")], cwd=ROOT)
    objs = [os.path.join(out_dir, n) for n in ("api.o", "softmax_net.o", "emu.o", "selftest.o")]
    subprocess.check_call([cxx, "-shared", "-Wl,-Bsymbolic", "-o", lib, *objs, "-ldl"], cwd=ROOT)
    asan = ""
    for cc in (cxx, "g++", "gcc", "/usr/bin/g++"):      # a compiler that knows where its libasan lives (a bare name back = it does not)
        try:
            cand = subprocess.check_output([cc, "-print-file-name=libasan.so"], text=True).strip()
        except (OSError, subprocess.CalledProcessError):
            continue
        if os.path.isabs(cand) and os.path.exists(cand):
            asan = cand
            break
    if not asan:
        print("asan: no libasan.so found next to any compiler -- nothing was run")
        return 2
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    p = subprocess.run([sys.executable, "-c", BODY], cwd=ROOT, env=env, capture_output=True, text=True)
    ok = p.returncode == 0 and "ASAN_RUN_OK" in p.stdout
    if not ok:
        print(p.stdout[-2000:], p.stderr[-6000:])
    print("asan:", "clean" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
