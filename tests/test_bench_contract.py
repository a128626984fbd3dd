"""bench.py end to end on the CPU-emulated build: the GPU arm's whole code path (Setup, initSess, sgm_sample e2e leg, device-resident
leg, roofline / e2e / cv_full_gradient / many-mode objects, extra workloads, cpu_baseline) at toy sizes, with torch.cuda stubbed.
bench.py cannot run in this container otherwise, and a Python error in it at round end would cost the round its only numbers.
The VALUES are meaningless here (emulated kernels); the test checks the contract: one JSON line with the keys the driver reads."""
import io
import json
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

from tests.emu import backend

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "roofline", "clocks", "e2e", "gpu_launches", "cpu_baseline"]


def tiny_workload(name, chains):
    import bench
    rs = np.random.RandomState(13)
    if name == "c3":
        N, d = 1500, 54
        X = rs.randn(N, d)
        y = (rs.rand(N) < 0.5).astype(np.float64)
        return dict(desc="toy c3", algo="sgld", cv=True, family="logistic", data={"X": X, "y": y}, params={"bias": np.array(0.0), "beta": np.zeros((d, 1))},
                    stepsize=5e-6, opt_stepsize=5e-6, n_iters_opt=120, n=300, chains=chains or 3, bytes_per_chain_iter=111100, fullgrad_bytes=N * 55 * 4, bound_note="toy")
    if name in ("c1", "c2"):
        w = bench.workload.__wrapped__(name, 4)
        w["data"] = {"X": w["data"]["X"][:2000]}
        return w
    if name == "c4":
        N = 300
        X = rs.rand(N, 20)
        Y = np.eye(4)[rs.randint(0, 4, N)]
        p = {"A": rs.randn(8, 4), "B": rs.randn(20, 8), "a": rs.randn(4), "b": rs.randn(8), "lambdaA": np.array(1.0), "lambdaB": np.array(1.0),
             "lambdaa": np.array(1.0), "lambdab": np.array(1.0)}
        return dict(desc="toy c4", algo="sghmc", cv=False, family="nn", data={"X": X, "y": Y}, params=p, stepsize=1e-6, n=40, chains=2, L=2, alpha=0.01,
                    bytes_per_chain_iter=1000, bound_note="toy")
    raise KeyError(name)


@pytest.mark.parametrize("main", ["c3", "c1"])
def test_bench_prints_one_contract_line(monkeypatch, main):
    import torch
    import bench
    import __graft_entry__
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_: None)
    monkeypatch.setattr(__graft_entry__, "build", lambda *a, **k: None)
    monkeypatch.setattr(bench, "measure_tf32_peak", lambda *a, **k: (900.0, 800.0))
    monkeypatch.setattr(bench, "ONE_CHAIN_ITERS", 150)   # (emulated kernels: keep the one-chain latency leg short)
    monkeypatch.setattr(bench, "QUICK_ITERS", 60)
    monkeypatch.setattr(bench, "probe_cluster_finder", lambda *a, **k: {"ok": True, "checks": []})   # (a child process of the real probe finds no GPU here)
    # the optional legs run in child processes on a GPU box; here (emulated build, loaded into THIS process) they run in place
    monkeypatch.setattr(bench, "run_leg_in_child", lambda leg, args, timeout_s=0: bench.leg_result(leg, bench.workload(args.workload, None), args))
    orig = bench.workload
    tiny_workload_w = tiny_workload
    bench.workload = types.FunctionType(tiny_workload_w.__code__, tiny_workload_w.__globals__)
    bench.workload.__wrapped__ = orig
    monkeypatch.setattr(bench, "oracle_model", lambda w: __import__("oracle.models", fromlist=["x"]).Logistic(54) if w["family"] == "logistic" else
                        (__import__("oracle.models", fromlist=["x"]).GaussMean(2, prior_sd=10.0) if w["family"] == "gauss" else __import__("oracle.models", fromlist=["x"]).Mixture2(2, 10.0)))
    try:
        args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="ours", workload=main, chains=None, iters_per_step=120, cpu_budget=0.5, no_cpu=False, modes=32,
                                     no_modes=False, peer_shards=False, no_extra=False)
        out = io.StringIO()
        with backend.emulated() as lib, redirect_stdout(out):
            bench.run_ours(args, bench.workload(main, None))
            assert not lib.emu_last_message()
    finally:
        bench.workload = orig
    lines = [l for l in out.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.getvalue()[-2000:]
    line = json.loads(lines[0])
    for k in REQUIRED:
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["workload"] == main
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and line["roofline"]["traffic"] is None
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "breakdown_ms"} and line["e2e"]["d2h_bytes_per_step"] > 0
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert line["gpu_launches"] >= 2 and "errors" not in line, line.get("errors")
    assert line["e2e"]["last_draw_finite"] is True
    assert "error" not in line["e2e"]["page_locked_output"] and line["e2e"]["page_locked_output"]["last_draw_finite"] is True, line["e2e"]["page_locked_output"]
    assert set(line["other_workloads"]) == {"c1", "c2", "c4"} - {main}
    assert all("error" not in v for v in line["other_workloads"].values()), line["other_workloads"]
    if main == "c3":
        f = line["mode_finder"]
        assert "error" not in f and f["one_cta_per_mode"]["init_ms"] > 0 and f["default_dispatch"]["us_per_sgd_step_upper_bound"] > 0, f
        assert f["one_chain"]["one_cta_per_chain"]["us_per_step"] > 0 and f["one_chain"]["default_dispatch"]["chain_iterations_per_s"] > 0, f
        assert line["cv_full_gradient"]["GBps"] > 0
        m = line["cv_full_gradient_modes"]
        assert m["TFLOPs"] > 0 and m["peak"] == 800.0 and abs(m["frac_of_3xtf32_ceiling"] - 3 * m["frac"]) < 1e-12


def _bench_worker(rank, world, port, peer_shards, out):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0",
                      SGM_EMU_SHM_IPC="1", SGM_EMU_SMS="2")
    import torch
    import torch.distributed as dist
    from tests.emu import build
    os.environ["SGM_EMU_NCCL"] = build.build_fake_nccl()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import __graft_entry__
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *_: None
    torch.cuda.synchronize = lambda *_: None
    __graft_entry__.build = lambda *a, **k: None
    bench.measure_tf32_peak = lambda *a, **k: (900.0, 800.0)
    # rank 0 reports a failed cluster probe in the --peer-shards run (both ranks must then pin one CTA per chain); the many-mode part passes
    bench.probe_cluster_finder = lambda *a, part=None, **k: {"ok": part == "many_mode" or rank != 0 or not peer_shards, "why": "test: rank 0 reports a failed probe"}
    rs = np.random.RandomState(13)
    N, d = (1200, 130) if peer_shards else (1500, 54)   # a sharded dataset takes the long-row kernel (rows resolved through the shard table)
    w = dict(desc="toy", algo="sgnht" if peer_shards else "sgld", cv=True, family="logistic", data={"X": rs.randn(N, d), "y": (rs.rand(N) < 0.5).astype(np.float64)},
             params={"bias": np.array(0.0), "beta": np.zeros((d, 1))}, stepsize=5e-6, opt_stepsize=5e-6, n_iters_opt=110, n=128, chains=2, a=0.01,
             bytes_per_chain_iter=1000, fullgrad_bytes=N * (d + 1) * 4, bound_note="toy")
    args = types.SimpleNamespace(gpus=world, steps=2, warmup=1, impl="ours", workload="c3", chains=None, iters_per_step=110, cpu_budget=0.5, no_cpu=True, modes=32,
                                 no_modes=False, peer_shards=peer_shards, no_extra=True)
    buf = io.StringIO()
    try:
        with backend.emulated() as lib, redirect_stdout(buf):
            bench.run_ours(args, w)
            msg = lib.emu_last_message()
        out[rank] = (buf.getvalue(), msg.decode() if msg else "")
    except BaseException as e:   # noqa: BLE001
        import traceback
        out[rank] = ("", f"{type(e).__name__}: {e}\n{traceback.format_exc()[-2000:]}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("peer_shards", [False, True])
def test_bench_two_ranks_prints_one_line_on_rank_0(peer_shards):
    """The N > 1 arm (what the driver launches under torchrun for the scaling record) as a 2-process job on the emulated build: gloo
    carries the rendezvous, tests/emu/fake_nccl.c the library's collectives; replicated dataset and --peer-shards."""
    import socket
    import torch.multiprocessing as mp
    from tests.emu import build
    backend.library()
    build.build_fake_nccl()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_bench_worker, args=(2, port, peer_shards, out), nprocs=2, join=True)
    for r in range(2):
        assert r in out and not out[r][1], out.get(r)
    lines0 = [l for l in out[0][0].splitlines() if l.startswith("{")]
    assert len(lines0) == 1 and not [l for l in out[1][0].splitlines() if l.startswith("{")]
    line = json.loads(lines0[0])
    assert "errors" not in line, line["errors"]
    assert line["n_gpus"] == 2 and line["config"]["chains_total"] == 4 and line["scaling"] == "weak"
    assert line["e2e"]["value"] > 0 and line["e2e"]["comm_init_ms_outside"] is not None
    assert line["cv_full_gradient"]["collective_ms"] is not None
    m = line["cv_full_gradient_modes"]
    assert m["modes"] == 32 and m["collective_ms"] is not None and m["TFLOPs_all_ranks"] > 0
    assert ("sharded over 2 ranks" in line["config"]["dataset"]) == peer_shards
    assert line["mode_finder_probe"]["pinned_to_one_cta_per_mode"] == peer_shards


def test_optional_leg_failure_costs_only_its_own_entry():
    """A leg's child process that cannot produce a result (here: no CUDA device in this container) comes back as an error object."""
    import bench
    args = types.SimpleNamespace(workload="c1", modes=16, chains=None)
    r = bench.run_leg_in_child("quick:c2", args, timeout_s=300)
    assert set(r) >= {"error"} and "child exited" in r["error"], r
    assert "CUDA device" in r.get("stderr_tail", "")


def test_smoke_runs_on_the_emulated_build(monkeypatch, capsys):
    """__graft_entry__.smoke(): the TMA-gather logistic step, the tcgen05 softmax-net steps and the tcgen05 many-mode full gradient, each
    checked against the oracle."""
    import __graft_entry__
    monkeypatch.setattr(__graft_entry__, "build", lambda *a, **k: None)
    monkeypatch.setattr(__graft_entry__, "SMOKE_CHILD_PROBE", False)   # (the child process would load the product library: no GPU here)
    with backend.emulated() as lib:
        lib.emu_note_count.restype = __import__("ctypes").c_longlong
        lib.emu_note_count.argtypes = [__import__("ctypes").c_char_p]
        before = lib.emu_note_count(b"cvg:pass1"), lib.emu_note_count(b"logistic_staged:gather4")
        __graft_entry__.smoke()
        assert lib.emu_note_count(b"cvg:pass1") > before[0] and lib.emu_note_count(b"logistic_staged:gather4") > before[1]
    assert capsys.readouterr().out.count("smoke ok") == 3
