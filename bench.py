#!/usr/bin/env python
"""Benchmark of the sgmcmc sampler hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

One "step" = one launch of the persistent chain kernel: ``--iters-per-step``
sgmcmcSteps for every chain of this rank (minibatch sampling + gather + gradient +
fused update + device-side chain storage).  The metric is BASELINE.json's: SGMCMC
iterations/s counted as chains x iters (whole job, all ranks), plus the
control-variate full-data gradient GB/s, each against the HBM roofline.

Workloads (BASELINE.json configs; synthetic data of the named shapes):
    c1  mvGauss SGLD            N=1e4  d=2   n=100
    c2  gaussMixture SGHMC L=5  N=1e4  d=2   n=100
    c3  logistic SGLDCV         N=581012 d=54 n=500      <- default (metric names both halves on it)
    c4  nn 784-100-10 SGHMC L=5 N=60000 n=100
    c5  SGNHTCV logistic        N=1.25e7/GPU d=1024 n=1024, 128 chains/GPU (1024 at 8 GPUs), data generated on device
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sgmcmc_iterations_per_s_chains_x_iters"
UNIT = "chain-iterations/s"


# --------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------
def workload(name: str, chains: int | None):
    """Returns dict(desc, setup_kwargs..., bytes_per_chain_iter, ...). Host data is synthetic, seed 13."""
    rs = np.random.RandomState(13)
    if name == "c1":
        X = rs.randn(10 ** 4, 2)
        return dict(desc="mvGauss vignette: SGLD, 2-d Gaussian mean, N=1e4, n=100", algo="sgld", cv=False, family="gauss",
                    data={"X": X}, params={"theta": np.zeros(2)}, stepsize=1e-5, n=100, chains=chains or 148 * 64,
                    bytes_per_chain_iter=824, bound_note="dataset (80 KB) is L1/L2 resident: not HBM-bound")
    if name == "c2":
        z = rs.rand(10 ** 4) < 0.5
        X = rs.randn(10 ** 4, 2) + np.where(z[:, None], 0.0, 0.1)
        return dict(desc="gaussMixture vignette: SGHMC L=5, 2-component 2-d mixture, N=1e4, n=100", algo="sghmc", cv=False,
                    family="mixture", data={"X": X}, params={"theta1": np.zeros(2), "theta2": np.full(2, 0.1)},
                    stepsize=2e-5, n=100, chains=chains or 148 * 64, L=5, alpha=0.01,
                    bytes_per_chain_iter=4352, bound_note="dataset (80 KB) is L1/L2 resident: not HBM-bound")
    if name == "c3":
        N, d = 581012, 54
        X = rs.randn(N, d)
        beta = rs.randn(d) * 0.5
        y = (rs.rand(N) < 1.0 / (1.0 + np.exp(-(0.1 + X @ beta / np.sqrt(d))))).astype(np.float64)
        return dict(desc="logisticRegression: SGLDCV, covertype-shaped N=581012 d=54, n=500, CV full gradient at mode",
                    algo="sgld", cv=True, family="logistic", data={"X": X, "y": y},
                    params={"bias": np.array(0.0), "beta": np.zeros((d, 1))}, stepsize=5e-6, opt_stepsize=5e-6,
                    n_iters_opt=10 ** 4, n=500, chains=chains or 148 * 8,
                    bytes_per_chain_iter=111100, fullgrad_bytes=N * 55 * 4,
                    bound_note="dataset is 130 MB (~L2 size): gathers are served partly from L2")
    if name == "c5":
        d = 1024
        return dict(desc="scale-out: SGNHTCV logistic, d=1024, n=1024, 148 chains/GPU (one per SM; 8 GPUs = 1184 >= the config's 1024), N=1.25e7 rows/GPU generated on device",
                    algo="sgnht", cv=True, family="logistic", data=None, synth=dict(rows=int(os.environ.get("SGM_C5_ROWS", 12500000)), dim=d, seed=13),
                    params={"bias": np.array(0.0), "beta": np.zeros((d, 1))}, stepsize=1e-8, opt_stepsize=1e-7,
                    n_iters_opt=200, n=1024, chains=chains or 148, a=0.01,
                    bytes_per_chain_iter=1024 * 1025 * 4 + (4 * 1025 + 4) * 4 + 2 * 1025 * 4,
                    fullgrad_bytes=int(os.environ.get("SGM_C5_ROWS", 12500000)) * 1025 * 4,
                    bound_note="51 GB dataset >> L2: HBM-bound gather")
    if name == "c4":
        N = 60000
        X = rs.rand(N, 784)
        Y = np.eye(10)[rs.randint(0, 10, N)]
        p = {"A": rs.randn(100, 10), "B": rs.randn(784, 100), "a": rs.randn(10), "b": rs.randn(100),
             "lambdaA": np.array(rs.gamma(1.0)), "lambdaB": np.array(rs.gamma(1.0)),
             "lambdaa": np.array(rs.gamma(1.0)), "lambdab": np.array(rs.gamma(1.0))}
        return dict(desc="nn vignette: SGHMC L=5, 784-100-10 softmax net, N=60000, n=100", algo="sghmc", cv=False, family="nn",
                    data={"X": X, "y": Y}, params=p, stepsize=1e-6, n=100, chains=chains or 148, L=5, alpha=0.01,
                    bytes_per_chain_iter=int(8.59e6), bound_note="per-chain weights (318 KB) stream from L2/HBM")
    raise SystemExit(f"unknown workload {name}")


def model_objects(family):
    import sgmcmc_b200 as S
    return {"gauss": (S.gaussMeanLogLik(), S.normalPrior(10.0)), "mixture": (S.mixture2LogLik(), S.normalPrior(10.0)),
            "logistic": (S.logisticLogLik(), S.laplacePrior()), "nn": (S.softmaxNetLogLik(), S.gammaPrecisionPrior())}[family]


def oracle_model(w):
    from oracle import models as OM
    if w["family"] == "gauss":
        return OM.GaussMean(2, prior_sd=10.0)
    if w["family"] == "mixture":
        return OM.Mixture2(2, 10.0)
    if w["family"] == "logistic":
        return OM.Logistic(w["params"]["beta"].shape[0])
    return OM.SoftmaxNet()


PIN_ONE_CTA_FINDER = False   # set when the cluster mode finder failed its probe (probe_cluster_finder below)


def make_sampler(w, rank, world, chains_local, seed=13, shared=True, synth=None, n_iters_opt=None):
    sg = _make_sampler(w, rank, world, chains_local, seed, shared, synth, n_iters_opt)
    if PIN_ONE_CTA_FINDER and w["family"] == "logistic":
        from sgmcmc_b200 import validation as V
        V.setOption(sg, "mode_finder_split", 1)
        V.setOption(sg, "chain_split", 1)
    return sg


def probe_cluster_finder(timeout_s=600, part=None):
    """tools/probe_cluster_finder.py in a process of its own: the cluster mode finder (default dispatch for a shared mode) had never run
    on hardware when this round ended.  Returns the probe's JSON object; {"ok": False, ...} makes this run keep the finder on one CTA
    per mode, so a fault in the new kernel costs the `mode_finder` entry of the line and not the line."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "probe_cluster_finder.py")] + (["--part", part] if part else []),
                           capture_output=True, text=True, timeout=timeout_s, env=env, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"ok": False, "why": f"no answer within {timeout_s} s (killed)"}
    for ln in reversed(p.stdout.strip().splitlines()):
        if ln.startswith("{"):
            try:
                r = json.loads(ln)
                r["ok"] = bool(r.get("ok")) and p.returncode == 0
                return r
            except ValueError:
                break
    return {"ok": False, "why": f"probe exited {p.returncode} without a result", "stderr_tail": p.stderr[-400:]}


def _make_sampler(w, rank, world, chains_local, seed=13, shared=True, synth=None, n_iters_opt=None):
    import sgmcmc_b200 as S
    ll, lp = model_objects(w["family"])
    kw = dict(logPrior=lp, minibatchSize=w["n"], seed=seed, nChains=chains_local, chainOffset=rank * chains_local,
              sharedMode=shared, device=int(os.environ.get("LOCAL_RANK", 0)))
    if synth or w.get("synth"):
        kw["synth"] = synth or w["synth"]
    data = w["data"] if w["data"] is not None else {}
    algo, cv = w["algo"], w["cv"]
    if algo == "sghmc":
        kw.update(alpha=w.get("alpha", 0.01), L=w.get("L", 5))
    if algo == "sgnht":
        kw.update(a=w.get("a", 0.01))
    if cv:
        fn = {"sgld": S.sgldcvSetup, "sghmc": S.sghmccvSetup, "sgnht": S.sgnhtcvSetup}[algo]
        return fn(ll, data, w["params"], w["stepsize"], w["opt_stepsize"], nItersOpt=n_iters_opt or w["n_iters_opt"], verbose=False, **kw)
    fn = {"sgld": S.sgldSetup, "sghmc": S.sghmcSetup, "sgnht": S.sgnhtSetup}[algo]
    return fn(ll, data, w["params"], w["stepsize"], **kw)


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop_ev, self.thread = index, [], threading.Event(), None

    def _loop(self):
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.05)

    def __enter__(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_ev.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons),
                "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1500.0, "fallback (B200_PROFILING.md dense bf16)"


def measure_tf32_peak(seconds=1.5):
    """cuBLAS TF32 (fp32 inputs, allow_tf32) 8192^3 on this GPU, now: (burst TFLOP/s = best of 10, sustained = back to back for `seconds`).
    MEASURED_PEAKS.json only has bf16; the many-mode control-variate pass computes in 3xTF32, so its ceiling is THIS number / 3."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda", dtype=torch.float32)
        b = torch.randn(n, n, device="cuda", dtype=torch.float32)
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, t0 = 0, time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(10):
                torch.matmul(a, b)
            reps += 10
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        sustained = reps * flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b
        return best, sustained
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def max_over_ranks(vals):
    """Element-wise maximum over the ranks of a short list of floats (timings): one all-reduce on the backend's own device."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def comm_init(sg, rank, world):
    from sgmcmc_b200 import distributed as D
    D.comm_init(sg, rank, world)


def attach_peer_shards(sg, rank, world):
    from sgmcmc_b200 import distributed as D
    D.attach_peer_shards(sg, rank, world)


def many_mode_full_gradient(w, rank, world, modes):
    """Control-variate full-data gradient with one mode per chain: two tensor-core contractions (csrc/logistic_cv_gemm.cuh).
    A separate sampler (sharedMode=False) on a bounded row count; returns the cv_full_gradient_modes object."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import validation as V
    synth = None
    if w.get("synth"):   # c5: the rank's whole row shard unless SGM_MODES_ROWS bounds it
        synth = dict(w["synth"], rows=int(os.environ.get("SGM_MODES_ROWS", w["synth"]["rows"])))
    # N > 1: the chains (and their modes) are split over the ranks, every rank contracts its 1/N of the rows against ALL modes
    # (all-gather of the modes, reduce-scatter of the sums): strong scaling of a fixed 4 N d M
    sg = make_sampler(w, rank, world, max(1, modes // world), shared=False, synth=synth, n_iters_opt=20)
    if world > 1:
        comm_init(sg, rank, world)
    S.initSess(sg, verbose=False)
    flops = V.lastFullGradFlops(sg)
    reduce_ms, gather_ms = V.lastCollectiveTiming(sg)
    reps = 3 if flops > 2e14 else 20
    ms, _ = V.benchFullGradient(sg, reps)
    bf16_peak, bf16_src = measured_tensor_peak()
    tf32_burst, tf32_sustained = measure_tf32_peak()
    rows = (synth["rows"] if synth else np.asarray(w["data"]["X"]).shape[0]) // world
    d = w["params"]["beta"].shape[0]
    tf = flops / (ms * 1e-3) / 1e12 if ms > 0 and flops > 0 else None
    del sg
    # ONE stated denominator: the TF32 throughput cuBLAS reaches on this GPU now (sustained).  Three TF32 MMAs are spent per float32
    # product, so the ceiling of this arithmetic is that / 3; the bf16 figure of MEASURED_PEAKS.json is printed for reference only.
    return {"modes": max(1, modes // world) * world, "rows_per_rank": rows, "d": d, "ms": ms, "flops": flops, "TFLOPs": tf, "TFLOPs_all_ranks": (tf * world) if tf else None,
            "bound": "tensor", "peak": tf32_sustained, "peak_source": "cuBLAS TF32 8192^3 measured in this run (sustained); burst %.1f" % tf32_burst,
            "frac": (tf / tf32_sustained) if tf else None,
            "frac_of_3xtf32_ceiling": (tf / (tf32_sustained / 3.0)) if tf else None,
            "tensor_pipe_TFLOPs": (3.0 * tf) if tf else None,
            "bf16_peak_reference": bf16_peak, "bf16_peak_source": bf16_src,
            "collective_ms": {"all_gather_modes": gather_ms, "reduce_scatter_sums": reduce_ms} if world > 1 else None,
            "arithmetic": "3xTF32 (hi*hi + hi*lo + lo*hi) on tcgen05, float32-level accuracy: three TF32 MMAs per float32 product, "
                          "so the ceiling of the ALGORITHMIC rate is the TF32 peak / 3; tensor_pipe_TFLOPs = 3 x the algorithmic rate",
            "passes": "R = y - sigmoid(X theta_hat^T + bias); G = X^T R (K split over row ranges), per row block",
            "sample": "whole dataset" if not synth or synth["rows"] == w["synth"]["rows"] else "bounded row sample of the c5 shape"}


# --------------------------------------------------------------------------------------
# CPU arm (reference-structured proxy; see oracle/cpu_baseline.py)
# --------------------------------------------------------------------------------------
def cpu_proxy(w, budget_s, threads):
    import torch
    from oracle.cpu_baseline import ReferenceStructuredSampler, time_full_gradient, time_iterations
    torch.set_num_threads(threads)
    if w["data"] is None:  # c5: bounded host sample of the device-generated shape
        rs = np.random.RandomState(13)
        d = w["synth"]["dim"]
        N = 20000
        X = rs.randn(N, d)
        y = (rs.rand(N) < 1 / (1 + np.exp(-(X @ (rs.randn(d) / np.sqrt(d)))))).astype(np.float64)
        data, sample_note = {"X": X, "y": y}, f"N reduced to {N} rows on the host (the 51 GB device dataset cannot be held by the CPU arm); "
    else:
        data, sample_note = w["data"], ""
    s = ReferenceStructuredSampler(oracle_model(w), data, w["params"], w["algo"], w["stepsize"], cv=w["cv"],
                                   alpha=w.get("alpha", 0.01), a=w.get("a", 0.01), L=w.get("L", 5),
                                   opt_stepsize=w.get("opt_stepsize", 0.0), n=w["n"])
    fg = None
    if w["cv"]:
        s.find_mode(50)  # enough to exercise the path; the timed part is the sampling loop + the full gradient
        fg = time_full_gradient(oracle_model(w), data, w["params"])
    ips, iters, dt = time_iterations(s, budget_s)
    cpu_proxy.last_sampler = s
    return ips, f"{sample_note}1 chain, {iters} sgmcmcSteps in {dt:.1f} s, torch CPU float32 graph replica, {threads} thread(s)", fg, data


def cpu_chain_leg(w, budget_s, steps, warmup):
    """Child of run_reference: ONE chain of the reference-structured proxy on one thread; W short warm-up samples, then K bounded samples
    (one per step of the arm); prints the K rates."""
    import torch
    from oracle.cpu_baseline import time_iterations
    torch.set_num_threads(1)
    ips, note, fg, _ = cpu_proxy(w, 1.0, 1)              # builds the sampler, finds a mode, times the full gradient; first (warm-up) sample
    s = cpu_proxy.last_sampler
    for _ in range(max(0, warmup - 1)):
        time_iterations(s, 1.0)
    rates, iters_tot, dt_tot = [], 0, 0.0
    for _ in range(steps):
        r, iters, dt = time_iterations(s, budget_s)
        rates.append(r); iters_tot += iters; dt_tot += dt
    return {"rates": rates, "note": f"1 chain, {iters_tot} sgmcmcSteps in {dt_tot:.1f} s over {steps} samples, torch CPU float32 graph replica, 1 thread", "full_gradient_s": fg}


def run_reference(args, w):
    """--impl reference: the reference's own structure on the host cores (the real R+TF stack is absent).  A chain of the reference is
    sequential and its per-op work is far too small for intra-op threads to help (torch with 8 threads: 27 iterations/s against 634 on one
    thread for c3, measured on the development host), so "all the host threads it can use" means what a user of the reference would do
    with a many-core host: one independent chain per core (one R process each).  P single-thread chain processes run side by side; each
    step of the arm is one bounded sample in every process at the same time; value = the sum of their rates, mean over the K timed steps."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))        # <= 32 processes: each holds its own copy of the dataset (0.7 GB for c3)
    budget = max(2.0, min(15.0, 100.0 / max(1, args.steps)))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env.update(OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
    cmd = [sys.executable, os.path.abspath(__file__), "--leg", "cpu_chain", "--workload", args.workload, "--gpus", "1", "--cpu-budget", str(budget),
           "--steps", str(args.steps), "--warmup", str(args.warmup)]
    children = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env, cwd=ROOT) for _ in range(procs)]
    per_step, note, failed = np.zeros(args.steps), "", 0
    for ch in children:
        try:
            out, _ = ch.communicate(timeout=budget * args.steps * 3 + 600)
            r = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
            per_step += np.asarray(r["rates"], dtype=np.float64); note = r["note"]
        except Exception:   # noqa: BLE001
            ch.kill(); failed += 1
    if failed == procs:
        print(json.dumps({"impl": "reference", "unavailable": "no CPU chain process produced a rate"}))
        return
    val = float(np.mean(per_step))
    sample = f"{procs - failed} independent single-thread chains at once, each: {note}" + (f"; {failed} process(es) failed" if failed else "")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * budget, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "desc": w["desc"], "chains": procs - failed},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs - failed, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "TF-proxy (torch CPU) -- not the reference binary: R and TensorFlow are absent from this image"}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def device_leg(sg, lib, V, _lib, I, steps, warmup, barrier):
    """W warm-up launches, then exactly K timed ones (device-resident chain): returns (wall seconds, per-launch kernel ms list, launches)."""
    for _ in range(max(3, warmup)):
        _lib.check(lib.sgm_reset_chain(sg.handle))
        _lib.check(lib.sgm_run(sg.handle, I, 1))
    kernel_ms, launches = [], 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.sgm_reset_chain(sg.handle)
        _lib.check(lib.sgm_run(sg.handle, I, 1))
        ms, nl = V.lastTiming(sg)
        kernel_ms.append(ms)
        launches += nl
    barrier()
    return time.perf_counter() - t0, kernel_ms, launches


QUICK_ITERS = 550   # sgmcmcSteps per launch of the compact c1 / c2 measurements


def quick_workload(name, steps=5, iters=None):
    """A compact device-resident measurement of another BASELINE config (N = 1 only), for the same JSON line."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import _lib
    from sgmcmc_b200 import validation as V
    import torch
    w = workload(name, None)
    I = iters or (QUICK_ITERS if name in ("c1", "c2") else 20)
    lib = _lib.load()
    sg = make_sampler(w, 0, 1, w["chains"])
    S.initSess(sg, verbose=False)
    _lib.check(lib.sgm_reserve_chain(sg.handle, I))

    def barrier():
        torch.cuda.synchronize()
    dt, kms, launches = device_leg(sg, lib, V, _lib, I, steps, 3, barrier)
    peak, _ = measured_peaks()
    k_ms = float(np.mean(kms))
    achieved = w["chains"] * I * w["bytes_per_chain_iter"] / (k_ms * 1e-3) / 1e9
    del sg
    return {"desc": w["desc"], "value": w["chains"] * I * steps / dt, "unit": UNIT, "chains": w["chains"], "iters_per_step": I, "steps": steps,
            "kernel_ms_per_launch": k_ms, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "bytes_per_chain_iter": w["bytes_per_chain_iter"], "note": w["bound_note"]}}



ONE_CHAIN_ITERS = 2000   # sgmcmcSteps of the one-chain latency measurement (one launch)


def mode_finder_ab(w, n_iters_opt=None):
    """Few chains / one shared mode: the thread-block cluster per chain (csrc/logistic_split.cuh, the default dispatch there) against one
    CTA per chain, each on a fresh sampler.
    mode finder (getMode: nItersOpt sequential SGD steps): device time of initSess, which is the finder plus one full-data gradient and
    a log-posterior check every 100 steps -- an upper bound on the time per SGD step;
    one chain (the reference's own way of running): device time of 2000 sgmcmcSteps of ONE chain in one launch."""
    import sgmcmc_b200 as S
    from sgmcmc_b200 import _lib
    from sgmcmc_b200 import validation as V
    lib = _lib.load()
    n_opt = n_iters_opt or min(int(w.get("n_iters_opt", 2000)), 2000)
    out = {"n_iters_opt": n_opt, "modes": 1}
    for name, split in (("one_cta_per_mode", 1), ("default_dispatch", 0)):
        sg = make_sampler(w, 0, 1, w["chains"], n_iters_opt=n_opt)
        V.setOption(sg, "mode_finder_split", split)
        S.initSess(sg, verbose=False)      # warm-up (first launch of every kernel)
        S.initSess(sg, verbose=False)
        ms, launches = V.lastTiming(sg)
        out[name] = {"init_ms": ms, "us_per_sgd_step_upper_bound": ms * 1e3 / n_opt, "launches": launches}
        del sg
    iters = ONE_CHAIN_ITERS
    one = {"iters": iters, "chains": 1}
    for name, split in (("one_cta_per_chain", 1), ("default_dispatch", 0)):
        sg = make_sampler(w, 0, 1, 1, n_iters_opt=min(n_opt, 200))
        V.setOption(sg, "chain_split", split)
        S.initSess(sg, verbose=False)
        _lib.check(lib.sgm_run(sg.handle, iters, 0))
        _lib.check(lib.sgm_run(sg.handle, iters, 0))
        ms, _ = V.lastTiming(sg)
        one[name] = {"ms": ms, "us_per_step": ms * 1e3 / iters, "chain_iterations_per_s": iters / (ms * 1e-3)}
        del sg
    out["one_chain"] = one
    return out


def run_leg_in_child(leg, args, timeout_s=300):
    """Runs one optional leg (`--leg NAME`) of this script in its own process (N = 1 only) and returns its JSON object.  A leg that
    faults on the device (a sticky CUDA error ends every later call of the process that raised it) or hangs then costs its own entry,
    not the main line.  The child measures with CUDA events exactly as the parent would; process start-up is outside its timed regions."""
    cmd = [sys.executable, os.path.abspath(__file__), "--leg", leg, "--workload", args.workload, "--modes", str(args.modes), "--gpus", "1"]
    if args.chains:
        cmd += ["--chains", str(args.chains)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"error": f"leg {leg}: no result within {timeout_s} s (child process killed)"}
    for ln in reversed(p.stdout.strip().splitlines()):
        if ln.startswith("{"):
            try:
                return json.loads(ln)
            except ValueError:
                break
    return {"error": f"leg {leg}: child exited {p.returncode} without a result", "stderr_tail": p.stderr[-400:]}


def leg_result(leg, w, args):
    """One optional measurement by name (the body of a `--leg` child)."""
    try:
        if leg == "modes":
            return many_mode_full_gradient(w, 0, 1, args.modes)
        if leg == "finder":
            return mode_finder_ab(w)
        if leg.startswith("quick:"):
            return quick_workload(leg.split(":", 1)[1])
        return {"error": f"unknown leg {leg}"}
    except Exception as e:   # noqa: BLE001
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}


def run_leg(args, w):
    """Child side of run_leg_in_child: one JSON object on stdout."""
    if args.leg == "cpu_chain":     # the reference arm's worker: CPU only
        print(json.dumps(cpu_chain_leg(w, args.cpu_budget, args.steps, args.warmup)))
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(0)
    print(json.dumps(leg_result(args.leg, w, args)))


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    own_group = world > 1 and not dist.is_initialized()   # (a test on the emulated build arrives with a gloo group already up)
    if own_group:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__
    if local == 0:
        __graft_entry__.build()   # no-op when the in-tree .so is current
    if world > 1:
        dist.barrier()
    import sgmcmc_b200 as S
    from sgmcmc_b200 import _lib
    from sgmcmc_b200 import validation as V
    lib = _lib.load()
    chains_local = w["chains"]
    I = args.iters_per_step
    total_chains = chains_local * world
    peak, peak_src = measured_peaks()
    errors = {}
    # ---- the cluster mode finder, tried in a child process before this process depends on it (rank 0 asks, every rank follows)
    global PIN_ONE_CTA_FINDER
    finder_probe = None
    finder_probe_all_ranks_say_many_mode_bad = False
    if w["family"] == "logistic":
        mm_probe = None
        if rank == 0:
            finder_probe = probe_cluster_finder()
            if world > 1 and w["cv"] and not args.no_modes:      # (at N = 1 the many-mode leg runs in a child of its own anyway)
                mm_probe = probe_cluster_finder(part="many_mode")
                finder_probe["many_mode"] = mm_probe
        bad = 0.0 if (finder_probe is None or finder_probe.get("ok")) else 1.0
        mm_bad = 0.0 if (mm_probe is None or mm_probe.get("ok")) else 1.0
        if world > 1:
            bad, mm_bad = max_over_ranks([bad, mm_bad])
        PIN_ONE_CTA_FINDER = bad > 0
        finder_probe_all_ranks_say_many_mode_bad = mm_bad > 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- process-level one-time cost, outside the end-to-end region and reported apart: the job's NCCL communicator (like the CUDA
    # context, it is created once per process; every sampler of the job then shares it -- sgmcmc_b200/distributed.py)
    comm_ms = None
    if world > 1 and w["cv"]:
        t0 = time.perf_counter()
        tiny = S.sgldcvSetup(S.gaussMeanLogLik(), {"X": np.zeros((64, 2))}, {"theta": np.zeros(2)}, 1e-4, 1e-4, logPrior=S.normalPrior(10.0), minibatchSize=8,
                             nItersOpt=1, verbose=False, seed=1, sharedMode=True, device=local)
        comm_init(tiny, rank, world)
        barrier()
        comm_ms = 1000.0 * (time.perf_counter() - t0)

    sharded = bool(args.peer_shards and world > 1 and w["family"] == "logistic")
    synth_main = dict(w["synth"], seed=w["synth"].get("seed", 13) + rank) if (sharded and w.get("synth")) else None   # different rows per rank

    def new_sampler():
        sg = make_sampler(w, rank, world, chains_local, synth=synth_main)
        if sharded:
            attach_peer_shards(sg, rank, world)
        if world > 1 and w["cv"]:
            comm_init(sg, rank, world)
        return sg

    def drop(sg):
        if sharded:   # unmap on every rank before any owner frees its shard
            _lib.check(lib.sgm_detach_peer_shards(sg.handle))
            barrier()

    # ---- device-resident leg FIRST (the line's `value` and roofline): W warm-up steps, then exactly K timed steps
    sg = new_sampler()
    sess = S.initSess(sg, verbose=False)
    init_ms, init_launches = V.lastTiming(sg)
    fg_ms, fg_bytes = V.lastFullGradTiming(sg) if w["cv"] else (0.0, 0)
    reduce_ms, _ = V.lastCollectiveTiming(sg) if w["cv"] else (0.0, 0.0)
    _lib.check(lib.sgm_reserve_chain(sg.handle, I))
    with ClockSampler(local) as clk:
        dt, kernel_ms, launches = device_leg(sg, lib, V, _lib, I, args.steps, args.warmup, barrier)
    if world > 1:
        dt = max_over_ranks([dt])[0]
    clocks = clk.summary()
    value = total_chains * I * args.steps / dt
    k_ms = float(np.mean(kernel_ms))
    achieved = chains_local * I * w["bytes_per_chain_iter"] / (k_ms * 1e-3) / 1e9
    dataset_mb = {"c1": 0.08, "c2": 0.08, "c3": 130.1, "c4": 191.0, "c5": 51400.0}.get(args.workload)
    if w.get("synth"):   # generated on the device: rows x padded row (d + y, a multiple of 4 floats) x 4 bytes
        dataset_mb = w["synth"]["rows"] * 4 * (w["synth"]["dim"] // 4 + 1) * 4 / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "chains_per_gpu": chains_local, "chains_total": total_chains,
                   "iters_per_step": I, "minibatch": w["n"], "store": "device-resident chain, every iteration",
                   "l2": ("dataset %.0f MB >> the 126 MB L2: gathers come from HBM" % dataset_mb) if (args.workload == "c5" and dataset_mb > 4 * 126) else
                         ("dataset %.0f MB (reduced from the config's 51,400 MB): NOT larger than the 126 MB L2, the gathers are served from L2" % dataset_mb) if args.workload == "c5" else
                         ("dataset %.1f MB is about the size of the 126 MB L2: the gathers are served partly from L2 (the figure is algorithmic bytes / time, "
                          "NOT an HBM-traffic claim; the per-chain state is re-read from L2 too)" % dataset_mb) if args.workload in ("c3", "c4") else
                         "dataset 80 KB lives in shared memory / L1: not an HBM-bound case (stated)",
                   "sampling": "device Philox, without replacement", "shared_mode": bool(w["cv"]),
                   "dataset": (f"sharded over {world} ranks, peers mapped by CUDA IPC: chains sample from {world}x the rows, "
                               f"{world - 1}/{world} of the gathered bytes arrive over NVLink") if sharded else "replicated per GPU"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "traffic_note": "no ncu capture of this commit exists (gpurun was closed for the whole round): measured DRAM traffic is not claimed",
                     "algorithmic_bytes_per_launch": chains_local * I * w["bytes_per_chain_iter"],
                     "kernel": "persistent chain kernel (gather + gradient + update + store)",
                     "kernel_ms_per_launch": k_ms, "bytes_per_chain_iter": w["bytes_per_chain_iter"], "peak_source": peak_src,
                     "note": w["bound_note"]},
        "clocks": clocks,
        "e2e": None,
        "gpu_launches": launches,
    }
    try:
        if w["cv"]:
            if w["family"] == "logistic":   # steady-state figure: 20 launches back to back (the one inside initSess sits between launch gaps)
                fg_ms, fg_bytes = V.benchFullGradient(sg, 20)
            line["cv_full_gradient"] = {"ms": fg_ms, "bytes": fg_bytes, "GBps": (fg_bytes / (fg_ms * 1e-3) / 1e9) if fg_ms > 0 else None,
                                        "frac_of_hbm_peak": (fg_bytes / (fg_ms * 1e-3) / 1e9 / peak) if fg_ms > 0 else None,
                                        "rows_per_rank": fg_bytes // (16 * (w["params"]["beta"].shape[0] // 4 + 1)) if w["family"] == "logistic" else None,
                                        "init_ms_total": init_ms,
                                        "collective_ms": reduce_ms if world > 1 else None,
                                        "note": "one mode (shared by all chains): HBM-bound stream, not a GEMM; bytes = rows x row stride; collective_ms = the all-reduce "
                                                "of the P-vector inside initSess, timed on the device (sharding this pass pays only while rows/rank x row bytes is tens of MB: "
                                                "the reduction of 55 doubles costs more than a 12 us kernel saves)"}
    except Exception as e:   # noqa: BLE001
        errors["cv_full_gradient"] = f"{type(e).__name__}: {str(e)[:300]}"
    drop(sg)
    del sess, sg

    # ---- end-to-end leg through the public API with HOST buffers: Setup (dataset H2D) + initSess + K x sgm_sample (runSGMCMC's loop in one
    # call: the chain streams to the host as float32 while later iterations run, widened to float64 into the caller's array).  A fresh sampler:
    # nothing of the device leg above is reused.
    try:
        barrier()
        t_e2e0 = time.perf_counter()
        sg = new_sampler()
        torch.cuda.synchronize()
        t_setup = time.perf_counter()
        sess = S.initSess(sg, verbose=False)
        torch.cuda.synchronize()
        t_init = time.perf_counter()
        host_chain = np.empty(chains_local * sg.P * I, dtype=np.float64)
        hptr = host_chain.ctypes.data_as(_lib._PD)
        e2e_launches = 0
        for _ in range(args.steps):
            _lib.check(lib.sgm_sample(sg.handle, I, 100, hptr))
            e2e_launches += V.lastTiming(sg)[1]
        barrier()
        t_e2e = time.perf_counter() - t_e2e0
        t_run = time.perf_counter() - t_init
        if world > 1:
            t_e2e, t_run, t_up, t_in = max_over_ranks([t_e2e, t_run, t_setup - t_e2e0, t_init - t_setup])
        else:
            t_up, t_in = t_setup - t_e2e0, t_init - t_setup
        finite = bool(np.isfinite(host_chain[-sg.P:]).all())
        d2h_bytes = chains_local * sg.P * I * 4
        h2d_total = sum(np.asarray(v).size * 8 for v in (w["data"] or {}).values())
        del host_chain
        # the same K steps into a PAGE-LOCKED array (sgm_host_alloc): the GPU writes the doubles itself, no staging and no host threads.
        # Reported beside the headline (which stays on the pageable array a caller of the reference would hand over).
        pinned = None
        try:
            from sgmcmc_b200 import api as A
            pbuf = A._pinned_doubles(chains_local * sg.P * I)
            pptr = pbuf.ctypes.data_as(_lib._PD)
            _lib.check(lib.sgm_sample(sg.handle, I, 100, pptr))      # warm-up of the direct path
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                _lib.check(lib.sgm_sample(sg.handle, I, 100, pptr))
            barrier()
            t_pin = time.perf_counter() - t0
            if world > 1:
                t_pin = max_over_ranks([t_pin])[0]
            pinned = {"steady_state_value": total_chains * I * args.steps / t_pin, "ms_per_step": 1000.0 * t_pin / args.steps,
                      "d2h_bytes_per_step": chains_local * sg.P * I * 8, "last_draw_finite": bool(np.isfinite(pbuf[-sg.P:]).all()),
                      "note": "page-locked caller array: written by the GPU directly as float64 (compare with steady_state_value of the pageable array)"}
            del pbuf
        except Exception as e:   # noqa: BLE001
            pinned = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        line["e2e"] = {"value": total_chains * I * args.steps / t_e2e, "unit": UNIT,
                       "h2d_bytes_per_step": h2d_total / args.steps, "d2h_bytes_per_step": d2h_bytes,
                       "seconds": t_e2e,
                       "breakdown_ms": {"upload_setup": 1000.0 * t_up, "init_mode_and_full_gradient": 1000.0 * t_in, "run_and_copy_out": 1000.0 * t_run,
                                        "run_and_copy_out_per_step": 1000.0 * t_run / args.steps},
                       "steady_state_value": total_chains * I * args.steps / t_run,
                       "page_locked_output": pinned,
                       "comm_init_ms_outside": comm_ms, "last_draw_finite": finite,
                       "includes": "Setup (dataset upload, float64 host arrays) + initSess (mode finder, full gradient) + K x sgm_sample (sampling with the chain streamed "
                                   "out as float32 in 100-iteration blocks on a second stream and widened to float64 into a host array; divergence check every 100 iterations)",
                       "gpu_launches": e2e_launches}
        drop(sg)
        del sess, sg
    except Exception as e:   # noqa: BLE001  (the device-timed line above must survive a failure of this leg)
        errors["e2e"] = f"{type(e).__name__}: {str(e)[:300]}"

    # ---- optional legs.  N = 1: each in its own process (a device fault there cannot take the main line with it)
    if w["cv"] and w["family"] == "logistic" and not args.no_modes:
        try:
            if world > 1 and finder_probe_all_ranks_say_many_mode_bad:
                line["cv_full_gradient_modes"] = {"error": "the many-mode full gradient failed in the probe process (mode_finder_probe): not run in process"}
            else:
                line["cv_full_gradient_modes"] = (run_leg_in_child("modes", args) if world == 1 else many_mode_full_gradient(w, rank, world, args.modes))
        except Exception as e:   # noqa: BLE001
            errors["cv_full_gradient_modes"] = f"{type(e).__name__}: {str(e)[:300]}"
    if finder_probe is not None:
        line["mode_finder_probe"] = dict(finder_probe, pinned_to_one_cta_per_mode=PIN_ONE_CTA_FINDER)
    if world == 1 and w["cv"] and w["family"] == "logistic" and not args.no_extra:
        line["mode_finder"] = run_leg_in_child("finder", args) if not PIN_ONE_CTA_FINDER else {"error": "the cluster finder failed its probe; see mode_finder_probe"}
    if world == 1 and not args.no_extra:
        line["other_workloads"] = {name: run_leg_in_child("quick:" + name, args) for name in ("c1", "c2", "c4") if name != args.workload}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            ips, note, fg, _ = cpu_proxy(w, args.cpu_budget, 1)
            line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": 1, "kind": "port", "sample": note,
                                    "full_gradient_s": fg, "label": "TF-proxy (torch CPU) -- not the reference binary"}
        except Exception as e:   # noqa: BLE001
            errors["cpu_baseline"] = f"{type(e).__name__}: {str(e)[:300]}"
    if errors:
        line["errors"] = errors
    if rank == 0:
        print(json.dumps(line))
    if own_group:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SGM_WORKLOAD", "c3"))
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU")
    ap.add_argument("--iters-per-step", type=int, default=550)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--modes", type=int, default=1024, help="modes of the many-mode control-variate full gradient leg (logistic *cv workloads)")
    ap.add_argument("--no-modes", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the compact c1 / c2 / c4 measurements added to the line at N = 1")
    ap.add_argument("--leg", default=None, help="internal: run ONE optional leg (modes | finder | quick:c1 | quick:c2 | quick:c4) and print its JSON object")
    ap.add_argument("--peer-shards", action="store_true",
                    help="N > 1, logistic workloads: every rank keeps its own rows and maps the peers' (the c5 config at N = world x rows/GPU)")
    args = ap.parse_args()
    w = workload(args.workload, args.chains)
    if args.leg:
        run_leg(args, w)
    elif args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
