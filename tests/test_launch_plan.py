"""Launch plans of the five BASELINE.json configurations, computed by the product library WITHOUT a device (sgm_debug_plan): every
kernel's grid / block / shared memory / tensor memory, every TMA tensor map, and the rules whose violation is an Xid on the GPU
(16-byte aligned TMA box starts and strides, boxes inside their allocation, <= 227 KB shared memory, <= 512 TMEM columns, grid
limits).  Round 1 lost its GPU access to one such violation at 8 GPUs (a row shard starting at an odd row); this sweep covers
world sizes 1, 2, 3, 4, 8 and shard starts at every residue modulo 32, at the full sizes, on a CPU.

The library itself runs the same check before it launches the tensor-core pass (csrc/api.cu: cvg_plan_check)."""
import ctypes as C
import json

import pytest

from sgmcmc_b200 import _lib

GAUSS, MIX, LOGISTIC, NET = _lib.MODEL_GAUSS_MEAN, _lib.MODEL_MIXTURE2, _lib.MODEL_LOGISTIC, _lib.MODEL_SOFTMAX_NET
SHARED = _lib.FLAG_SHARED_MODE

# BASELINE.json configs at their full sizes; c5: 1.25e7 rows and 128 chains per GPU
CONFIGS = {
    "c1": dict(model=GAUSS, cv=0, N=10 ** 4, d=2, n=100, n_chains=148 * 64),
    "c2": dict(model=MIX, cv=0, N=10 ** 4, d=2, n=100, n_chains=148 * 64),
    "c3": dict(model=LOGISTIC, cv=1, N=581012, d=54, n=500, n_chains=148 * 8, flags=SHARED),
    "c3_modes": dict(model=LOGISTIC, cv=1, N=581012, d=54, n=500, n_chains=1024),
    "c4": dict(model=NET, cv=0, N=60000, d=784, d_hidden=100, d_out=10, n=100, n_chains=148),
    "c5": dict(model=LOGISTIC, cv=1, N=12500000, d=1024, n=1024, n_chains=148, flags=SHARED),
    "c5_modes": dict(model=LOGISTIC, cv=1, N=12500000, d=1024, n=1024, n_chains=128),
}


def plan(lib, world=1, rank=0, row0=-1, row1=-1, sharded=0, sms=148, **cfg):
    q = _lib.SgmPlanQuery()
    q.model, q.cv, q.N, q.d = cfg["model"], cfg["cv"], cfg["N"], cfg["d"]
    q.d_hidden, q.d_out = cfg.get("d_hidden", 0), cfg.get("d_out", 0)
    q.n, q.n_chains, q.flags = cfg["n"], cfg["n_chains"], cfg.get("flags", 0)
    q.world, q.rank, q.sharded_dataset, q.row0, q.row1, q.sms = world, rank, sharded, row0, row1, sms
    buf = C.create_string_buffer(1 << 16)
    rc = lib.sgm_debug_plan(C.byref(q), buf, len(buf))
    assert rc >= 0, lib.sgm_last_error()
    js = json.loads(buf.value.decode())
    assert rc == len(js["violations"])
    return js


@pytest.mark.parametrize("name", sorted(CONFIGS))
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_baseline_configs_plan_cleanly_at_every_world_size(lib, name, world):
    for rank in range(world):
        js = plan(lib, world=world, rank=rank, **CONFIGS[name])
        assert js["violations"] == [], (name, world, rank, js["violations"])
        for l in js["launches"]:
            assert l["smem"] <= 227 * 1024 and l["tmem_cols"] <= 512 and 1 <= l["block"] <= 1024


@pytest.mark.parametrize("name", ["c3_modes", "c5_modes"])
def test_row_ranges_starting_anywhere(lib, name):
    """The many-mode pass on a row range that starts at ANY row (every residue modulo 32, including the odd starts that were an
    illegal instruction in round 1) and ends anywhere."""
    cfg = CONFIGS[name]
    N = cfg["N"]
    for start in list(range(0, 40)) + [N // 8 + 1, N // 3, N // 2 + 17, N - 4099]:
        for length in (1, 127, 4099, N // 8):
            if start + length > N:
                continue
            js = plan(lib, world=8, rank=1, row0=start, row1=start + length, **cfg)
            assert js["violations"] == [], (name, start, length, js["violations"])


def test_the_sharded_dataset_plan_uses_the_peer_table(lib):
    js = plan(lib, world=8, rank=3, sharded=1, **CONFIGS["c5"])
    assert js["violations"] == []
    assert js["rows"] == [0, CONFIGS["c5"]["N"]], "a sharded dataset sums the whole local shard"
    assert any("peer-shard table" in l["note"] for l in js["launches"])


def test_plan_reports_what_the_bench_runs(lib):
    """The kernels named in profiles/*launch_list* are the ones the plan (= the dispatch code) selects for the bench configs."""
    c3 = plan(lib, **CONFIGS["c3"])
    k = {l["kernel"]: l for l in c3["launches"]}
    assert "logistic_staged_kernel" in k and "R = 2" in k["logistic_staged_kernel"]["note"] and "gather4" in k["logistic_staged_kernel"]["note"]
    assert "logistic_fullgrad_stream_kernel<short>" in k and "logistic_cache_rhat_kernel" in k
    c5 = plan(lib, **CONFIGS["c5"])
    k = {l["kernel"]: l for l in c5["launches"]}
    assert "<G, V> = <32, 9>" in k["logistic_chain_kernel"]["note"] and "logistic_fullgrad_stream_kernel<long>" in k
    c5m = plan(lib, **CONFIGS["c5_modes"])
    assert sum(l["kernel"] == "cvg_kernel" for l in c5m["launches"]) >= 2
    c4 = plan(lib, **CONFIGS["c4"])
    assert "TMA-fed" in c4["launches"][0]["note"] and c4["launches"][0]["tmem_cols"] == 512


def test_violations_are_reported(lib):
    """A plan the hardware would refuse is reported, not launched: a minibatch whose staging does not fit, a mode count beyond the
    grid limit."""
    big = dict(CONFIGS["c3"], n=2000)            # 2000 x 224 B > 227 KB: the short-row kernel is not selected, the generic one is
    js = plan(lib, **big)
    assert js["violations"] == [] and js["launches"][0]["kernel"] == "logistic_chain_kernel"
    js = plan(lib, **dict(CONFIGS["c3_modes"], n_chains=70000, d=5))   # 70000 per-chain modes on the SIMT grid: grid.y > 65535 ...
    # ... is avoided because the tensor-core path takes over from 16 modes on
    assert js["violations"] == []
