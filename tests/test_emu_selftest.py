"""The CPU emulation of the kernel sources (tests/emu) must REPORT the failure classes it exists to catch: deliberately broken
kernels (tests/emu/selftest.cpp) -- a read that misses its mbarrier wait, a deadlock, out-of-range global / shared accesses, a
TMA box whose innermost start coordinate is not 16-byte aligned (the Xid 13 of round 1's 8-GPU run), invalid tensor maps."""
import ctypes as C

import pytest

from tests.emu import backend


@pytest.fixture(scope="module")
def emu():
    lib = backend.library()
    for name in ("selftest_deadlock", "selftest_oob_global", "selftest_oob_shared", "selftest_divergent_barrier"):
        getattr(lib, name).restype = C.c_char_p
    lib.selftest_missing_wait.restype = C.c_char_p
    lib.selftest_missing_wait.argtypes = [C.c_int, C.POINTER(C.c_float)]
    lib.selftest_tma_unaligned.restype = C.c_char_p
    lib.selftest_tma_unaligned.argtypes = [C.c_int]
    lib.selftest_cluster_exchange.restype = C.c_char_p
    lib.selftest_cluster_exchange.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
    yield lib
    lib.emu_reset_error()


def test_read_without_the_mbarrier_wait_sees_stale_data(emu):
    v = C.c_float()
    assert emu.selftest_missing_wait(1, C.byref(v)) == b"" and v.value == 5.0
    if emu.emu_async_is_lazy():
        assert emu.selftest_missing_wait(0, C.byref(v)) == b""
        assert v.value == -1.0, "asynchronous copies land at the last moment the memory model allows"


def test_deadlock_is_reported_not_hung(emu):
    assert b"DEADLOCK" in emu.selftest_deadlock()
    assert b"DEADLOCK" in emu.selftest_divergent_barrier()


def test_out_of_range_accesses_fault(emu):
    assert b"illegal address" in emu.selftest_oob_global()
    assert b"shared-memory access out of range" in emu.selftest_oob_shared()


def test_unaligned_tma_coordinate_is_an_illegal_instruction(emu):
    assert emu.selftest_tma_unaligned(32) == b""
    assert b"not 16-byte aligned" in emu.selftest_tma_unaligned(33)


def test_cluster_protocol_errors_are_reported(emu):
    """Thread-block clusters (the cluster mode finder, csrc/logistic_split.cuh): the CTAs of a cluster run one after another as far
    as each can get, so a CTA that should have waited for its peers reads their shared memory before it is written, and a CTA that
    leaves without the final barrier is gone when a peer comes to read it."""
    v = C.c_float()
    for cluster in (2, 4, 8):
        assert emu.selftest_cluster_exchange(cluster, 0, C.byref(v)) == b"" and v.value == cluster * (cluster + 1) / 2
    msg = emu.selftest_cluster_exchange(4, 1, C.byref(v))
    assert msg == b"" and v.value != v.value, "a read of a peer's slot before the cluster barrier sees unwritten (poisoned) shared memory"
    assert b"already exited" in emu.selftest_cluster_exchange(4, 2, C.byref(v))
    emu.emu_reset_error()


def test_refilled_stage_is_seen_by_late_readers(emu):
    assert emu.selftest_refill_race() > 0


@pytest.mark.parametrize("which,ok", [(0, 1), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0)])
def test_tensor_map_validation(emu, which, ok):
    assert emu.selftest_encode(which) == ok
