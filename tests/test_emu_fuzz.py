"""A fixed-seed batch of tests/fuzz_emu.py inside the CPU suite: random shapes of every family through the emulated build against the
oracle, plus the regression cases the fuzz has found so far."""
import numpy as np
import pytest

from tests import fuzz_emu
from tests.emu import backend


@pytest.fixture(autouse=True)
def emulated_library():
    with backend.emulated() as lib:
        yield lib
        msg = lib.emu_last_message()
        assert not msg, msg.decode()


@pytest.mark.parametrize("seed", [101, 102])
def test_random_shapes(seed):
    rs = np.random.RandomState(seed)
    for _ in range(8):
        kind, name, algo, mode = fuzz_emu.random_case(rs)
        try:
            fuzz_emu.run_case(kind, name, algo, mode)
        except Exception as e:   # noqa: BLE001
            raise AssertionError(f"{kind} {name} {algo} {mode}: {e}") from e


@pytest.mark.parametrize("kind,name,algo,mode", [
    # d = 72: ld = 76 floats, four unpadded rows are 1216 B = 9.5 lines of 128 B: every second tile::gather4 destination was 64 bytes
    # off a 128-byte boundary (a misaligned TMA destination on the device); such shapes now take one bulk copy per row
    ("steps", "logistic72_n532_N532", "sgld", "cv_shared"),
    ("steps", "logistic8_n100_N300", "sghmc", "plain"),
    # degenerate minibatch (n = N - 1 = 8): the control-variate gradient cancels to ~1e-3; tolerance scaled by the summand mass
    ("steps", "logistic18_n8_N9", "sghmc", "cv"),
    # n = 2 of N = 491 rows: N / n = 245 multiplies the float32 rounding of the net's O(1) pre-activations; judged against the summand mass
    ("steps", "softmaxnet_76_12_8_n2_N491", "sgld", "plain"),
])
def test_shapes_the_fuzz_found(kind, name, algo, mode):
    fuzz_emu.run_case(kind, name, algo, mode)


def test_gather4_is_only_chosen_for_rows_that_tile_128_byte_lines():
    """Launch plans (no device): tile::gather4 for ld % 8 == 0 unpadded rows (d = 54: ld = 56), bulk copy per row otherwise (d = 72: ld = 76)."""
    import ctypes as C
    import json
    from sgmcmc_b200 import _lib
    lib = _lib.load()
    for d, want in ((54, "TMA tile::gather4"), (72, "TMA bulk copy per row"), (8, "TMA bulk copy per row"), (20, "TMA tile::gather4")):
        q = _lib.SgmPlanQuery(model=_lib.MODEL_LOGISTIC, cv=1, N=5000, d=d, d_hidden=0, d_out=0, n=300, n_chains=4, flags=0, world=1, rank=0,
                              sharded_dataset=0, row0=-1, row1=-1, sms=148)
        buf = C.create_string_buffer(1 << 16)
        assert lib.sgm_debug_plan(C.byref(q), buf, len(buf)) == 0
        plan = json.loads(buf.value.decode())
        notes = [l["note"] for l in plan["launches"] if l["kernel"] == "logistic_staged_kernel"]
        assert notes and all(want in n for n in notes), (d, notes)
