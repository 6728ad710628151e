"""Runs with a stop criterion (FK.py:216-218, FK_DTI.py:199-206) are executed speculatively: whole
chunks of host-scheduled launches behind a checkpoint, stop tests evaluated per chunk afterwards, a trip
replayed from the checkpoint up to the stopping step.  The result must be the device-scheduled loop's
(TGTK_SPECULATE=0: every launch runs the tests itself and freezes the state) bit for bit -- final state,
step counts, stop step, volume, snapshots -- wherever the stopping step falls inside a chunk."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(autouse=True)
def _restore_env():
    yield
    os.environ.pop("TGTK_SPECULATE", None)


def _inputs(shape=(20, 26, 23), seed=3):
    rng = np.random.default_rng(seed)
    wm = rng.uniform(0.2, 0.9, shape)
    gm = 1.0 - wm
    x, y, z = np.meshgrid(*[np.arange(n) - n // 2 for n in shape], indexing="ij")
    u0 = 0.6 * np.exp(-(x * x + y * y + z * z) / 12.0)
    return wm, gm, u0


def _both_modes(fn):
    out = {}
    for mode in ("0", "1"):
        os.environ["TGTK_SPECULATE"] = mode
        out[mode] = fn()
    return out["0"], out["1"]


def _same(a, b):
    assert a["steps_run"] == b["steps_run"] and a["stop"] == b["stop"] and a["t_stop"] == b["t_stop"]
    assert a["volume"] == b["volume"]
    sa, sb = a["state"], b["state"]
    if isinstance(sa, dict):
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), k
    else:
        assert np.array_equal(sa, sb)
    if a["snapshots"] is not None:
        la = a["snapshots"] if not isinstance(a["snapshots"], dict) else a["snapshots"]["P"]
        lb = b["snapshots"] if not isinstance(b["snapshots"], dict) else b["snapshots"]["P"]
        assert len(la) == len(lb)
        for p, q in zip(la, lb):
            assert np.array_equal(p, q)


@pytest.mark.parametrize("k", [0, 1, 5, 30, 31, 32, 33, 63, 64, 70, 95])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_fk_stop_anywhere_in_a_chunk(oracle, k, precision):
    from tumorgrowthtoolkit_b200 import _loop
    wm, gm, u0 = _inputs()
    args = (u0, wm, gm, 0.1, 1.0, 10.0, 0.4, 0.02, 1.0, 1.0, 1.0)
    os.environ["TGTK_SPECULATE"] = "1"
    free = _loop.fk_loop(*args, 100, precision=precision)
    assert free["steps_run"] == 100 and free["stop"] == "time"
    # the growing tumour's volume is monotone: a threshold between steps k-1 and k trips exactly at k
    record = np.unique(np.linspace(0, 99, 12, dtype=int))

    def solve_to(th):
        return lambda: _loop.fk_loop(*args, 100, stopping_volume=th, record_steps=record, precision=precision)

    # find the threshold from a speculative run's own volume log: stop at step k
    from tumorgrowthtoolkit_b200 import _native as nat
    p = nat.make_params(nat.MODEL_FK, _loop.dtype_code(precision), u0.shape, (1.0, 1.0, 1.0), 0.02, rho=0.4, Dw=1.0,
                        ratio=10.0, th_matter=0.1)
    with nat.Sim(p) as sim:
        sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm); sim.upload(nat.FIELD_U, u0)
        sim.prepare()
        sim.run(100)
        sim.sync()
        v = sim.volumes(0, 100)
    assert np.all(np.diff(v) > 0)
    th = 0.5 * (v[k] + (v[k - 1] if k > 0 else 0.0))
    a, b = _both_modes(solve_to(th))
    _same(a, b)
    assert b["steps_run"] == k + 1 and b["stop"] == "volume" and b["t_stop"] == k
    assert len(b["snapshots"]) == int(np.sum(record < k))
    if precision == "fp64":
        ref = oracle.fk_loop(*args, 100, stopping_volume=th)
        assert ref["steps_run"] == k + 1
        assert np.max(np.abs(b["state"] - ref["state"])) <= 1e-12


def test_fk2c_and_dti_stops_match_the_device_scheduled_loop():
    from tumorgrowthtoolkit_b200 import _loop
    wm, gm, u0 = _inputs((18, 21, 37))
    S0 = np.where(wm + gm >= 0.1, 1.0, 0.0)
    args2c = (u0, np.zeros_like(u0), S0, wm, gm, 0.1, 0.9, 0.9, 100.0, 1.3, 0.5, 0.35, 0.5, 0.05, 0.02, 1.0, 1.0, 1.0)
    free = _loop.fk2c_loop(*args2c, 80)
    th = free["volume"] * 0.97
    a, b = _both_modes(lambda: _loop.fk2c_loop(*args2c, 80, stopping_volume=th, record_steps=np.arange(0, 80, 7)))
    _same(a, b)
    assert b["stop"] == "volume" and 0 < b["steps_run"] < 80
    # FK_DTI: decay below the 1e-6 volume floor (FK_DTI.py:203-206) and a volume stop
    rgb = np.stack([wm, 0.7 * wm, 0.4 * wm], axis=-1)
    tiny = 1e-6 * u0
    a, b = _both_modes(lambda: _loop.dti_loop(tiny, rgb, 1.0, -5.0, 0.02, 1.0, 1.0, 1.0, 90))
    _same(a, b)
    assert b["stop"] == "small" and b["steps_run"] < 90
    free = _loop.dti_loop(u0, rgb, 1.0, 0.5, 0.02, 1.0, 1.0, 1.0, 70)
    a, b = _both_modes(lambda: _loop.dti_loop(u0, rgb, 1.0, 0.5, 0.02, 1.0, 1.0, 1.0, 70, stopping_volume=free["volume"] * 0.9,
                                              record_steps=np.arange(0, 70, 9)))
    _same(a, b)
    assert b["stop"] == "volume"


def test_no_trip_runs_to_the_end_and_reset_reruns():
    from tumorgrowthtoolkit_b200 import _native as nat
    wm, gm, u0 = _inputs()
    p = nat.make_params(nat.MODEL_FK, nat.F64, u0.shape, (1.0, 1.0, 1.0), 0.02, rho=0.4, Dw=1.0, ratio=10.0, th_matter=0.1,
                        stopping_volume=1e9)
    with nat.Sim(p) as sim:
        sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm); sim.upload(nat.FIELD_U, u0)
        sim.prepare()
        sim.run(75)
        st = sim.sync()
        first = sim.download(nat.FIELD_U)
        assert st.steps_run == 75 and st.stop_reason == nat.STOP_NONE
        sim.reset()
        sim.run(40); sim.run(35)
        st = sim.sync()
        assert st.steps_run == 75
        assert np.array_equal(first, sim.download(nat.FIELD_U))
