"""Fused slab driver (tgtk_sim_slab_attach: halo pushes and epoch hand-shakes inside the step kernels) on ONE rank:
the periodic ring closes on the rank itself, so every mechanism of the decomposed run -- restricted plane range,
in-kernel pushes into the ghost planes, epoch words, deferred chunk reduction + commit, checkpoint / replay of a
tripped stop criterion, snapshots, reset -- is exercised and must reproduce the undecomposed loop bit for bit.
The multi-rank runs are in tests/test_gpu_multi.py (torchrun, skipped on a one-GPU box)."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

pytestmark = pytest.mark.gpu

shapes = st.tuples(st.integers(1, 20), st.integers(1, 40), st.integers(1, 70))


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(autouse=True)
def _env():
    os.environ["TGTK_VARIANT"] = "4"       # the undecomposed side runs the same two-voxel kernels (fp32 FK_DTI defaults to z pairs)
    os.environ["TGTK_TB"] = "0"
    yield
    os.environ.pop("TGTK_VARIANT", None)
    os.environ.pop("TGTK_TB", None)


def _fields(shape, seed):
    rng = np.random.default_rng(seed)
    wm = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.8)
    gm = rng.uniform(0, 1, shape) * (1 - wm)
    u0 = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.6)
    rgb = rng.uniform(0, 3, shape + (3,)) * (rng.uniform(0, 1, shape) < 0.8)[..., None]
    return wm, gm, u0, rgb


def _pair(model, shape, seed, nsteps, precision, **kw):
    from tumorgrowthtoolkit_b200 import _loop, slab
    wm, gm, u0, rgb = _fields(shape, seed)
    if model == "fk":
        args = (u0, wm, gm, 0.1, 1.1, 10.0, 0.2, 0.02, 1.0, 1.3, 0.7, nsteps)
        return _loop.fk_loop(*args, precision=precision, **kw), slab.fk_loop(*args, precision=precision, **kw)
    if model == "dti":
        args = (u0, rgb, 0.9, 0.2, 0.004, 1.0, 1.3, 0.7, nsteps)
        skw = dict(kw, small_volume=1e-6)
        return _loop.dti_loop(*args, precision=precision, **kw), slab.dti_loop(*args, precision=precision, **skw)
    S0 = np.where(wm + gm >= 0.1, 1.0, 0.0) * np.random.default_rng(seed + 1).uniform(0.3, 1.0, shape)
    args = (0.9 * u0, 0.3 * u0[::-1].copy(), S0, wm, gm, 0.1, 0.9, 0.9, 50.0, 1.3, 0.3, 0.35, 0.5, 0.05, 0.02, 1.0, 1.3, 0.7, nsteps)
    return _loop.fk2c_loop(*args, precision=precision, **kw), slab.fk2c_loop(*args, precision=precision, **kw)


def _same(a, b):
    if isinstance(a, dict):
        return all(np.array_equal(a[k], b[k]) for k in a)
    return np.array_equal(a, b)


@settings(max_examples=24, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(model=st.sampled_from(["fk", "dti", "fk2c"]), shape=shapes, seed=st.integers(0, 2 ** 20), nsteps=st.integers(1, 70),
       precision=st.sampled_from(["fp64", "fp32"]))
def test_fused_ring_on_one_rank_is_bitwise(model, shape, seed, nsteps, precision):
    a, b = _pair(model, shape, seed, nsteps, precision)
    assert a["steps_run"] == b["steps_run"] == nsteps
    assert _same(a["state"], b["state"])
    assert abs(a["volume"] - b["volume"]) <= 1e-12 * max(1.0, abs(a["volume"]))


@pytest.mark.parametrize("model", ["fk", "dti", "fk2c"])
def test_fused_stop_criterion_and_snapshots(model):
    """A volume threshold crossed mid-chunk: same stop step, same final state, same valid snapshots."""
    shape = (13, 21, 37)
    free, _ = _pair(model, shape, 11, 50, "fp64")
    first, _ = _pair(model, shape, 11, 1, "fp64")
    th = 0.5 * (free["volume"] + first["volume"])
    rec = np.arange(0, 50, 7)
    a, b = _pair(model, shape, 11, 50, "fp64", stopping_volume=th, record_steps=rec)
    assert a["stop"] == b["stop"] == "volume" and a["steps_run"] == b["steps_run"] and a["t_stop"] == b["t_stop"]
    assert 1 < b["steps_run"] < 50
    assert _same(a["state"], b["state"])
    sa, sb = a["snapshots"], b["snapshots"]
    if isinstance(sa, dict):
        sa, sb = sa["P"], sb["P"]
    assert len(sa) == len(sb) > 0
    for p, q in zip(sa, sb):
        assert np.array_equal(p, q)


def test_fused_session_reset_and_rerun():
    from tumorgrowthtoolkit_b200 import _loop, slab
    shape = (9, 17, 33)
    _, _, u0, rgb = _fields(shape, 5)
    sess = slab.session_dti(u0, rgb, 0.9, 0.2, 0.004, 1.0, 1.3, 0.7)
    ref = _loop.dti_loop(u0, rgb, 0.9, 0.2, 0.004, 1.0, 1.3, 0.7, 45)
    for _ in range(3):
        sess.reset()
        sess.run(45)
        assert sess.sync_ms() > 0
        assert np.array_equal(sess.gather(), ref["state"])
        assert abs(sess.volume() - ref["volume"]) <= 1e-12 * abs(ref["volume"])
    sess.close()


def test_fused_small_volume_exit():
    """FK_DTI.py:203-206: `volume < 1e-6` ends the loop; decomposed run stops at the same step."""
    from tumorgrowthtoolkit_b200 import _loop, slab
    shape = (8, 12, 20)
    _, _, u0, rgb = _fields(shape, 2)
    u0 = u0 * 1e-6
    args = (u0, rgb, 0.9, -3.0, 0.02, 1.0, 1.3, 0.7, 120)
    a = _loop.dti_loop(*args)
    b = slab.dti_loop(*args, small_volume=1e-6)
    assert a["stop"] == b["stop"] == "small" and a["steps_run"] == b["steps_run"] < 120
    assert np.array_equal(a["state"], b["state"])
