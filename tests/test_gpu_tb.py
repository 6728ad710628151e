"""Temporal blocking (tgtk_kernels_tb.cuh): FK_DTI with two time steps per launch must be the one-step
loop bit for bit -- same arithmetic per voxel, only the launch structure differs -- for any box, any step
count (pairs + single-step remainders), with snapshots, stop criteria and decomposed slabs."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

pytestmark = pytest.mark.gpu

shapes = st.tuples(st.integers(1, 24), st.integers(1, 40), st.integers(1, 70))


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(autouse=True)
def _restore_env():
    yield
    for k in ("TGTK_TB", "TGTK_SPECULATE", "TGTK_L2PF", "TGTK_VARIANT", "TGTK_TB_ROWS"):
        os.environ.pop(k, None)


def _run(shape, seed, nsteps, precision, tb, record=None, stopping_volume=np.inf, rho=0.2, l2pf="-1", rows="18"):
    from tumorgrowthtoolkit_b200 import _loop
    os.environ["TGTK_TB"], os.environ["TGTK_L2PF"], os.environ["TGTK_TB_ROWS"] = str(tb), str(l2pf), str(rows)
    # the fp32 z-pair kernels (auto for FK_DTI) contract mul/sub into FFMA2; the two-step kernel rounds every
    # operation like the two-voxel kernels, which are therefore the bitwise reference in fp32
    os.environ["TGTK_VARIANT"] = "4"
    rng = np.random.default_rng(seed)
    rgb = rng.uniform(0, 3, shape + (3,)) * (rng.uniform(0, 1, shape) < 0.8)[..., None]
    u0 = rng.uniform(0, 1, shape)
    return _loop.dti_loop(u0, rgb, 0.9, rho, 0.004, 1.0, 1.3, 0.7, nsteps, stopping_volume=stopping_volume,
                          record_steps=record, precision=precision), (u0, rgb)


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(shape=shapes, seed=st.integers(0, 2 ** 20), nsteps=st.integers(1, 75), precision=st.sampled_from(["fp64", "fp32"]),
       l2pf=st.sampled_from(["0", "2"]), rows=st.sampled_from(["10", "18"]), with_record=st.booleans())
def test_two_step_launches_equal_one_step_launches(shape, seed, nsteps, precision, l2pf, rows, with_record):
    record = np.unique(np.linspace(0, nsteps - 1, 3, dtype=int)) if with_record else np.array([nsteps - 1])
    a, _ = _run(shape, seed, nsteps, precision, 0, record, l2pf=l2pf)
    b, _ = _run(shape, seed, nsteps, precision, 1, record, l2pf=l2pf, rows=rows)
    assert a["steps_run"] == b["steps_run"] == nsteps
    assert np.array_equal(a["state"], b["state"])
    assert abs(a["volume"] - b["volume"]) <= 1e-12 * max(1.0, abs(a["volume"]))
    assert len(a["snapshots"]) == len(b["snapshots"])
    for p, q in zip(a["snapshots"], b["snapshots"]):
        assert np.array_equal(p, q)
    if nsteps >= 5 and not with_record:
        assert b["launches"] < a["launches"]


def test_bitwise_equal_to_the_oracle(oracle):
    shape, nsteps = (9, 33, 61), 23
    b, (u0, rgb) = _run(shape, 4, nsteps, "fp64", 1)
    ref = oracle.dti_loop(u0, rgb, 0.9, 0.2, 0.004, 1.0, 1.3, 0.7, nsteps)
    assert np.array_equal(b["state"], ref["state"])


@pytest.mark.parametrize("nsteps", [5, 7, 38])
def test_fp32_zpair_kernel_on_odd_z_with_two_step_launches_requested(oracle, nsteps):
    """TGTK_TB=1 + the fp32 z-pair kernel (auto for FK_DTI) + odd Z + nsteps % 4 != 0 mixed
    two-step launches, which do not write the z = Z mirror element, with z-pair leftovers that read it as their
    periodic neighbour.  The combination now falls back to one-step launches: same bits as TGTK_TB=0, oracle-close."""
    from tumorgrowthtoolkit_b200 import _loop
    shape = (11, 14, 37)
    rng = np.random.default_rng(3)
    rgb = rng.uniform(0.5, 3, shape + (3,))
    u0 = rng.uniform(0, 1, shape)
    args = (u0, rgb, 0.9, 0.2, 0.004, 1.0, 1.3, 0.7, nsteps)
    os.environ.pop("TGTK_VARIANT", None)
    os.environ["TGTK_TB"] = "0"
    a = _loop.dti_loop(*args, precision="fp32")
    os.environ["TGTK_TB"] = "1"
    b = _loop.dti_loop(*args, precision="fp32")
    assert np.array_equal(a["state"], b["state"])
    ref = oracle.dti_loop(*args)
    assert np.max(np.abs(b["state"] - ref["state"])) <= 2e-5


@pytest.mark.parametrize("spec", ["0", "1"])
def test_stop_criterion_with_two_step_launches(spec):
    os.environ["TGTK_SPECULATE"] = spec
    shape = (12, 20, 37)
    free, _ = _run(shape, 9, 60, "fp64", 0, rho=2.0)
    th = 0.5 * (free["volume"] + _run(shape, 9, 1, "fp64", 0, rho=2.0)[0]["volume"])
    a, _ = _run(shape, 9, 60, "fp64", 0, np.arange(0, 60, 7), stopping_volume=th, rho=2.0)
    b, _ = _run(shape, 9, 60, "fp64", 1, np.arange(0, 60, 7), stopping_volume=th, rho=2.0)
    assert a["stop"] == b["stop"] == "volume" and a["steps_run"] == b["steps_run"] and a["t_stop"] == b["t_stop"]
    assert 1 < b["steps_run"] < 60
    assert np.array_equal(a["state"], b["state"])
    assert len(a["snapshots"]) == len(b["snapshots"])


@pytest.mark.parametrize("driver", ["p2p", "native", "torch"])
def test_slab_blocks_with_two_step_launches(oracle, driver):
    """One rank, periodic ring on itself (world == 1): halo 4 -> two 2-step launches per block, halo 5 -> two
    2-step launches and one single step; bitwise the undecomposed loop."""
    from tumorgrowthtoolkit_b200 import slab
    os.environ["TGTK_TB"] = "1"
    rng = np.random.default_rng(1)
    shape = (16, 13, 21)
    rgb = rng.uniform(0, 3, shape + (3,))
    u0 = rng.uniform(0, 1, shape)
    ref = oracle.dti_loop(u0, rgb, 0.9, 0.2, 0.004, 1.0, 1.3, 0.7, 22)
    for halo in (4, 5):
        got = slab.dti_loop(u0, rgb, 0.9, 0.2, 0.004, 1.0, 1.3, 0.7, 22, halo=halo, driver=driver)
        assert np.array_equal(got["state"], ref["state"]), halo


# ---- FK -------------------------------------------------------------------------------------------------
def _run_fk(shape, seed, nsteps, precision, tb, record=None, stopping_volume=np.inf, l2pf="-1"):
    from tumorgrowthtoolkit_b200 import _loop
    os.environ["TGTK_TB"], os.environ["TGTK_L2PF"] = str(tb), str(l2pf)
    rng = np.random.default_rng(seed)
    wm = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.8)
    gm = rng.uniform(0, 1, shape) * (1 - wm)
    u0 = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.6)
    args = (u0, wm, gm, 0.1, 1.1, 10.0, 0.2, 0.02, 1.0, 1.3, 0.7)
    return _loop.fk_loop(*args, nsteps, stopping_volume=stopping_volume, record_steps=record, precision=precision), args


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(shape=shapes, seed=st.integers(0, 2 ** 20), nsteps=st.integers(1, 75), precision=st.sampled_from(["fp64", "fp32"]),
       l2pf=st.sampled_from(["0", "2"]))
def test_fk_two_step_launches(oracle, shape, seed, nsteps, precision, l2pf):
    """Same operator, different association of the face sums than fk_step_v4: equal to rounding, not bitwise."""
    tol = 1e-12 if precision == "fp64" else 2e-5
    a, args = _run_fk(shape, seed, nsteps, precision, 0, [nsteps - 1], l2pf=l2pf)
    b, _ = _run_fk(shape, seed, nsteps, precision, 1, [nsteps - 1], l2pf=l2pf)
    assert a["steps_run"] == b["steps_run"] == nsteps
    assert np.max(np.abs(a["state"] - b["state"])) <= tol
    assert abs(a["volume"] - b["volume"]) <= 10 * tol * max(1.0, abs(a["volume"]))
    if nsteps >= 5:
        assert b["launches"] < a["launches"]
    ref = oracle.fk_loop(*args, nsteps)
    assert np.max(np.abs(b["state"] - ref["state"])) <= tol


def test_fk_stop_criterion_and_slab_with_two_step_launches(oracle):
    from tumorgrowthtoolkit_b200 import slab
    shape = (14, 19, 37)
    free, args = _run_fk(shape, 5, 50, "fp64", 1)
    th = free["volume"] * 0.97 if free["volume"] > _run_fk(shape, 5, 1, "fp64", 1)[0]["volume"] else free["volume"] * 1.03
    a, _ = _run_fk(shape, 5, 50, "fp64", 0, np.arange(0, 50, 6), stopping_volume=th)
    b, _ = _run_fk(shape, 5, 50, "fp64", 1, np.arange(0, 50, 6), stopping_volume=th)
    assert a["stop"] == b["stop"] and a["steps_run"] == b["steps_run"] and a["t_stop"] == b["t_stop"]
    assert np.max(np.abs(a["state"] - b["state"])) <= 1e-12
    os.environ["TGTK_TB"] = "1"
    ref = oracle.fk_loop(*args, 22)
    for driver in ("p2p", "native", "torch"):
        got = slab.fk_loop(*args, 22, halo=4, driver=driver)
        assert np.max(np.abs(got["state"] - ref["state"])) <= 1e-12, driver
