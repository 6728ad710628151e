"""The TMA-staged FK_2c kernel (variant 7, tgtk_kernels_tma.cuh: cp.async.bulk.tensor tiles, producer warp, mbarrier
pipeline, wrap strips) against the CPU oracle on ragged, odd and tiny boxes, both dtypes, every stage / tile setting,
with and without a stop criterion, and against the default kernel over the benchmarked loop.  The variant is chosen at
creation (TGTK_VARIANT=7: its tensor maps need an even fp64 row pitch)."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

pytestmark = pytest.mark.gpu

shapes = st.tuples(st.integers(1, 24), st.integers(1, 40), st.integers(1, 80))


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(autouse=True)
def _env():
    os.environ["TGTK_VARIANT"] = "7"
    yield
    for k in ("TGTK_VARIANT", "TGTK_TMA_ROWS", "TGTK_TMA_STAGES", "TGTK_L2PF"):
        os.environ.pop(k, None)


def _case(shape, seed):
    rng = np.random.default_rng(seed)
    wm = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.8)
    gm = rng.uniform(0, 1, shape) * (1 - wm)
    P0 = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.6)
    N0 = rng.uniform(0, 0.6, shape) * (rng.uniform(0, 1, shape) < 0.5)
    S0 = rng.uniform(0.2, 1.0, shape)
    return wm, gm, P0, N0, S0


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(shape=shapes, seed=st.integers(0, 2 ** 20), nsteps=st.integers(1, 40), precision=st.sampled_from(["fp64", "fp32"]),
       rows=st.sampled_from(["4", "8"]), stages=st.sampled_from(["2", "3"]), l2pf=st.sampled_from(["0", "2"]),
       h=st.tuples(st.floats(0.5, 2.0), st.floats(0.5, 2.0), st.floats(0.5, 2.0)))
def test_tma_kernel_any_box(oracle, shape, seed, nsteps, precision, rows, stages, l2pf, h):
    from tumorgrowthtoolkit_b200 import _native as nat
    os.environ["TGTK_TMA_ROWS"], os.environ["TGTK_TMA_STAGES"], os.environ["TGTK_L2PF"] = rows, stages, l2pf
    wm, gm, P0, N0, S0 = _case(shape, seed)
    Dw, Ds, rho = 1.1, 1.3, 0.2
    dt = 0.05 * min(h) ** 2 / max(Dw, Ds)
    p = nat.make_params(nat.MODEL_FK_2C, nat.F64 if precision == "fp64" else nat.F32, shape, h, dt, rho=rho, Dw=Dw, ratio=10.0,
                        th_matter=0.1, th_necro=0.9, Ds=Ds, lambda_np=0.35, sigma_np=0.5, lambda_s=0.05)
    with nat.Sim(p) as sim:
        sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm)
        sim.upload(nat.FIELD_P, P0); sim.upload(nat.FIELD_N, N0); sim.upload(nat.FIELD_S, S0)
        sim.prepare()
        assert sim.tuning()["variant"] == 7
        sim.run(nsteps)
        st_ = sim.sync()
        got = {k: sim.download(f) for k, f in zip("PNS", (nat.FIELD_P, nat.FIELD_N, nat.FIELD_S))}
    ref = oracle.fk2c_loop(P0, N0, S0, wm, gm, 0.1, 0.9, Dw, 10.0, Ds, rho, 0.35, 0.5, 0.05, dt, *h, nsteps)
    assert st_.steps_run == nsteps
    tol = 1e-12 if precision == "fp64" else 5e-5
    for k in "PNS":
        assert np.max(np.abs(got[k] - ref["state"][k])) <= tol, k
    assert abs(st_.last_volume - ref["volume"]) <= (1e-10 if precision == "fp64" else 1e-4) * max(1.0, abs(ref["volume"]))


def test_tma_kernel_stop_criterion_and_snapshots(oracle):
    from tumorgrowthtoolkit_b200 import _loop
    shape = (14, 21, 45)
    wm, gm, P0, N0, S0 = _case(shape, 3)
    args = (P0 * 0.9, N0 * 0.3, S0, wm, gm, 0.1, 0.9, 0.9, 50.0, 1.3, 0.3, 0.35, 0.5, 0.05, 0.02, 1.0, 1.3, 0.7)
    free = oracle.fk2c_loop(*args, 60)
    first = oracle.fk2c_loop(*args, 1)
    th = 0.5 * (free["volume"] + first["volume"])
    rec = np.arange(0, 60, 7)
    ref = oracle.fk2c_loop(*args, 60, stopping_volume=th, record_steps=rec)
    got = _loop.fk2c_loop(*args, 60, stopping_volume=th, record_steps=rec)
    assert got["stop"] == ref["stop"] == "volume" and got["steps_run"] == ref["steps_run"] and got["t_stop"] == ref["t_stop"]
    for k in "PNS":
        assert np.max(np.abs(got["state"][k] - ref["state"][k])) <= 1e-12
        assert len(got["snapshots"][k]) == len(ref["snapshots"][k])
