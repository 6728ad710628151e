"""hypothesis-driven shape / parameter fuzzing of the step kernels against the CPU oracle
(test plan item 5): every kernel family, degenerate boxes, anisotropic spacing."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

pytestmark = pytest.mark.gpu

shapes = st.tuples(st.integers(1, 40), st.integers(1, 40), st.integers(1, 70))
# library tunables read at tgtk_sim_create: L2 prefetch distance (-1 auto, 0 off) and programmatic dependent launch
tunables = st.tuples(st.sampled_from([-1, 0, 1, 2, 5]), st.sampled_from([0, 1]))


def _set_tunables(t):
    import os
    os.environ["TGTK_L2PF"], os.environ["TGTK_PDL"] = str(t[0]), str(t[1])


@pytest.fixture(autouse=True)
def _restore_env():
    import os
    yield
    os.environ.pop("TGTK_L2PF", None)
    os.environ.pop("TGTK_PDL", None)


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(shape=shapes, seed=st.integers(0, 2 ** 20), variant=st.sampled_from([1, 2, 3, 4]),
       h=st.tuples(st.floats(0.5, 2.0), st.floats(0.5, 2.0), st.floats(0.5, 2.0)), nsteps=st.integers(1, 6), tun=tunables)
def test_fk_and_fk2c_any_box(oracle, shape, seed, variant, h, nsteps, tun):
    from tumorgrowthtoolkit_b200 import _native as nat
    _set_tunables(tun)
    rng = np.random.default_rng(seed)
    wm = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.8)
    gm = rng.uniform(0, 1, shape) * (1 - wm)
    u0 = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.6)
    N0 = rng.uniform(0, 0.6, shape) * (rng.uniform(0, 1, shape) < 0.5)
    S0 = rng.uniform(0.2, 1.0, shape)
    Dw, Ds, rho = 1.1, 1.3, 0.2
    dt = 0.05 * min(h) ** 2 / max(Dw, Ds)
    # FK
    p = nat.make_params(nat.MODEL_FK, nat.F64, shape, h, dt, rho=rho, Dw=Dw, ratio=10.0, th_matter=0.1)
    with nat.Sim(p) as sim:
        sim.set_variant(variant)
        sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm); sim.upload(nat.FIELD_U, u0)
        sim.prepare()
        sim.run(nsteps)
        st_ = sim.sync()
        got = sim.download(nat.FIELD_U)
    ref = oracle.fk_loop(u0, wm, gm, 0.1, Dw, 10.0, rho, dt, *h, nsteps)
    assert st_.steps_run == nsteps
    assert np.max(np.abs(got - ref["state"])) <= 1e-12
    assert abs(st_.last_volume - ref["volume"]) <= 1e-10 * max(1.0, abs(ref["volume"]))
    # FK_2c
    p = nat.make_params(nat.MODEL_FK_2C, nat.F64, shape, h, dt, rho=rho, Dw=Dw, ratio=10.0, th_matter=0.1, th_necro=0.9,
                        Ds=Ds, lambda_np=0.35, sigma_np=0.5, lambda_s=0.05)
    with nat.Sim(p) as sim:
        sim.set_variant(variant)
        sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm)
        sim.upload(nat.FIELD_P, u0); sim.upload(nat.FIELD_N, N0); sim.upload(nat.FIELD_S, S0)
        sim.prepare()
        sim.run(nsteps)
        sim.sync()
        got = {k: sim.download(f) for k, f in (("P", nat.FIELD_P), ("N", nat.FIELD_N), ("S", nat.FIELD_S))}
    ref = oracle.fk2c_loop(u0, N0, S0, wm, gm, 0.1, 0.9, Dw, 10.0, Ds, rho, 0.35, 0.5, 0.05, dt, *h, nsteps)
    for k in "PNS":
        assert np.max(np.abs(got[k] - ref["state"][k])) <= 1e-12, k


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(shape=shapes, seed=st.integers(0, 2 ** 20), variant=st.sampled_from([1, 2, 3, 4]), nsteps=st.integers(1, 5), tun=tunables)
def test_dti_any_box_bitwise(oracle, shape, seed, variant, nsteps, tun):
    from tumorgrowthtoolkit_b200 import _native as nat
    _set_tunables(tun)
    rng = np.random.default_rng(seed)
    rgb = rng.uniform(0, 3, shape + (3,)) * (rng.uniform(0, 1, shape) < 0.8)[..., None]
    u0 = rng.uniform(0, 1, shape)
    h, Dw, rho, dt = (1.0, 1.3, 0.7), 0.9, 0.2, 0.004
    p = nat.make_params(nat.MODEL_DTI, nat.F64, shape, h, dt, rho=rho, Dw=Dw)
    with nat.Sim(p) as sim:
        sim.set_variant(variant)
        for a in range(3):
            sim.upload(nat.FIELD_COEF0 + a, rgb[..., a] * Dw)
        sim.upload(nat.FIELD_U, u0)
        sim.prepare()
        sim.run(nsteps)
        sim.sync()
        got = sim.download(nat.FIELD_U)
    ref = oracle.dti_loop(u0, rgb, Dw, rho, dt, *h, nsteps)
    assert np.array_equal(got, ref["state"])
