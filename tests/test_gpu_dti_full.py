"""Opt-in full-tensor FK_DTI operator (19-point).  The packaged reference has no such mode
(SURVEY.md finding 1), so there is no reference parity: the checks are (i) reduction to the
reference's diagonal operator bit for bit, (ii) a numpy restatement with np.roll, (iii) a
closed-form discrete solution."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


def numpy_full_step(u, D, off, rho, dt, h):
    """u' = u + dt*(sum_a D_a d2_a u + 2 sum_{a<b} D_ab d_a d_b u + rho u (1-u)), periodic."""
    sp = 0
    terms = []
    for a in range(3):
        t = 1 / (h[a] * h[a]) * (D[..., a] * (np.roll(u, 1, a) - u) - D[..., a] * (u - np.roll(u, -1, a)))
        terms.append(t)
    sp = (terms[0] + terms[1]) + terms[2]
    cross = []
    for idx, (a, b) in enumerate(((0, 1), (0, 2), (1, 2))):
        upp = np.roll(np.roll(u, -1, a), -1, b)
        upm = np.roll(np.roll(u, -1, a), 1, b)
        ump = np.roll(np.roll(u, 1, a), -1, b)
        umm = np.roll(np.roll(u, 1, a), 1, b)
        m = (upp - upm) - (ump - umm)
        cross.append((2.0 / (4.0 * h[a] * h[b]) * off[..., idx]) * m)
    cr = (cross[0] + cross[1]) + cross[2]
    return u + ((sp + cr) + rho * (u * (1 - u))) * dt


def test_zero_offdiagonals_reduce_to_reference_operator_bitwise(oracle):
    from tumorgrowthtoolkit_b200 import _loop
    rng = np.random.default_rng(1)
    shape = (11, 9, 13)
    rgb = rng.uniform(0, 3, shape + (3,))
    u0 = rng.uniform(0, 1, shape)
    args = (0.8, 0.2, 0.01, 1.0, 1.3, 0.7, 4)
    ref = oracle.dti_loop(u0, rgb, *args)                       # the pinned reference operator
    got = _loop.dti_full_loop(u0, rgb, np.zeros(shape + (3,)), *args)
    assert np.array_equal(got["state"], ref["state"])


def test_against_numpy_restatement():
    from tumorgrowthtoolkit_b200 import _loop
    rng = np.random.default_rng(2)
    shape = (10, 12, 7)
    rgb = rng.uniform(0.5, 2, shape + (3,))
    off = rng.uniform(-0.4, 0.4, shape + (3,))
    u = rng.uniform(0, 1, shape)
    h, Dw, rho, dt = (1.0, 1.3, 0.7), 0.9, 0.3, 0.01
    got = _loop.dti_full_loop(u, rgb, off, Dw, rho, dt, *h, 3, record_steps=[0, 1, 2])
    for st in range(3):
        u = numpy_full_step(u, rgb * Dw, off * Dw, rho, dt, h)
        assert np.max(np.abs(got["snapshots"][st] - u)) <= 1e-14


def test_diffusion_only_along_the_tensor_axis_closed_form():
    """D = lam * e e^T with e = (1,1,0)/sqrt(2) and a field that varies only across e,
    u = sin(2 pi (i-j)/M): the continuous operator gives exactly zero flux; the discrete 19-point
    operator gives u' = u (1 - 4 dt lam sin^4(pi/M)) (derivation in DESIGN.md / the docstring of
    dti_full_step_v1)."""
    from tumorgrowthtoolkit_b200 import _loop
    M, lam, dt = 24, 1.7, 0.05
    i, j, k = np.meshgrid(np.arange(M), np.arange(M), np.arange(5), indexing="ij")
    u0 = 0.4 * np.sin(2 * np.pi * (i - j) / M)
    rgb = np.zeros((M, M, 5, 3))
    rgb[..., 0] = rgb[..., 1] = lam / 2
    off = np.zeros((M, M, 5, 3))
    off[..., 0] = lam / 2
    got = _loop.dti_full_loop(u0, rgb, off, 1.0, 0.0, dt, 1.0, 1.0, 1.0, 1)
    want = u0 * (1 - 4 * dt * lam * np.sin(np.pi / M) ** 4)
    assert np.max(np.abs(got["state"] - want)) <= 1e-14
    # the diagonal-only operator would smooth the same field ~1/sin^2(pi/M) = 58x faster
    diag = _loop.dti_loop(u0, rgb, 1.0, 0.0, dt, 1.0, 1.0, 1.0, 1)
    assert np.max(np.abs(diag["state"] - u0)) > 20 * np.max(np.abs(got["state"] - u0))


def test_solver_option(gold_solve):
    from tumorgrowthtoolkit_b200.FK_DTI import FK_DTI_Solver
    from tests.solve_cases import DTI_CASES, dti_params
    p = dti_params(gold_solve, DTI_CASES["dti_a"])
    base = FK_DTI_Solver(dict(p)).solve()
    p["dti_full_tensor"] = True
    full = FK_DTI_Solver(p).solve()
    assert full["success"] and full["final_state"].shape == base["final_state"].shape
    d = np.max(np.abs(full["final_state"] - base["final_state"]))
    assert 0 < d < 0.5                      # off-diagonals change the result, moderately


def test_staged_kernel_equals_baseline_kernel_bitwise_any_box():
    """dti_full_step_v3 (marching, four-plane shared-memory window; the default) against dti_full_step_v1 (one voxel per
    thread, neighbours through L1) on ragged and degenerate boxes, fp64 and fp32."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from tumorgrowthtoolkit_b200 import _native as nat

    @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(shape=st.tuples(st.integers(1, 20), st.integers(1, 30), st.integers(1, 70)), seed=st.integers(0, 2 ** 20),
           nsteps=st.integers(1, 5), dtype=st.sampled_from([nat.F64, nat.F32]))
    def check(shape, seed, nsteps, dtype):
        rng = np.random.default_rng(seed)
        coefs = [rng.uniform(0.2, 2.0, shape) for _ in range(3)] + [rng.uniform(-0.4, 0.4, shape) for _ in range(3)]
        u0 = rng.uniform(0, 1, shape)
        out = {}
        for variant in (1, 0):
            p = nat.make_params(nat.MODEL_DTI_FULL, dtype, shape, (1.0, 1.3, 0.7), 0.004, rho=0.2, Dw=1.0)
            with nat.Sim(p) as sim:
                sim.set_variant(variant)
                for c, arr in enumerate(coefs):
                    sim.upload(nat.FIELD_COEF0 + c, arr)
                sim.upload(nat.FIELD_U, u0)
                sim.prepare()
                sim.run(nsteps)
                st_ = sim.sync()
                out[variant] = (sim.download(nat.FIELD_U), st_.last_volume, sim.tuning()["planes_per_segment"])
        assert out[1][2] == 0 and out[0][2] > 0          # baseline vs marching geometry
        assert np.array_equal(out[0][0], out[1][0])
        assert abs(out[0][1] - out[1][1]) <= 1e-10 * max(1.0, abs(out[1][1]))

    check()
