"""Operator-level parity of the CUDA kernels, through the C ABI, against (a) vectors produced
by the reference itself (tests/golden/ops.npz) and (b) the CPU oracle on seeded random boxes
(odd sizes, dx != dy != dz, zero tissue, tissue everywhere so the periodic wrap is active).

Tolerances: fp64 1e-12 per step for the compact kernels, EXACT (0.0) for the kernels that
keep the reference's rounding order; fp32 1e-5 per step.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = ("D_plus_x", "D_plus_y", "D_plus_z", "D_minus_x", "D_minus_y", "D_minus_z")


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as ge
    ge.build()
    from tumorgrowthtoolkit_b200 import _native
    assert _native.device_count() >= 1
    return _native


def _maxerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


def test_get_D_and_m_tildas_bitwise(L, gold_ops):
    from tumorgrowthtoolkit_b200.FK import Solver
    from tumorgrowthtoolkit_b200.FK_2c import Solver as S2
    g = gold_ops
    for ci in range(int(g["ncases"])):
        t = f"c{ci}"
        th, Dw, ratio, rho, dt = g[t + "_fk_sc"]
        D = Solver({}).get_D(g[t + "_wm"], g[t + "_gm"], th, Dw, ratio)
        for k in KEYS:
            assert np.array_equal(D[k], g[f"{t}_fk_{k}"]), (ci, k)
        th, th_necro, Dw, ratio, Ds = g[t + "_2c_sc"][:5]
        D = S2({}).get_D_with_necro(g[t + "_wm"], g[t + "_gm"], th, Dw, ratio, g[t + "_2c_P0"], g[t + "_2c_N0"], th_necro)
        for k in KEYS:
            assert np.array_equal(D[k], g[f"{t}_2c_{k}"]), (ci, k)
        M = Solver({}).m_Tildas(g[t + "_wm"], g[t + "_gm"], th)
        # D_minus = Dw*(WM_t + GM_t/ratio) recomposed from the tildas must equal get_D
        th, Dw, ratio, rho, dt = g[t + "_fk_sc"]
        assert np.array_equal(Dw * (M["WM_t_x"] + M["GM_t_x"] / ratio), g[f"{t}_fk_D_minus_x"])


def test_fk_update_method_bitwise_fp64(L, gold_ops):
    """Public FK_update(A, D_domain, ...) with explicit coefficients: in place, same object,
    bit-identical to numpy."""
    from tumorgrowthtoolkit_b200.FK import Solver
    g = gold_ops
    for ci in range(int(g["ncases"])):
        t = f"c{ci}"
        th, Dw, ratio, rho, dt = g[t + "_fk_sc"]
        dx, dy, dz = g[t + "_h"]
        D = {k: g[f"{t}_fk_{k}"] for k in KEYS}
        u = g[t + "_fk_u0"].copy()
        for st in range(3):
            r = Solver({}).FK_update(u, D, rho, dt, dx, dy, dz)
            assert r is u
            assert np.array_equal(u, g[f"{t}_fk_u{st + 1}"]), (ci, st, _maxerr(u, g[f"{t}_fk_u{st + 1}"]))


def test_fk_update_accepts_strided_views(L, gold_ops):
    from tumorgrowthtoolkit_b200.FK import Solver
    g = gold_ops
    t = "c1"
    th, Dw, ratio, rho, dt = g[t + "_fk_sc"]
    dx, dy, dz = g[t + "_h"]
    shp = g[t + "_fk_u0"].shape
    big = np.zeros(tuple(s + 5 for s in shp))
    view = big[2:2 + shp[0], 1:1 + shp[1], 3:3 + shp[2]]
    view[...] = g[t + "_fk_u0"]
    D = {}
    for k in KEYS:
        b = np.zeros(tuple(s + 3 for s in shp))
        b[1:1 + shp[0], 2:2 + shp[1], 0:shp[2]] = g[f"{t}_fk_{k}"]
        D[k] = b[1:1 + shp[0], 2:2 + shp[1], 0:shp[2]]
    Solver({}).FK_update(view, D, rho, dt, dx, dy, dz)
    assert np.array_equal(view, g[t + "_fk_u1"])
    assert big[0].sum() == 0 and big[:, 0].sum() == 0          # nothing written outside the view
    tr = np.ascontiguousarray(g[t + "_fk_u0"].transpose(2, 0, 1)).transpose(1, 2, 0)   # sz != 1
    assert tr.strides[2] != tr.itemsize
    Solver({}).FK_update(tr, {k: g[f"{t}_fk_{k}"] for k in KEYS}, rho, dt, dx, dy, dz)
    assert np.array_equal(tr, g[t + "_fk_u1"])


def test_fk2c_update_method_fp64(L, gold_ops):
    from tumorgrowthtoolkit_b200.FK_2c import Solver
    g = gold_ops
    for ci in range(int(g["ncases"])):
        t = f"c{ci}"
        th, th_necro, Dw, ratio, Ds, rho, lnp, sig, ls, dt = g[t + "_2c_sc"]
        dx, dy, dz = g[t + "_h"]
        D = {k: g[f"{t}_2c_{k}"] for k in KEYS}
        E = {k: g[f"{t}_2cS_{k}"] for k in KEYS}
        P, N, S = g[t + "_2c_P0"].copy(), g[t + "_2c_N0"].copy(), g[t + "_2c_S0"].copy()
        out = Solver({}).FK_2c_update(P, N, S, D, rho, lnp, sig, E, ls, dt, dx, dy, dz)
        assert out["P"] is P and out["N"] is N and out["S"] is S
        for a, name in ((P, "P"), (N, "N"), (S, "S")):
            assert _maxerr(a, g[f"{t}_2c_{name}1"]) <= 1e-14, (ci, name)      # exp() ulp only


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-12), ("fp32", 1e-5)])
def test_fk_loop_three_steps(L, gold_ops, precision, tol):
    from tumorgrowthtoolkit_b200 import _loop
    g = gold_ops
    for ci in range(int(g["ncases"])):
        t = f"c{ci}"
        th, Dw, ratio, rho, dt = g[t + "_fk_sc"]
        dx, dy, dz = g[t + "_h"]
        res = _loop.fk_loop(g[t + "_fk_u0"], g[t + "_wm"], g[t + "_gm"], th, Dw, ratio, rho, dt, dx, dy, dz, 3,
                            record_steps=[0, 1, 2], precision=precision)
        assert res["steps_run"] == 3 and len(res["snapshots"]) == 3
        for st in range(3):
            assert _maxerr(res["snapshots"][st], g[f"{t}_fk_u{st + 1}"]) <= tol, (ci, st)
        vol = dx * dy * dz * np.sum(g[t + "_fk_u3"])
        assert abs(res["volume"] - vol) <= (1e-12 if precision == "fp64" else 1e-5) * max(1.0, vol)


def test_fk_loop_bit_exact_mode(L, gold_ops):
    from tumorgrowthtoolkit_b200 import _loop
    g = gold_ops
    for ci in range(int(g["ncases"])):
        t = f"c{ci}"
        th, Dw, ratio, rho, dt = g[t + "_fk_sc"]
        dx, dy, dz = g[t + "_h"]
        res = _loop.fk_loop(g[t + "_fk_u0"], g[t + "_wm"], g[t + "_gm"], th, Dw, ratio, rho, dt, dx, dy, dz, 3,
                            record_steps=[0, 1, 2], bit_exact=True)
        for st in range(3):
            assert np.array_equal(res["snapshots"][st], g[f"{t}_fk_u{st + 1}"]), (ci, st)


@pytest.mark.parametrize("precision,tol", [("fp64", 0.0), ("fp32", 1e-5)])
def test_dti_loop_three_steps(L, gold_ops, precision, tol):
    from tumorgrowthtoolkit_b200 import _loop
    g = gold_ops
    for ci in range(int(g["ncases"])):
        t = f"c{ci}"
        Dw, rho, dt = g[t + "_dti_sc"]
        dx, dy, dz = g[t + "_h"]
        res = _loop.dti_loop(g[t + "_dti_u0"], g[t + "_dti_rgb"], Dw, rho, dt, dx, dy, dz, 3,
                             record_steps=[0, 1, 2], precision=precision)
        for st in range(3):
            assert _maxerr(res["snapshots"][st], g[f"{t}_dti_u{st + 1}"]) <= tol, (ci, st)


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-12), ("fp32", 2e-5)])
def test_fk2c_loop_three_steps(L, gold_ops, precision, tol):
    from tumorgrowthtoolkit_b200 import _loop
    g = gold_ops
    for ci in range(int(g["ncases"])):
        t = f"c{ci}"
        th, th_necro, Dw, ratio, Ds, rho, lnp, sig, ls, dt = g[t + "_2c_sc"]
        dx, dy, dz = g[t + "_h"]
        res = _loop.fk2c_loop(g[t + "_2c_P0"], g[t + "_2c_N0"], g[t + "_2c_S0"], g[t + "_wm"], g[t + "_gm"],
                              th, th_necro, Dw, ratio, Ds, rho, lnp, sig, ls, dt, dx, dy, dz, 3,
                              record_steps=[0, 1, 2], precision=precision)
        for st in range(3):
            for name in "PNS":
                e = _maxerr(res["snapshots"][name][st], g[f"{t}_2c_{name}{st + 1}"])
                assert e <= tol, (ci, st, name, e)


SHAPES = [(1, 1, 1), (1, 7, 3), (2, 2, 2), (3, 1, 9), (33, 29, 31), (17, 40, 65), (64, 5, 130)]


@pytest.mark.parametrize("shape", SHAPES)
def test_random_boxes_against_oracle(L, oracle, shape):
    """Seeded fuzz incl. degenerate boxes (size-1 and size-2 axes: a voxel is its own /
    twice its neighbour under the periodic wrap), all three models, 5 steps."""
    from tumorgrowthtoolkit_b200 import _loop
    rng = np.random.default_rng(abs(hash(shape)) % (2 ** 31))
    wm = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.8)
    gm = rng.uniform(0, 1, shape) * (1 - wm)
    u0 = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.5)
    h = (1.0, 1.3, 0.8)
    dt = 0.02
    ref = oracle.fk_loop(u0, wm, gm, 0.1, 1.1, 10.0, 0.2, dt, *h, 5)
    got = _loop.fk_loop(u0, wm, gm, 0.1, 1.1, 10.0, 0.2, dt, *h, 5)
    assert _maxerr(got["state"], ref["state"]) <= 1e-12
    got = _loop.fk_loop(u0, wm, gm, 0.1, 1.1, 10.0, 0.2, dt, *h, 5, bit_exact=True)
    assert np.array_equal(got["state"], ref["state"])
    rgb = rng.uniform(0, 3, shape + (3,))
    ref = oracle.dti_loop(u0, rgb, 0.9, 0.2, dt / 3, *h, 5)
    got = _loop.dti_loop(u0, rgb, 0.9, 0.2, dt / 3, *h, 5)
    assert np.array_equal(got["state"], ref["state"])
    N0 = rng.uniform(0, 0.6, shape) * (rng.uniform(0, 1, shape) < 0.5)
    S0 = rng.uniform(0.2, 1, shape)
    a2 = (u0, N0, S0, wm, gm, 0.1, 0.9, 1.1, 10.0, 1.3, 0.2, 0.35, 0.5, 0.05, dt, *h, 5)
    ref = oracle.fk2c_loop(*a2)
    got = _loop.fk2c_loop(*a2)
    for k in "PNS":
        assert _maxerr(got["state"][k], ref["state"][k]) <= 1e-12, k
    assert abs(got["volume"] - ref["volume"]) <= 1e-10 * max(1.0, abs(ref["volume"]))


def test_bool_tissue_maps_follow_numpy_logical_or(L):
    """numpy evaluates bool WM + roll(WM) as logical OR: an all-WM face gets
    Dw*0.5.  The reference semantics restated with numpy ufuncs on bool arrays."""
    from tumorgrowthtoolkit_b200.FK import Solver
    rng = np.random.default_rng(5)
    shp = (6, 7, 5)
    lab = rng.integers(0, 4, shp)
    wm, gm = lab == 3, lab == 2
    th, Dw, ratio = 0.1, 1.3, 10.0
    D = Solver({}).get_D(wm, gm, th, Dw, ratio)
    for a, name in enumerate("xyz"):
        cond = np.logical_and(np.roll(wm, -1, a) + np.roll(gm, -1, a) >= th, wm + gm >= th)
        wt = np.where(cond, (np.roll(wm, -1, a) + wm) / 2, 0)
        gt = np.where(cond, (np.roll(gm, -1, a) + gm) / 2, 0)
        assert np.array_equal(D["D_minus_" + name], Dw * (wt + gt / ratio))


def test_stop_on_volume_and_snapshots(L, oracle):
    from tumorgrowthtoolkit_b200 import _loop
    rng = np.random.default_rng(11)
    shp = (20, 18, 22)
    wm = np.full(shp, 0.7)
    gm = np.full(shp, 0.3)
    u0 = np.zeros(shp)
    u0[8:12, 8:12, 8:12] = 0.5
    args = (u0, wm, gm, 0.1, 1.0, 10.0, 0.8, 0.05, 1.0, 1.0, 1.0, 200)
    free = oracle.fk_loop(*args)
    target = 0.5 * (free["volume"] + 32.0)
    rec = np.linspace(0, 199, 17, dtype=int)
    ref = oracle.fk_loop(*args, stopping_volume=target, record_steps=rec)
    got = _loop.fk_loop(*args, stopping_volume=target, record_steps=rec)
    assert ref["stop"] == "volume" and got["stop"] == "volume"
    assert got["t_stop"] == ref["t_stop"] and got["steps_run"] == ref["steps_run"]
    assert len(got["snapshots"]) == len(ref["snapshots"]) > 0
    assert _maxerr(got["state"], ref["state"]) <= 1e-12
    assert _maxerr(got["snapshots"][-1], ref["snapshots"][-1]) <= 1e-12
    assert abs(got["volume"] - ref["volume"]) <= 1e-10 * ref["volume"]


def test_volume_log_matches_per_step_sums(L, oracle):
    from tumorgrowthtoolkit_b200 import _native as nat
    shp = (12, 14, 10)
    rng = np.random.default_rng(2)
    wm, gm = rng.uniform(0.2, 0.6, shp), rng.uniform(0.1, 0.4, shp)
    u0 = rng.uniform(0, 0.3, shp)
    p = nat.make_params(nat.MODEL_FK, nat.F64, shp, (1.0, 1.0, 2.0), 0.03, rho=0.3, Dw=1.0, ratio=10.0, th_matter=0.1)
    with nat.Sim(p) as sim:
        sim.upload(nat.FIELD_WM, wm)
        sim.upload(nat.FIELD_GM, gm)
        sim.upload(nat.FIELD_U, u0)
        sim.prepare()
        sim.run(4)
        sim.run(3)          # a second run continues the same loop
        st = sim.sync()
        vols = sim.volumes(0, 7)
    assert st.steps_run == 7 and st.kernel_launches == 7
    u = u0
    Dm = oracle.faces(wm, gm, 0.1, 1.0, 10.0)
    for t in range(7):
        u = oracle.fk_step_faces(u, Dm, 0.3, 0.03, 1.0, 1.0, 2.0)
        assert abs(vols[t] - 2.0 * u.sum()) <= 1e-11 * max(1.0, 2.0 * u.sum())


def test_errors_are_reported_not_swallowed(L):
    from tumorgrowthtoolkit_b200 import _native as nat
    p = nat.make_params(nat.MODEL_FK, nat.F64, (4, 4, 4), (1, 1, 1), 0.1)
    with nat.Sim(p) as sim:
        with pytest.raises(nat.NativeError):
            sim.run(1)                       # not prepared
        with pytest.raises(nat.NativeError):
            sim.prepare()                    # tissue maps missing
        with pytest.raises(ValueError):
            sim.upload(nat.FIELD_U, np.zeros((3, 4, 4)))
    with pytest.raises(nat.NativeError):
        nat.Sim(nat.make_params(99, nat.F64, (4, 4, 4), (1, 1, 1), 0.1))
    with pytest.raises(nat.NativeError):
        nat.Sim(nat.make_params(nat.MODEL_FK, nat.F64, (0, 4, 4), (1, 1, 1), 0.1))


@pytest.mark.parametrize("factor", [0.5, 0.65, 1.0, 2.0, 1.7])
def test_resample_matches_scipy_zoom(L, factor):
    """The device resample against scipy.ndimage.zoom(order=1) itself (the reference's call,
    FK/FK.py:150-151): plain zoom, strided input, and the fused restore_tumor + zoom."""
    from scipy.ndimage import zoom
    from tumorgrowthtoolkit_b200 import _host
    rng = np.random.default_rng(17)
    a = rng.uniform(0, 1, (23, 31, 18))
    ref = zoom(a, factor, order=1)
    got = _host.resample_linear(a, factor)
    assert got.shape == ref.shape and _maxerr(got, ref) <= 1e-14
    big = rng.uniform(0, 1, (30, 40, 25))
    view = big[3:26, 2:33, 5:23]
    assert _maxerr(_host.resample_linear(view, factor), zoom(view, factor, order=1)) <= 1e-14
    # crop pasted into zeros, then zoomed back to a "full" grid
    low_shape = (40, 44, 30)
    grid = _host.LowResGrid(low_shape, tuple(int(round(n / 0.65)) for n in low_shape), 0.65, 1, 1, 1, (0.5, 0.5, 0.5))
    box = (np.array([5, 6, 4]), np.array([28, 37, 22]))
    full = np.zeros(low_shape)
    full[5:28, 6:37, 4:22] = a
    ref = zoom(full, grid.up_factor, order=1)
    got = grid.upsample_crop(a, box)
    assert got.shape == ref.shape == grid.full_shape and _maxerr(got, ref) <= 1e-14


def test_slab_single_rank_equals_plain_loop(L):
    """world = 1: the slab driver (ghost planes, blocks of `halo` steps, ring closing on the rank
    itself, volume over owned planes only) must reproduce the plain loop bitwise."""
    from tumorgrowthtoolkit_b200 import _loop, slab
    rng = np.random.default_rng(9)
    shape = (19, 12, 21)
    wm = rng.uniform(0.2, 0.7, shape)
    gm = 1.0 - wm
    u0 = rng.uniform(0, 0.5, shape)
    args = (u0, wm, gm, 0.1, 1.0, 10.0, 0.3, 0.02, 1.0, 1.1, 0.9, 11)
    ref = _loop.fk_loop(*args)
    for halo, driver in ((1, "p2p"), (3, "native"), (4, "torch"), (4, "p2p")):
        got = slab.fk_loop(*args, halo=halo, driver=driver)
        assert np.array_equal(got["state"], ref["state"]), (halo, driver)
        assert abs(got["volume"] - ref["volume"]) <= 1e-12 * ref["volume"]
    P0, N0, S0 = u0, 0.3 * u0[::-1], np.ones(shape)
    a2 = (P0, N0, S0, wm, gm, 0.1, 0.9, 1.0, 10.0, 1.3, 0.3, 0.35, 0.5, 0.05, 0.02, 1.0, 1.1, 0.9, 9)
    ref = _loop.fk2c_loop(*a2)
    got = slab.fk2c_loop(*a2, halo=2, driver="p2p")
    for k in "PNS":
        assert np.array_equal(got["state"][k], ref["state"][k]), k


@pytest.mark.parametrize("key,exponent,linear", [("dti_rgb_exp1", 1, 0), ("dti_rgb_exp15", 1.5, 0.2)])
def test_rgb_from_tensor_kernels_match_reference(L, gold_solve, key, exponent, linear):
    """tools.makeXYZ_rgb_from_tensor on the device (tgtk_dti_rgb_host) against the reference's own output
    (tests/golden/solve.npz: dti_rgb_*): only the summation order of the two brain means differs."""
    from tumorgrowthtoolkit_b200 import synthetic, _ops
    T = synthetic.dti_tensor_field(gold_solve["wm"], gold_solve["gm"], seed=7)
    got = _ops.dti_rgb(T, exponent, linear)
    want = gold_solve[key]
    assert got.shape == want.shape
    assert _maxerr(got, want) <= 1e-12
    assert np.array_equal(got == 0, want == 0)
