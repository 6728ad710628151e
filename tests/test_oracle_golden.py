"""Pins the CPU oracle (oracle/) against vectors produced by the reference itself
(tests/golden/ops.npz, written by tests/make_golden.py).  CPU only."""
import numpy as np
import pytest

KEYS = ("D_plus_x", "D_plus_y", "D_plus_z", "D_minus_x", "D_minus_y", "D_minus_z")


def _cases(g):
    return range(int(g["ncases"]))


def test_faces_match_reference_get_D_bitwise(oracle, gold_ops):
    g = gold_ops
    for ci in _cases(g):
        t = f"c{ci}"
        th, Dw, ratio, rho, dt = g[t + "_fk_sc"]
        D = oracle.get_D(g[t + "_wm"], g[t + "_gm"], th, Dw, ratio)
        for k in KEYS:
            assert np.array_equal(D[k], g[f"{t}_fk_{k}"]), (ci, k)


def test_faces_with_necro_match_reference_bitwise(oracle, gold_ops):
    g = gold_ops
    for ci in _cases(g):
        t = f"c{ci}"
        th, th_necro, Dw, ratio, Ds = g[t + "_2c_sc"][:5]
        D = oracle.get_D(g[t + "_wm"], g[t + "_gm"], th, Dw, ratio, g[t + "_2c_P0"], g[t + "_2c_N0"], th_necro)
        E = oracle.get_D(g[t + "_wm"], g[t + "_gm"], th, Ds, 1)
        for k in KEYS:
            assert np.array_equal(D[k], g[f"{t}_2c_{k}"]), (ci, k)
            assert np.array_equal(E[k], g[f"{t}_2cS_{k}"]), (ci, k)


def test_fk_update_bitwise(oracle, gold_ops):
    g = gold_ops
    for ci in _cases(g):
        t = f"c{ci}"
        th, Dw, ratio, rho, dt = g[t + "_fk_sc"]
        dx, dy, dz = g[t + "_h"]
        D = {k: g[f"{t}_fk_{k}"] for k in KEYS}
        Dm = oracle.faces(g[t + "_wm"], g[t + "_gm"], th, Dw, ratio)
        u = g[t + "_fk_u0"]
        v = u.copy()
        for st in range(3):
            u = oracle.fk_step(u, D, rho, dt, dx, dy, dz)
            v = oracle.fk_step_faces(v, Dm, rho, dt, dx, dy, dz)
            assert np.array_equal(u, g[f"{t}_fk_u{st + 1}"]), (ci, st)
            assert np.array_equal(v, g[f"{t}_fk_u{st + 1}"]), (ci, st)


def test_dti_update_bitwise(oracle, gold_ops):
    g = gold_ops
    for ci in _cases(g):
        t = f"c{ci}"
        Dw, rho, dt = g[t + "_dti_sc"]
        dx, dy, dz = g[t + "_h"]
        res = oracle.dti_loop(g[t + "_dti_u0"], g[t + "_dti_rgb"], Dw, rho, dt, dx, dy, dz, 3,
                              record_steps=[0, 1, 2])
        for st in range(3):
            assert np.array_equal(res["snapshots"][st], g[f"{t}_dti_u{st + 1}"]), (ci, st)


def test_fk2c_update_explicit_coefficients(oracle, gold_ops):
    """FK_2c_update with the reference's own coefficient dicts: differs from numpy only by
    libm exp vs numpy's SIMD exp (<= 1 ulp of exp), tolerance 1e-14."""
    g = gold_ops
    for ci in _cases(g):
        t = f"c{ci}"
        th, th_necro, Dw, ratio, Ds, rho, lnp, sig, ls, dt = g[t + "_2c_sc"]
        dx, dy, dz = g[t + "_h"]
        D = {k: g[f"{t}_2c_{k}"] for k in KEYS}
        E = {k: g[f"{t}_2cS_{k}"] for k in KEYS}
        P, N, S = oracle.fk2c_step(g[t + "_2c_P0"], g[t + "_2c_N0"], g[t + "_2c_S0"], D, E, rho, lnp, sig, ls,
                                   dt, dx, dy, dz)
        for a, name in ((P, "P"), (N, "N"), (S, "S")):
            assert np.max(np.abs(a - g[f"{t}_2c_{name}1"])) <= 1e-14, (ci, name)


def test_fk2c_loop_three_steps(oracle, gold_ops):
    """Three steps incl. the per-step rebuild of the necro-masked faces (FK_2c.py:229-230)."""
    g = gold_ops
    for ci in _cases(g):
        t = f"c{ci}"
        th, th_necro, Dw, ratio, Ds, rho, lnp, sig, ls, dt = g[t + "_2c_sc"]
        dx, dy, dz = g[t + "_h"]
        res = oracle.fk2c_loop(g[t + "_2c_P0"], g[t + "_2c_N0"], g[t + "_2c_S0"], g[t + "_wm"], g[t + "_gm"],
                               th, th_necro, Dw, ratio, Ds, rho, lnp, sig, ls, dt, dx, dy, dz, 3,
                               record_steps=[0, 1, 2])
        for st in range(3):
            for name in "PNS":
                d = np.max(np.abs(res["snapshots"][name][st] - g[f"{t}_2c_{name}{st + 1}"]))
                assert d <= 1e-13, (ci, st, name, d)


def test_zoom_linear_matches_scipy(oracle):
    from scipy.ndimage import zoom
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 1, (13, 10, 9))
    for f in (0.5, 0.65, 1.0, 2.0):
        ref = zoom(a, f, order=1)
        out = oracle.zoom_linear(a, ref.shape)
        assert np.max(np.abs(out - ref)) <= 1e-14


def test_atlas_loops_match_reference(oracle, gold_atlas, atlas_labels):
    """Oracle loops on the SRI atlas at resolution_factor 0.5 against the reference's own
    loop output: FK 300 steps (bitwise), FK_2c 560 steps (<= 1e-10, exp ulp only)."""
    from scipy.ndimage import zoom
    seg = atlas_labels
    wm, gm = (seg == 3).astype(np.float64), (seg == 2).astype(np.float64)
    wml, gml = zoom(wm, 0.5, order=1), zoom(gm, 0.5, order=1)
    g = gold_atlas
    mn, mx = g["fk/crop"]
    sl = tuple(slice(int(a), int(b)) for a, b in zip(mn, mx))
    cw, cg = np.ascontiguousarray(wml[sl]), np.ascontiguousarray(gml[sl])
    Nx, Ny, Nz = wml.shape
    dx = 2.0
    x, y, z = np.meshgrid(np.arange(Nx) - int(0.3 * Nx), np.arange(Ny) - int(0.7 * Ny),
                          np.arange(Nz) - int(0.5 * Nz), indexing="ij")
    A = 250 / np.power(4 * np.pi * 5.0, 3 / 2) * np.exp(-((x * dx) ** 2 + (y * dx) ** 2 + (z * dx) ** 2) / 20.0)
    A = np.where(A > 0.1, A, 0)
    A = np.where(A > 1, 1.0, A)[sl]
    steps = int(g["fk/steps"])
    Nt = 100 * 1.0 / dx ** 2 * 8 + 100
    res = oracle.fk_loop(A, cw, cg, 0.1, 1.0, 10, 0.10, 100 / Nt, dx, dx, dx, steps)
    assert res["steps_run"] == 300
    assert np.max(np.abs(res["state"] - g["fk/final_lowres_crop"])) <= 1e-13
    assert abs(res["volume"] - float(g["fk/final_volume"])) <= 1e-9 * float(g["fk/final_volume"])

    steps = int(g["2c/steps"])
    Nt = 100 * 1.3 / dx ** 2 * 8 + 300
    S0 = np.where(cw + cg >= 0.1, 1.0, 0.0)
    res = oracle.fk2c_loop(A, np.zeros_like(A), S0, cw, cg, 0.1, 0.9, 0.9, 100, 1.3, 0.14, 0.35, 0.5, 0.05,
                           100 / Nt, dx, dx, dx, steps)
    for k in "PNS":
        d = np.max(np.abs(res["state"][k] - g[f"2c/final_lowres_crop_{k}"]))
        assert d <= 1e-10, (k, d)
    assert abs(res["volume"] - float(g["2c/final_volume"])) <= 1e-9 * float(g["2c/final_volume"])
