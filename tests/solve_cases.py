"""Parameter sets of tests/golden/solve.npz (must match tests/make_golden.py:gen_solve) and
the comparison of a result dict with the stored reference result."""
import numpy as np

SEED = dict(NxT1_pct=0.35, NyT1_pct=0.6, NzT1_pct=0.5)

FK_CASES = {
    "fk_a": dict(Dw=1.0, rho=0.10, RatioDw_Dg=10, th_matter=0.1, init_scale=1.0,
                 resolution_factor=0.5, time_series_solution_Nt=5, stopping_time=40),
    "fk_b": dict(Dw=0.6, rho=0.25, resolution_factor=1.0, stopping_time=12, dx_mm=1.0, dy_mm=1.2, dz_mm=0.9),
    "fk_vol": dict(Dw=1.5, rho=0.3, resolution_factor=0.65, stopping_time=60, stopping_volume=900.0,
                   time_series_solution_Nt=7),
    "fk_outside": dict(Dw=1.0, rho=0.1, resolution_factor=0.5, NxT1_pct=0.02, NyT1_pct=0.02, NzT1_pct=0.02),
    "fk_ts_gt_steps": dict(Dw=0.05, rho=0.02, resolution_factor=0.5, stopping_time=2, time_series_solution_Nt=200),
}
C2_CASES = {
    "2c_a": dict(Dw=0.9, rho=0.14, lambda_np=0.35, sigma_np=0.5, D_s=1.3, lambda_s=0.05, RatioDw_Dg=100,
                 Nt_multiplier=8, resolution_factor=0.5, time_series_solution_Nt=4, stopping_time=60),
    "2c_b": dict(Dw=2.0, rho=0.5, lambda_np=0.6, sigma_np=0.3, D_s=0.8, lambda_s=0.3, resolution_factor=1.0,
                 Nt_multiplier=10, stopping_time=15, th_necro=0.8),
    "2c_vol": dict(Dw=1.2, rho=0.4, lambda_np=0.2, sigma_np=0.4, D_s=2.0, lambda_s=0.1,
                   resolution_factor=0.65, stopping_volume=1500.0, stopping_time=80),
}
DTI_CASES = {
    "dti_a": dict(Dw=1.0, rho=0.2, init_scale=1.0, diffusionEllipsoidScaling=1, resolution_factor=0.5,
                  stopping_time=20, time_series_solution_Nt=3),
    "dti_vol": dict(Dw=1.0, rho=0.4, init_scale=0.5, diffusionEllipsoidScaling=1, resolution_factor=1.0,
                    stopping_time=100, stopping_volume=700.0, diffusionTensorExponent=1.5,
                    diffusionTensorLinear=0.2),
    "dti_outside": dict(Dw=1.0, rho=0.2, diffusionEllipsoidScaling=1, resolution_factor=0.5,
                        NxT1_pct=0.02, NyT1_pct=0.02, NzT1_pct=0.02, stopping_time=5),
    "dti_small": dict(Dw=3.0, rho=0.0, init_scale=0.2, diffusionEllipsoidScaling=1, resolution_factor=1.0,
                      stopping_time=3),
    "dti_decay": dict(Dw=1.0, rho=-2.0, init_scale=1.0, diffusionEllipsoidScaling=1, resolution_factor=0.5,
                      stopping_time=30, time_series_solution_Nt=6),
    "dti_elong": dict(Dw=1.0, rho=0.2, init_scale=1.0, diffusionEllipsoidScaling=2.0,
                      resolution_factor=0.5, stopping_time=10),
}


def base_params(gold, extra):
    p = dict(gm=gold["gm"].copy(), wm=gold["wm"].copy(), **SEED)
    p.update(extra)
    return p


def dti_params(gold, extra):
    from tumorgrowthtoolkit_b200 import synthetic
    T = synthetic.dti_tensor_field(gold["wm"], gold["gm"], seed=7)
    p = dict(diffusionTensors=T, **SEED)
    p.update(extra)
    return p


def series_digest(series):
    series = np.asarray(series, dtype=np.float64)
    if series.size == 0:
        return np.zeros((0, 3))
    w = np.cos(0.37 * np.arange(series[0].size)).reshape(series[0].shape)
    return np.array([[s.sum(), (s * s).sum(), (s * w).sum()] for s in series])


def _cmp_array(key, got, want, tol):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (key, got.shape, want.shape)
    if want.size:
        err = float(np.max(np.abs(got - want)))
        assert err <= tol, (key, err)


def _cmp_series(key, got, gold, tol):
    dg = gold[key + "/digest"]
    assert len(got) == len(dg), (key, len(got), len(dg))
    if len(dg):
        _cmp_array(key + "/first", got[0], gold[key + "/first"], tol)
        _cmp_array(key + "/last", got[-1], gold[key + "/last"], tol)
        d = series_digest(got)
        scale = np.maximum(np.abs(dg), 1.0)
        assert np.max(np.abs(d - dg) / scale) <= tol * 1e3, key


def compare_result(name, res, gold, tol, vol_rtol=1e-9):
    keys = sorted(k[len(name) + 1:] for k in gold.files if k.startswith(name + "/"))
    top = sorted({k.split("/")[0] for k in keys})
    assert sorted(res.keys()) == top, (sorted(res.keys()), top)
    for k in top:
        v = res[k]
        if k == "time_series":
            if f"{name}/time_series" in gold.files:          # stored as the "__none__" marker
                assert v is None
            elif isinstance(v, dict):
                for f in "PNS":
                    _cmp_series(f"{name}/time_series/{f}", v[f], gold, tol)
            else:
                _cmp_series(f"{name}/time_series", v, gold, tol)
        elif isinstance(v, dict):
            for f, arr in v.items():
                _cmp_array(f"{name}/{k}/{f}", arr, gold[f"{name}/{k}/{f}"], tol)
        else:
            want = gold[f"{name}/{k}"]
            if want.dtype.kind in "US":
                assert str(v) == str(want), (k, v, want)
            elif want.dtype.kind == "b":
                assert bool(v) == bool(want), (k, v, want)
            elif want.ndim == 0:
                assert abs(float(v) - float(want)) <= vol_rtol * max(1.0, abs(float(want))), (k, v, want)
            else:
                _cmp_array(f"{name}/{k}", v, want, tol)


# ---- SURVEY.md 8(d) input 3: FK_DTI on the synthetic tensor volume over the SRI atlas (tests/golden/atlas_dti_res05.npz) ----
ATLAS_DTI_CASES = {"stop": dict(init_scale=0.1, stopping_volume=15000, stopping_time=1000),
                   "fixed": dict(init_scale=1.0, stopping_time=100)}
_atlas_dti_tensors = {}


def atlas_dti_params(labels, gold, case):
    """Parameters of tests/make_golden.py:gen_dti_atlas (the synthetic tensor volume is regenerated, seeded)."""
    import contextlib, io
    from tumorgrowthtoolkit_b200 import synthetic
    from tumorgrowthtoolkit_b200.FK_DTI import tools
    if "T" not in _atlas_dti_tensors:
        with contextlib.redirect_stdout(io.StringIO()):
            _atlas_dti_tensors["T"] = tools.get_tensor_from_lower6(synthetic.dti_lower6_field(labels))
    pct = synthetic.seed_inside(labels == 3, (0.49, 0.30, 0.7))
    assert np.allclose(pct, gold["seed_pct"])
    p = dict(Dw=1.0, rho=0.2, diffusionEllipsoidScaling=1, diffusionTensors=_atlas_dti_tensors["T"], NxT1_pct=pct[0], NyT1_pct=pct[1],
             NzT1_pct=pct[2], resolution_factor=0.5)
    p.update(ATLAS_DTI_CASES[case])
    return p


def compare_atlas_dti(case, res, gold, tol):
    assert res["success"], res.get("error")
    assert str(res["stopping_criteria"]) == str(gold[case + "/stopping_criteria"])
    # same stop step; dt = T / Nt carries the ~1e-16 relative difference of max(rgb) (device vs numpy summation order of the brain means)
    ft = float(gold[case + "/final_time"])
    assert abs(float(res["final_time"]) - ft) <= 1e-12 * max(1.0, ft), (res["final_time"], ft)
    fv = float(gold[case + "/final_volume"])
    assert abs(float(res["final_volume"]) - fv) <= max(tol, 1e-9) * fv
    fs = res["final_state"]
    assert fs.shape == (240, 240, 155)
    err = float(np.max(np.abs(fs[::2, ::2, ::2] - gold[case + "/final_state_sub2"])))
    assert err <= tol, err
    w = np.cos(0.37 * np.arange(fs.size)).reshape(fs.shape)
    dg = gold[case + "/final_state_digest"]
    d = np.array([fs.sum(), (fs * fs).sum(), (fs * w).sum()])
    assert np.max(np.abs(d - dg) / np.maximum(np.abs(dg), 1.0)) <= tol * 1e3
    ini = res["initial_state"]
    di = gold[case + "/initial_state_digest"]
    assert abs(ini.sum() - di[0]) <= 1e-9 * max(1.0, di[0]) and abs((ini ** 2).sum() - di[1]) <= 1e-9 * max(1.0, di[1])
