"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (unmodified, imported from
/root/reference).  Run once in the development container:

    python tests/make_golden.py            # all groups
    python tests/make_golden.py ops solve  # selected groups

The reference cannot travel to the GPU box, so its outputs are committed as fixtures next
to this script.  Nothing here is imported by the product.

Groups
  ops    operator-level: get_D / FK_update / get_D_with_necro / FK_2c_update /
         get_D_from_DTI on small random boxes (odd sizes, dx != dy != dz, zero tissue,
         a tissue-everywhere case that exercises the periodic wrap), 3 steps each.
  solve  whole solve() calls of the three solvers on a small synthetic phantom
         (result dict arrays, step counts, stop criteria; finite stopping_volume; seed
         outside the brain; time series).
  dti_atlas  FK_DTI_Solver.solve() of the reference on the synthetic tensor volume of SURVEY.md 8(d) input 3 at res 0.5.
  atlas  the bundled SRI label atlas (240x240x155) packed to uint8, and reference loop
         results on it at resolution_factor 0.5 (FK 300 steps, FK_2c 560 steps) obtained by
         calling the reference's own update methods on the reference's own cropped arrays.
"""
import gzip
import os
import struct
import sys
import time
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(HERE))


def import_reference():
    for m in ("matplotlib", "matplotlib.pyplot", "nibabel", "dipy", "dipy.core", "dipy.core.gradients",
              "dipy.data", "dipy.io", "dipy.io.image", "dipy.reconst", "dipy.reconst.dti",
              "dipy.segment", "dipy.segment.mask"):
        sys.modules.setdefault(m, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from TumorGrowthToolkit.FK import Solver as FK
    from TumorGrowthToolkit.FK_2c import Solver as FK2c
    from TumorGrowthToolkit.FK_DTI import FK_DTI_Solver as DTI
    import TumorGrowthToolkit.FK_DTI.tools as tools
    return FK, FK2c, DTI, tools


def read_nifti(path):
    """Minimal NIfTI-1 reader (gzip + struct): data in Fortran order, scl_slope/inter applied."""
    raw = gzip.open(path, "rb").read()
    dim = struct.unpack("<8h", raw[40:56])
    datatype, = struct.unpack("<h", raw[70:72])
    vox_offset, slope, inter = struct.unpack("<3f", raw[108:120])
    dt = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 512: np.uint16}[datatype]
    shape = tuple(dim[1:1 + dim[0]])
    n = int(np.prod(shape))
    a = np.frombuffer(raw, dtype=dt, count=n, offset=int(vox_offset)).reshape(shape, order="F")
    if slope not in (0.0,) and not (slope == 1.0 and inter == 0.0):
        a = a.astype(np.float64) * slope + inter
    return a


# ---------------------------------------------------------------------------------------
def gen_ops():
    FK, FK2c, DTI, _ = import_reference()
    rng = np.random.default_rng(20261019)
    cases = {}
    shapes = [((7, 6, 5), (1.0, 1.5, 0.75), "holes"), ((9, 8, 11), (2.0, 2.0, 2.0), "holes"),
              ((12, 10, 13), (0.8, 1.1, 1.3), "full"), ((5, 17, 6), (1.0, 1.0, 1.0), "sparse")]
    for ci, (shp, h, kind) in enumerate(shapes):
        wm = rng.uniform(0, 1, shp)
        gm = rng.uniform(0, 1, shp) * (1 - wm)
        if kind == "holes":
            hole = rng.uniform(0, 1, shp) < 0.3
            wm[hole] = 0
            gm[hole] *= 0.05
        elif kind == "sparse":
            hole = rng.uniform(0, 1, shp) < 0.7
            wm[hole] = 0
            gm[hole] = 0
        elif kind == "full":
            wm = 0.3 + 0.5 * wm
            gm = 1.0 - wm           # tissue everywhere: periodic wrap fully active
        dx, dy, dz = h
        th, Dw, ratio, rho = 0.1, 0.7 + 0.2 * ci, 10.0 + 3 * ci, 0.11 + 0.02 * ci
        dt = 0.04 * min(h) ** 2 / Dw
        tag = f"c{ci}"
        cases[tag + "_wm"], cases[tag + "_gm"] = wm, gm
        cases[tag + "_h"] = np.array(h)
        cases[tag + "_fk_sc"] = np.array([th, Dw, ratio, rho, dt])
        # --- FK
        s = FK({})
        D = s.get_D(wm, gm, th, Dw, ratio)
        for k, v in D.items():
            cases[f"{tag}_fk_{k}"] = v
        u = rng.uniform(0, 1, shp) * (rng.uniform(0, 1, shp) < 0.6)
        cases[tag + "_fk_u0"] = u.copy()
        for st in range(3):
            u = s.FK_update(u, D, rho, dt, dx, dy, dz)
            cases[f"{tag}_fk_u{st + 1}"] = u.copy()
        # --- FK_2c
        s2 = FK2c({})
        th_necro, Ds, lnp, sig, ls = 0.9, 1.3 - 0.1 * ci, 0.35, 0.5, 0.05 + 0.01 * ci
        dt2 = 0.04 * min(h) ** 2 / max(Dw, Ds)
        cases[tag + "_2c_sc"] = np.array([th, th_necro, Dw, ratio, Ds, rho, lnp, sig, ls, dt2])
        P = rng.uniform(0, 1.0, shp) * (rng.uniform(0, 1, shp) < 0.7)
        N = rng.uniform(0, 0.6, shp) * (rng.uniform(0, 1, shp) < 0.5)
        S = rng.uniform(0.2, 1.0, shp) * (wm + gm >= th)
        cases[tag + "_2c_P0"], cases[tag + "_2c_N0"], cases[tag + "_2c_S0"] = P.copy(), N.copy(), S.copy()
        Dn = s2.get_D_with_necro(wm, gm, th, Dw, ratio, P, N, th_necro)
        Dss = s2.get_D(wm, gm, th, Ds, 1)
        for k, v in Dn.items():
            cases[f"{tag}_2c_{k}"] = v
        for k, v in Dss.items():
            cases[f"{tag}_2cS_{k}"] = v
        for st in range(3):
            upd = s2.FK_2c_update(P, N, S, Dn, rho, lnp, sig, Dss, ls, dt2, dx, dy, dz)
            Dn = s2.get_D_with_necro(wm, gm, th, Dw, ratio, P, upd["N"], th_necro)
            cases[f"{tag}_2c_P{st + 1}"] = P.copy()
            cases[f"{tag}_2c_N{st + 1}"] = N.copy()
            cases[f"{tag}_2c_S{st + 1}"] = S.copy()
        # --- FK_DTI
        s3 = DTI({})
        rgb = rng.uniform(0, 3, shp + (3,)) * (rng.uniform(0, 1, shp) < 0.8)[..., None]
        cases[tag + "_dti_rgb"] = rgb.copy()
        Dd = s3.get_D_from_DTI(rgb, Dw)
        dt3 = 0.04 * min(h) ** 2 / (Dw * 3)
        cases[tag + "_dti_sc"] = np.array([Dw, rho, dt3])
        u = rng.uniform(0, 1, shp) * (rng.uniform(0, 1, shp) < 0.6)
        cases[tag + "_dti_u0"] = u.copy()
        for st in range(3):
            u = s3.FK_update(u, Dd, rho, dt3, dx, dy, dz)
            cases[f"{tag}_dti_u{st + 1}"] = u.copy()
    cases["ncases"] = np.array(len(shapes))
    np.savez_compressed(os.path.join(GOLD, "ops.npz"), **cases)
    print("ops.npz", os.path.getsize(os.path.join(GOLD, "ops.npz")))


# ---------------------------------------------------------------------------------------
def series_digest(series):
    """(n,3) digest of a list of snapshots: sum, sum of squares, and a position-weighted sum
    (keeps the fixture small; the first and last snapshots are stored in full)."""
    series = np.asarray(series, dtype=np.float64)
    if series.size == 0:
        return np.zeros((0, 3))
    w = np.cos(0.37 * np.arange(series[0].size)).reshape(series[0].shape)
    return np.array([[s.sum(), (s * s).sum(), (s * w).sum()] for s in series])


def _pack_series(key, series, out):
    out[key + "/digest"] = series_digest(series)
    if len(series):
        out[key + "/first"] = np.asarray(series[0])
        out[key + "/last"] = np.asarray(series[-1])


def _pack_result(prefix, res, out):
    for k, v in res.items():
        if k == "time_series" and v is not None:
            if isinstance(v, dict):
                for kk, vv in v.items():
                    _pack_series(f"{prefix}/{k}/{kk}", vv, out)
            else:
                _pack_series(f"{prefix}/{k}", v, out)
        elif isinstance(v, dict):
            for kk, vv in v.items():
                out[f"{prefix}/{k}/{kk}"] = np.asarray(vv)
        elif v is None:
            out[f"{prefix}/{k}"] = np.array("__none__")
        else:
            out[f"{prefix}/{k}"] = np.asarray(v)


def _scalars(res):
    return {k: v for k, v in res.items() if np.ndim(v) == 0 and not isinstance(v, dict)}


def gen_solve():
    from tumorgrowthtoolkit_b200 import synthetic
    FK, FK2c, DTI, tools = import_reference()
    wm, gm = synthetic.brain_phantom((32, 40, 28))
    out = {"wm": wm, "gm": gm}
    base = dict(gm=gm, wm=wm, NxT1_pct=0.35, NyT1_pct=0.6, NzT1_pct=0.5)

    fk_cases = {
        "fk_a": dict(Dw=1.0, rho=0.10, RatioDw_Dg=10, th_matter=0.1, init_scale=1.0,
                     resolution_factor=0.5, time_series_solution_Nt=5, stopping_time=40),
        "fk_b": dict(Dw=0.6, rho=0.25, resolution_factor=1.0, stopping_time=12, dx_mm=1.0, dy_mm=1.2,
                     dz_mm=0.9),
        "fk_vol": dict(Dw=1.5, rho=0.3, resolution_factor=0.65, stopping_time=60, stopping_volume=900.0,
                       time_series_solution_Nt=7),
        "fk_outside": dict(Dw=1.0, rho=0.1, resolution_factor=0.5, NxT1_pct=0.02, NyT1_pct=0.02,
                           NzT1_pct=0.02),
        "fk_ts_gt_steps": dict(Dw=0.05, rho=0.02, resolution_factor=0.5, stopping_time=2,
                               time_series_solution_Nt=200),
    }
    for name, extra in fk_cases.items():
        p = dict(base)
        p.update(extra)
        t = time.time()
        res = FK(p).solve()
        print(name, f"{time.time() - t:.1f}s", _scalars(res))
        _pack_result(name, res, out)

    c2_cases = {
        "2c_a": dict(Dw=0.9, rho=0.14, lambda_np=0.35, sigma_np=0.5, D_s=1.3, lambda_s=0.05, RatioDw_Dg=100,
                     Nt_multiplier=8, resolution_factor=0.5, time_series_solution_Nt=4, stopping_time=60),
        "2c_b": dict(Dw=2.0, rho=0.5, lambda_np=0.6, sigma_np=0.3, D_s=0.8, lambda_s=0.3, resolution_factor=1.0,
                     Nt_multiplier=10, stopping_time=15, th_necro=0.8),
        "2c_vol": dict(Dw=1.2, rho=0.4, lambda_np=0.2, sigma_np=0.4, D_s=2.0, lambda_s=0.1,
                       resolution_factor=0.65, stopping_volume=1500.0, stopping_time=80),
    }
    for name, extra in c2_cases.items():
        p = dict(base)
        p.update(extra)
        t = time.time()
        res = FK2c(p).solve()
        print(name, f"{time.time() - t:.1f}s", _scalars(res))
        _pack_result(name, res, out)

    T = synthetic.dti_tensor_field(wm, gm, seed=7)
    dti_base = dict(diffusionTensors=T, NxT1_pct=0.35, NyT1_pct=0.6, NzT1_pct=0.5)
    dti_cases = {
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
    for name, extra in dti_cases.items():
        p = dict(dti_base)
        p.update(extra)
        t = time.time()
        res = DTI(p).solve()
        print(name, f"{time.time() - t:.1f}s", _scalars(res))
        _pack_result(name, res, out)
    # the intermediate the DTI solver feeds its loop with (tools.py:166-187), for host-logic tests
    out["dti_rgb_exp1"] = tools.makeXYZ_rgb_from_tensor(T, exponent=1, linear=0)
    out["dti_rgb_exp15"] = tools.makeXYZ_rgb_from_tensor(T, exponent=1.5, linear=0.2)
    np.savez_compressed(os.path.join(GOLD, "solve.npz"), **out)
    print("solve.npz", os.path.getsize(os.path.join(GOLD, "solve.npz")))


# ---------------------------------------------------------------------------------------
def gen_atlas():
    from scipy.ndimage import zoom
    FK, FK2c, DTI, tools = import_reference()
    seg = read_nifti(os.path.join(REF, "dataset", "sub-mni152_tissue-with-antsN4_space-sri.nii.gz"))
    seg8 = np.ascontiguousarray(np.rint(seg).astype(np.uint8))
    assert seg8.shape == (240, 240, 155) and set(np.unique(seg8)) <= {0, 1, 2, 3}
    np.savez_compressed(os.path.join(GOLD, "sri_tissue_labels.npz"), seg=seg8)
    print("sri_tissue_labels.npz", os.path.getsize(os.path.join(GOLD, "sri_tissue_labels.npz")))
    wm, gm = (seg8 == 3).astype(np.float64), (seg8 == 2).astype(np.float64)

    out = {}
    # ---- FK, FK_exampleAtlas.ipynb parameters, the reference's own solve()
    p = dict(Dw=1.0, rho=0.10, RatioDw_Dg=10, gm=gm, wm=wm, NxT1_pct=0.3, NyT1_pct=0.7, NzT1_pct=0.5,
             init_scale=1.0, resolution_factor=0.5, th_matter=0.1, verbose=True, time_series_solution_Nt=None)
    t = time.time()
    res = FK(p).solve()
    print("atlas FK solve", f"{time.time() - t:.1f}s", res["final_volume"], res["final_time"])
    out["fk/final_volume"] = np.array(res["final_volume"])
    out["fk/final_time"] = np.array(res["final_time"])
    # low-res cropped loop result, by driving the reference's own methods like solve() does
    s = FK(p)
    rf = 0.5
    gml, wml = zoom(gm, rf, order=1), zoom(wm, rf, order=1)
    Nx, Ny, Nz = gml.shape
    dx = 1.0 / rf
    Nt = np.max([100 * 1.0 / dx ** 2 * 8 + 100, 100 * 0.1 * 1.1])
    dt, steps = 100 / Nt, int(np.ceil(Nt))
    xv, yv, zv = np.meshgrid(np.arange(Nx), np.arange(Ny), np.arange(Nz), indexing="ij")
    A = np.array(s.gauss_sol3d(xv - int(0.3 * Nx), yv - int(0.7 * Ny), zv - int(0.5 * Nz), dx, dx, dx, 1.0))
    cg, cw, A, (mn, mx) = s.crop_tissues_and_tumor(gml, wml, A, margin=2, threshold=0.5)
    D = s.get_D(cw, cg, 0.1, 1.0, 10)
    A = np.array(A)
    for _ in range(steps):
        A = s.FK_update(A, D, 0.10, dt, dx, dx, dx)
    vol = dx ** 3 * np.sum(A)
    assert abs(vol - res["final_volume"]) < 1e-9 * vol, (vol, res["final_volume"])
    out["fk/steps"] = np.array(steps)
    out["fk/crop"] = np.array([mn, mx])
    out["fk/final_lowres_crop"] = A
    print("atlas FK loop ok", steps, A.shape)

    # ---- FK_2c, FK_2c_example.ipynb parameters
    p2 = dict(Dw=0.9, rho=0.14, lambda_np=0.35, sigma_np=0.5, D_s=1.3, lambda_s=0.05, gm=gm, wm=wm,
              NxT1_pct=0.3, NyT1_pct=0.7, NzT1_pct=0.5, init_scale=1.0, resolution_factor=0.5, th_matter=0.1,
              th_necro=0.9, RatioDw_Dg=100, Nt_multiplier=8, verbose=True, time_series_solution_Nt=None)
    s2 = FK2c(p2)
    Nt = 100 * max(0.9, 1.3) / dx ** 2 * 8 + 300
    dt, steps = 100 / Nt, int(np.ceil(Nt))
    init = s2.get_initial_configuration(int(0.3 * Nx), int(0.7 * Ny), int(0.5 * Nz), Nx, Ny, Nz, dx, dx, dx, 1.0,
                                        gml, wml, 0.1)
    cg, cw, st, (mn, mx) = s2.crop_tissues_and_tumor(gml, wml, init["P"], init["N"], init["S"], margin=2,
                                                     threshold=0.5)
    st = {k: np.array(v) for k, v in st.items()}
    Dn = s2.get_D_with_necro(cw, cg, 0.1, 0.9, 100, st["P"], st["N"], 0.9)
    Dss = s2.get_D(cw, cg, 0.1, 1.3, 1)
    t = time.time()
    for _ in range(steps):
        upd = s2.FK_2c_update(st["P"], st["N"], st["S"], Dn, 0.14, 0.35, 0.5, Dss, 0.05, dt, dx, dx, dx)
        Dn = s2.get_D_with_necro(cw, cg, 0.1, 0.9, 100, st["P"], upd["N"], 0.9)
    vol = dx ** 3 * np.sum(st["P"]) + np.sum(st["N"])
    print("atlas FK_2c loop", f"{time.time() - t:.1f}s", steps, vol, st["P"].sum(), st["N"].sum(), st["N"].max())
    out["2c/steps"] = np.array(steps)
    out["2c/crop"] = np.array([mn, mx])
    out["2c/final_volume"] = np.array(vol)
    for k in "PNS":
        out[f"2c/final_lowres_crop_{k}"] = st[k]
    np.savez_compressed(os.path.join(GOLD, "atlas_res05.npz"), **out)
    print("atlas_res05.npz", os.path.getsize(os.path.join(GOLD, "atlas_res05.npz")))


def gen_dti_atlas():
    """SURVEY.md 8(d) input 3 / VERDICT r1 item 10: the synthetic FSL_HCP1065-shaped lower-6 tensor volume on the SRI
    label atlas -> tools.get_tensor_from_lower6 -> FK_DTI_Solver.solve() of the unmodified reference with the
    FK_DTI_example.ipynb parameters at resolution_factor 0.5, once with the notebook's stop criterion
    (stopping_volume 15000, stopping_time 1000) and once as the fixed-step variant (stopping_time 100).  The
    full-resolution final states (71 MB each) are stored sub-sampled by 2 per axis plus three digests of the full array."""
    from tumorgrowthtoolkit_b200 import synthetic
    FK, FK2c, DTI, tools = import_reference()
    seg = np.load(os.path.join(GOLD, "sri_tissue_labels.npz"))["seg"]
    lower6 = synthetic.dti_lower6_field(seg)
    tensors = tools.get_tensor_from_lower6(lower6)
    pct = synthetic.seed_inside(seg == 3, (0.49, 0.30, 0.7))
    out = {"seed_pct": np.array(pct)}
    base = dict(Dw=1.0, rho=0.2, diffusionEllipsoidScaling=1, diffusionTensors=tensors, NxT1_pct=pct[0], NyT1_pct=pct[1],
                NzT1_pct=pct[2], resolution_factor=0.5, verbose=True)
    cases = {"stop": dict(init_scale=0.1, stopping_volume=15000, stopping_time=1000),
             "fixed": dict(init_scale=1.0, stopping_time=100)}
    for name, extra in cases.items():
        t = time.time()
        res = DTI(dict(base, **extra)).solve()
        print("atlas DTI", name, f"{time.time() - t:.1f}s", _scalars(res))
        assert res["success"], res
        fs = res["final_state"]
        w = np.cos(0.37 * np.arange(fs.size)).reshape(fs.shape)
        out[name + "/final_state_sub2"] = np.ascontiguousarray(fs[::2, ::2, ::2])
        out[name + "/final_state_digest"] = np.array([fs.sum(), (fs * fs).sum(), (fs * w).sum()])
        out[name + "/final_time"] = np.array(res["final_time"])
        out[name + "/final_volume"] = np.array(res["final_volume"])
        out[name + "/stopping_criteria"] = np.array(res["stopping_criteria"])
        out[name + "/initial_state_digest"] = np.array([res["initial_state"].sum(), (res["initial_state"] ** 2).sum()])
    np.savez_compressed(os.path.join(GOLD, "atlas_dti_res05.npz"), **out)
    print("atlas_dti_res05.npz", os.path.getsize(os.path.join(GOLD, "atlas_dti_res05.npz")))


def gen_segmap():
    """create_segmentation_map of the unmodified reference script (synthetic_gens/generatePatient_FK_2c.py:67-83) on seeded
    fields, three threshold sets drawn with the script's own law: pins the label map the ensemble path produces."""
    import importlib.util
    import_reference()                                   # also mocks nibabel for the script's imports
    spec = importlib.util.spec_from_file_location("ref_generate_patient", os.path.join(REF, "synthetic_gens", "generatePatient_FK_2c.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(20261021)
    shape = (24, 20, 18)
    P = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.7)
    N = rng.uniform(0, 0.5, shape) * (rng.uniform(0, 1, shape) < 0.4)
    P[0, 0, :6] = [0.15, 0.4, 0.12, 0.2, 0.25, 0.6]      # values exactly on thresholds
    N[0, 1, :3] = [0.12, 0.05, 0.2]
    out = {"P": P, "N": N}
    ths = [(0.12, 0.4, 0.15), (0.05, 0.25, 0.12), (0.2, 0.6, 0.2)]
    out["thresholds"] = np.array(ths)
    for i, th in enumerate(ths):
        out[f"seg{i}"] = mod.create_segmentation_map(P, N, *th)
    np.savez_compressed(os.path.join(GOLD, "segmap.npz"), **out)
    print("segmap.npz", os.path.getsize(os.path.join(GOLD, "segmap.npz")))


def elongate_inputs():
    """Seeded (6,7,5,3,3) tensor field for the elongation (tools.py:109-164): SPD tensors with separated
    eigenvalues, isotropic and axis-aligned degenerate tensors, zeros (background)."""
    rng = np.random.default_rng(20261020)
    shape = (6, 7, 5)
    n = int(np.prod(shape))
    q, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    lam = np.sort(rng.uniform(0.2, 1.0, (n, 3)), axis=-1) * np.array([1.0, 1.6, 2.6]) * 1e-3
    t = np.einsum("nij,nj,nkj->nik", q, lam, q)
    t = 0.5 * (t + np.swapaxes(t, -1, -2))
    special = [np.diag([1e-3] * 3), np.diag([1e-3, 2e-3, 2e-3]), np.diag([2e-3, 2e-3, 1e-3]), np.diag([2e-3, 1e-3, 2e-3]),
               np.diag([3e-3, 1e-3, 2e-3]), np.zeros((3, 3)), np.diag([0.0, 0.0, 1.5e-3]), np.diag([7e-4, 7e-4, 0.0])]
    for i, m in enumerate(special):
        t[3 * i] = m
    return t.reshape(shape + (3, 3))


def gen_elongate():
    """tools.elongate_tensor_along_main_axis_torch of the unmodified reference (float32 torch.linalg.eigh)."""
    _, _, _, tools = import_reference()
    t = elongate_inputs()
    out = {"tensors": t}
    for scale in (2.0, 0.6, 1.25):
        out["scale_%g" % scale] = tools.elongate_tensor_along_main_axis_torch(t.copy(), scale)
    np.savez_compressed(os.path.join(GOLD, "elongate.npz"), **out)
    print("elongate.npz", os.path.getsize(os.path.join(GOLD, "elongate.npz")))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    groups = sys.argv[1:] or ["ops", "solve", "atlas", "elongate", "dti_atlas", "segmap"]
    for g in groups:
        {"ops": gen_ops, "solve": gen_solve, "atlas": gen_atlas, "elongate": gen_elongate, "dti_atlas": gen_dti_atlas,
         "segmap": gen_segmap}[g]()
