"""solve()-level host logic (resample, seed, step-count law, crop/restore, stop handling,
result dict) checked on the CPU against results of the reference's own solve()
(tests/golden/solve.npz).  The device time loop is replaced by the CPU oracle's loop here
(the oracle as the checker); the same comparisons run against the CUDA loop in
tests/test_gpu_solve.py."""
import numpy as np
import pytest

from tumorgrowthtoolkit_b200 import _loop
from tumorgrowthtoolkit_b200.FK import Solver as FK
from tumorgrowthtoolkit_b200.FK_2c import Solver as FK2c
from tumorgrowthtoolkit_b200.FK_DTI import FK_DTI_Solver as DTI
from tests.solve_cases import (FK_CASES, C2_CASES, DTI_CASES, base_params, dti_params, compare_result, atlas_dti_params,
                               compare_atlas_dti)


@pytest.fixture()
def cpu_loops(monkeypatch, oracle):
    def drop(kw):
        for k in ("precision", "device", "bit_exact", "devices", "output"):
            kw.pop(k, None)
        return kw
    monkeypatch.setattr(_loop, "fk_loop", lambda *a, **kw: oracle.fk_loop(*a, **drop(kw)))
    monkeypatch.setattr(_loop, "fk2c_loop", lambda *a, **kw: oracle.fk2c_loop(*a, **drop(kw)))
    monkeypatch.setattr(_loop, "dti_loop", lambda *a, **kw: oracle.dti_loop(*a, **drop(kw)))

    def cpu_resample(a, out_shape, lo=None, virt_shape=None, device=0):
        # the device resample replaced by the oracle's (restore_tumor + zoom, FK.py:75-90,233-238)
        a = np.asarray(a, dtype=np.float64)
        if virt_shape is not None:
            full = np.zeros(tuple(int(v) for v in virt_shape))
            full[tuple(slice(int(l), int(l) + n) for l, n in zip(lo, a.shape))] = a
            a = full
        return oracle.zoom_linear(a, out_shape)
    from tumorgrowthtoolkit_b200 import _host
    monkeypatch.setattr(_host, "_device_resample", cpu_resample)
    from tumorgrowthtoolkit_b200.FK_DTI import tools
    monkeypatch.setattr(tools, "makeXYZ_rgb_from_tensor_cuda", lambda t, exponent=1, linear=0, device=0: tools.makeXYZ_rgb_from_tensor(t, exponent, linear))


@pytest.mark.parametrize("name", sorted(FK_CASES))
def test_fk_solve_host_logic(name, gold_solve, cpu_loops):
    res = FK(base_params(gold_solve, FK_CASES[name])).solve()
    compare_result(name, res, gold_solve, tol=1e-12)


@pytest.mark.parametrize("name", sorted(C2_CASES))
def test_fk2c_solve_host_logic(name, gold_solve, cpu_loops):
    res = FK2c(base_params(gold_solve, C2_CASES[name])).solve()
    compare_result(name, res, gold_solve, tol=1e-10)


@pytest.mark.parametrize("name", sorted(DTI_CASES))
def test_dti_solve_host_logic(name, gold_solve, cpu_loops):
    tol = 2e-4 if name == "dti_elong" else 1e-12   # float32 eigh inside the reference's elongation
    p = dti_params(gold_solve, DTI_CASES[name])
    p["elongation_backend"] = "torch"              # no GPU here: the literal restatement of the reference's CPU eigh
    res = DTI(p).solve()
    compare_result(name, res, gold_solve, tol=tol)


def test_atlas_dti_solve_host_logic(atlas_labels, gold_atlas_dti, cpu_loops):
    """The config-3 input through the solver mirror with the oracle's loop: resample of the 4-D rgb volume, seed
    re-picked inside WM, step-count law on the full-resolution maximum, the notebook's stopping_volume."""
    res = DTI(atlas_dti_params(atlas_labels, gold_atlas_dti, "stop")).solve()
    compare_atlas_dti("stop", res, gold_atlas_dti, 1e-10)


def test_rgb_from_tensor_matches_reference(gold_solve):
    from tumorgrowthtoolkit_b200 import synthetic
    import tumorgrowthtoolkit_b200.FK_DTI.tools as tools
    T = synthetic.dti_tensor_field(gold_solve["wm"], gold_solve["gm"], seed=7)
    assert np.array_equal(tools.makeXYZ_rgb_from_tensor(T, 1, 0), gold_solve["dti_rgb_exp1"])
    assert np.max(np.abs(tools.makeXYZ_rgb_from_tensor(T, 1.5, 0.2) - gold_solve["dti_rgb_exp15"])) < 1e-14


def test_missing_required_key_raises(gold_solve):
    p = base_params(gold_solve, FK_CASES["fk_a"])
    del p["rho"]
    with pytest.raises(KeyError):
        FK(p).solve()


def test_bad_seed_fraction_asserts(gold_solve):
    p = base_params(gold_solve, FK_CASES["fk_a"])
    p["NxT1_pct"] = 1.5
    with pytest.raises(AssertionError):
        FK(p).solve()


def test_params_not_mutated(gold_solve, cpu_loops):
    p = base_params(gold_solve, FK_CASES["fk_a"])
    wm0, gm0 = p["wm"].copy(), p["gm"].copy()
    s = FK(p)
    assert s.params is p
    s.solve()
    assert np.array_equal(p["wm"], wm0) and np.array_equal(p["gm"], gm0)


def test_install_as_alias():
    import tumorgrowthtoolkit_b200
    tumorgrowthtoolkit_b200.install_as("TumorGrowthToolkit_alias_for_test")
    from TumorGrowthToolkit_alias_for_test.FK_2c import Solver
    assert Solver is FK2c
