"""Parity on the configurations bench.py measures (VERDICT r1 "what's weak" 1): the cropped res-1.0 atlas box
145x178x139 = 3.59 M voxels at the FULL step counts of the notebooks' parameters -- FK 900, FK_2c 1340, FK_DTI ~1630
(synthetic FSL_HCP1065-shaped tensor volume, SURVEY.md 8(d) input 3) -- CUDA through the C ABI against the CPU oracle
(OpenMP C restatement of the reference loops, pinned on reference-generated vectors) on the same inputs.

Gates (BASELINE.json): fp64 <= 1e-6 on every field (asserted: 1e-9), fp32 <= 1e-4, identical thresholded
segmentation (Dice = 1.0).  FK_2c in pure fp32 cannot meet the 1e-4 gate on P and N -- neither can the reference's
own arithmetic run in float32 (4.7e-3 / 1.5e-3 at res 0.5) -- because the `P+N <= th_necro` face test
(FK_2c.py:13-15) is discontinuous; what is asserted for it is stated in the test's name."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_cache = {}


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


def _problem_and_reference(model):
    """(problem, oracle final fields, oracle final volume) -- the oracle runs every time step of the loop once per module."""
    if model not in _cache:
        import bench
        pb = bench.build_problem(model)
        st = bench.OracleStepper(pb)
        for _ in range(pb["nsteps"]):
            st.step()
        _cache[model] = (pb, {k: v.copy() for k, v in st.fields().items()}, st.volume)
    return _cache[model]


def _gpu(pb, precision):
    import bench
    sim, out_fields = bench.make_sim(pb, precision, 0)
    sim.run(pb["nsteps"])
    st = sim.sync()
    assert st.steps_run == pb["nsteps"]
    names = "PNS" if pb["model"] == "fk2c" else "u"
    got = {k: sim.download(f) for k, f in zip(names, out_fields)}
    sim.close()
    return got, st.last_volume


@pytest.mark.parametrize("model,nsteps", [("fk", 900), ("fk2c", 1340), ("dti", None)])
def test_fp64_full_loop_on_the_benchmarked_box(model, nsteps):
    import bench
    pb, ref, vol = _problem_and_reference(model)
    assert pb["shape"] == (145, 178, 139)
    if nsteps is not None:
        assert pb["nsteps"] == nsteps                      # the step counts the reference prints at these parameters
    got, gvol = _gpu(pb, "fp64")
    for k in ref:
        err = float(np.max(np.abs(got[k] - ref[k])))
        assert err <= 1e-9, (model, k, err)
    assert abs(gvol - vol) <= 1e-9 * abs(vol)
    assert bench.dice_of(got, ref, model) == 1.0


@pytest.mark.parametrize("model", ["fk", "dti"])
def test_fp32_full_loop_on_the_benchmarked_box(model):
    import bench
    pb, ref, vol = _problem_and_reference(model)
    got, gvol = _gpu(pb, "fp32")
    err = float(np.max(np.abs(got["u"] - ref["u"])))
    assert err <= 1e-4, (model, err)
    assert abs(gvol - vol) <= 1e-4 * abs(vol)
    assert bench.dice_of(got, ref, model) == 1.0


def test_fk2c_pure_fp32_full_loop_misses_the_1e_4_gate_only_at_isolated_closure_flip_voxels():
    """Pure fp32 FK_2c: S within 1e-4 everywhere, mean error of P and N <= 1e-6, at most 0.01 % of the voxels of P / N
    above 1e-4 (isolated `P+N <= th_necro` flips, one step early or late), Dice >= 0.999.  NOT the 1e-4 max-abs gate:
    FK_2c meets its gate in fp64 only ("FK_2c and reduced precision")."""
    import bench
    pb, ref, vol = _problem_and_reference("fk2c")
    got, gvol = _gpu(pb, "fp32")
    errs = {k: np.abs(got[k] - ref[k]) for k in "PNS"}
    report = {k: (float(e.max()), float(e.mean()), int((e > 1e-4).sum())) for k, e in errs.items()}
    print("fp32 FK_2c, 1340 steps, max / mean / voxels above 1e-4:", report)
    assert errs["S"].max() <= 1e-4, report
    for k in "PN":
        assert errs[k].mean() <= 1e-6, report
        assert (errs[k] > 1e-4).sum() <= 1e-4 * errs[k].size, report
    assert bench.dice_of(got, ref, "fk2c") >= 0.999
