"""Ensemble mode: results independent of concurrency / sharding and equal to single solves."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


def test_ensemble_equals_single_solves_and_is_order_independent(gold_solve):
    from tumorgrowthtoolkit_b200 import ensemble
    from tumorgrowthtoolkit_b200.FK_2c import Solver
    wm, gm = gold_solve["wm"], gold_solve["gm"]
    params = ensemble.draw_ensemble(5)
    for p in params:                       # keep the test short and the seed inside the phantom
        p["NxT1_pct"], p["NyT1_pct"], p["NzT1_pct"] = 0.35, 0.6, 0.5
    kw = dict(resolution_factor=0.65, stopping_time=20)
    r1, st1 = ensemble.EnsembleRunner(wm, gm, concurrency=1, **kw).run(params)
    r4, st4 = ensemble.EnsembleRunner(wm, gm, concurrency=4, **kw).run(params[::-1])
    assert st1["patients"] == 5 and st1["voxel_updates"] == st4["voxel_updates"]
    for a, b in zip(r1, r4[::-1]):
        for k in "PNS":
            assert np.array_equal(a["final_state"][k], b["final_state"][k])
        assert np.array_equal(a["segmentation"], b["segmentation"])
    # against the drop-in solver with the same parameters (generatePatient_FK_2c.py:107-135)
    p = dict(params[0], gm=gm, wm=wm, th_matter=0.1, th_necro=0.9, Nt_multiplier=10, stopping_time=20,
             init_scale=1.0, dx_mm=1.0, dy_mm=1.0, dz_mm=1.0)
    res = Solver(p).solve()
    for k in "PNS":
        assert np.max(np.abs(res["final_state"][k] - r1[0]["final_state"][k])) <= 1e-12
    assert abs(res["final_volume"] - r1[0]["final_volume"]) <= 1e-9 * max(1.0, res["final_volume"])


def test_device_outputs_equal_host_postprocessing(gold_solve):
    """The asynchronous output path (restore + zoom + label map on the device, page-locked buffers) returns exactly
    what the synchronous download and the host's create_segmentation_map (generatePatient_FK_2c.py:67-83) return."""
    from tumorgrowthtoolkit_b200 import ensemble
    from tumorgrowthtoolkit_b200.FK_2c import Solver
    wm, gm = gold_solve["wm"], gold_solve["gm"]
    params = ensemble.draw_ensemble(3)
    for p in params:
        p["NxT1_pct"], p["NyT1_pct"], p["NzT1_pct"] = 0.35, 0.6, 0.5
    for precision in ("fp64", "fp32"):
        runner = ensemble.EnsembleRunner(wm, gm, concurrency=2, resolution_factor=0.65, stopping_time=20, precision=precision)
        res, _ = runner.run(params)
        for r in res:
            fs = r["final_state"]
            assert r["segmentation"].dtype == np.uint8 and r["segmentation"].shape == wm.shape
            want = ensemble.create_segmentation_map(fs["P"], fs["N"], r["params"].get("th_necro_n", 0.12),
                                                    r["params"].get("th_enhancing_p", 0.4), r["params"].get("th_edema_p", 0.15))
            assert np.array_equal(r["segmentation"], want)
            assert set(np.unique(r["segmentation"])) <= {0, 1, 3, 4}
        p = dict(params[1], gm=gm, wm=wm, th_matter=0.1, th_necro=0.9, Nt_multiplier=10, stopping_time=20, init_scale=1.0,
                 precision=precision)
        ref = Solver(p).solve()
        for k in "PNS":
            assert np.array_equal(ref["final_state"][k], res[1]["final_state"][k]), (precision, k)
    # results of a kept run are the caller's own arrays, not views of the runner's ring
    a = res[0]["final_state"]["P"]
    runner.run(params[:1])
    assert np.array_equal(a, res[0]["final_state"]["P"])


def test_pinned_arrays_round_trip():
    from tumorgrowthtoolkit_b200 import _native as nat
    a = nat.pinned_empty((7, 5, 3))
    a[...] = np.arange(105.0).reshape(7, 5, 3)
    b = nat.pinned_empty((4, 4), np.uint8)
    b[...] = 7
    assert a.flags.c_contiguous and a.sum() == 105 * 104 / 2 and b.sum() == 112
    v = a[2:4]
    del a
    assert v[0, 0, 0] == 30.0          # the view keeps the page-locked block alive
