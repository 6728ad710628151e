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
