"""Host-side ensemble logic: the parameter law of run_gen_FK_2c.py and the rank sharding (CPU only)."""
import numpy as np


def test_parameter_law_matches_reference_script():
    """First draws of np.random.seed(42000) in the dict order of run_gen_FK_2c.py:8-23."""
    from tumorgrowthtoolkit_b200 import ensemble
    np.random.seed(42000)
    want = [np.random.uniform(0.05, 4), np.random.uniform(0.01, 0.8), np.random.uniform(0.05, 0.7)]
    p = ensemble.draw_ensemble(1)[0]
    assert [p["Dw"], p["rho"], p["lambda_np"]] == want
    assert p["resolution_factor"] == 0.65 and p["RatioDw_Dg"] == 10


def test_shard_partitions_and_balances():
    from tumorgrowthtoolkit_b200 import ensemble
    params = ensemble.draw_ensemble(64)
    parts = [ensemble.shard(params, r, 8) for r in range(8)]
    assert sorted(i for part in parts for i in part) == list(range(64))
    loads = [sum(ensemble.step_count(params[i])[0] for i in part) for part in parts]
    assert max(loads) <= 1.15 * min(loads)


def test_segmentation_map_matches_reference_script():
    """ensemble.create_segmentation_map against labels produced by the reference script's own function
    (tests/make_golden.py:gen_segmap imports synthetic_gens/generatePatient_FK_2c.py:67-83), values on the thresholds included.
    tests/test_gpu_ensemble.py holds the device label kernel to this host function on every patient."""
    import os
    from tumorgrowthtoolkit_b200 import ensemble
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "segmap.npz"))
    for i, th in enumerate(g["thresholds"]):
        got = ensemble.create_segmentation_map(g["P"], g["N"], *th)
        assert np.array_equal(got, g[f"seg{i}"])
        assert set(np.unique(got)) <= {0, 1, 3, 4}
