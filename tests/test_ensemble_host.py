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
