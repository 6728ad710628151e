"""Multi-GPU paths through torchrun (skipped on a one-GPU box): slab decomposition over NCCL must
reproduce the undecomposed loop bitwise; the ensemble shards must cover every patient."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    from tumorgrowthtoolkit_b200 import _native
    import __graft_entry__ as ge
    ge.build()
    return _native.device_count()


def _torchrun(n, script, *args, port=29561):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("driver", ["fused", "p2p", "native", "torch"])
@pytest.mark.parametrize("model,world", [("dti", 2), ("fk2c", 2), ("dti", 4), ("fk", 8)])
def test_slab_ranks_are_bitwise(model, world, driver):
    """74 planes over 4 ranks = 19,19,18,18: uneven slabs exercise the neighbour-specific ghost offsets."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    halo = "1" if driver == "fused" else "3"
    r = _torchrun(world, "run_slab.py", "--model", model, "--steps", "45", "--halo", halo, "--factor", "0.5", "--check",
                  "--driver", driver, "--reps", "1")
    assert r["n_gpus"] == world and r["max_abs_diff_vs_undecomposed"] == 0.0
    assert r["volume_rel_diff"] <= 1e-12


@pytest.mark.parametrize("model,world", [("dti", 2), ("fk2c", 2), ("dti", 4)])
def test_fused_slab_stop_criterion_across_ranks(model, world):
    """stopping_volume crossed mid-chunk: every rank replays from its checkpoint to the same step; state bitwise."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun(world, "run_slab.py", "--model", model, "--steps", "150", "--factor", "0.5", "--check", "--stop-fraction", "0.6",
                  "--reps", "1", port=29563)
    assert r["stop"] == "volume" and r["same_stop"] and 1 < r["steps_run"] < 150
    assert r["max_abs_diff_vs_undecomposed"] == 0.0


def test_solver_api_on_two_ranks():
    """FK_DTI_Solver({..., 'devices': 'slab'}).solve() with the notebook's stopping_volume: same result dict as one GPU."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "run_slab.py", "--solve", "--factor", "0.5", port=29564)
    assert r["success"] and r["same_keys"] and r["same_stop"], r
    assert r["final_state_max_abs_diff"] == 0.0
    assert abs(r["final_volume"] - r["one_gpu_final_volume"]) <= 1e-12 * r["one_gpu_final_volume"]


def test_ensemble_two_ranks():
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "run_ensemble.py", "--patients", "6", "--lowres", port=29562)
    assert r["n_gpus"] == 2 and r["patients"] == 6 and r["glups"] > 0
