"""Wall time of whole Solver.solve() calls at the reference notebooks' parameters (incl. the resample
pre/post-processing that dominates the reference, SURVEY.md finding 5)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tumorgrowthtoolkit_b200 import synthetic
from tumorgrowthtoolkit_b200.FK import Solver as FK
from tumorgrowthtoolkit_b200.FK_2c import Solver as FK2c
from tumorgrowthtoolkit_b200.FK_DTI import FK_DTI_Solver as DTI

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wm, gm = synthetic.tissue_from_labels(np.load(os.path.join(root, "tests", "golden", "sri_tissue_labels.npz"))["seg"])
seed = dict(NxT1_pct=0.3, NyT1_pct=0.7, NzT1_pct=0.5)
fk = dict(Dw=1.0, rho=0.10, RatioDw_Dg=10, gm=gm, wm=wm, init_scale=1.0, th_matter=0.1, **seed)
c2 = dict(Dw=0.9, rho=0.14, lambda_np=0.35, sigma_np=0.5, D_s=1.3, lambda_s=0.05, gm=gm, wm=wm, init_scale=1.0,
          th_matter=0.1, th_necro=0.9, RatioDw_Dg=100, Nt_multiplier=8, **seed)
cases = [("FK   res 0.5, 64 snapshots (FK_exampleAtlas.ipynb)", FK, dict(fk, resolution_factor=0.5, time_series_solution_Nt=64)),
         ("FK   res 1.0, no snapshots", FK, dict(fk, resolution_factor=1.0)),
         ("FK_2c res 0.5, 8 snapshots (FK_2c_example.ipynb)", FK2c, dict(c2, resolution_factor=0.5, time_series_solution_Nt=8)),
         ("FK_2c res 1.0, no snapshots", FK2c, dict(c2, resolution_factor=1.0))]
T = synthetic.dti_tensor_field(wm, gm)
cases.append(("FK_DTI res 1.0, stopping_volume 15000 (FK_DTI_example.ipynb)", DTI,
              dict(Dw=1.0, rho=0.2, init_scale=0.1, diffusionEllipsoidScaling=1, stopping_volume=15000, stopping_time=1000,
                   diffusionTensors=T, resolution_factor=1.0, NxT1_pct=0.49, NyT1_pct=0.30, NzT1_pct=0.7)))
# where the wall time of a solve() goes: wrap the host-side stages
import functools
from tumorgrowthtoolkit_b200 import _host, _loop
from tumorgrowthtoolkit_b200.FK_DTI import tools as dti_tools
PHASES = {}


def _timed(mod, fn, label):
    orig = getattr(mod, fn)

    @functools.wraps(orig)
    def wrapper(*a, **k):
        t = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            PHASES[label] = PHASES.get(label, 0.0) + time.perf_counter() - t
    setattr(mod, fn, wrapper)


for mod, fn, label in ((_host, "resample_linear", "resample in"), (_host, "seed_gaussian", "seed"), (_host, "bounding_box", "bbox"),
                       (_loop, "fk_loop", "loop+outputs"), (_loop, "fk2c_loop", "loop+outputs"), (_loop, "dti_loop", "loop+outputs"),
                       (dti_tools, "makeXYZ_rgb_from_tensor", "rgb from tensor"), (_host.LowResGrid, "upsample", "upsample initial")):
    _timed(mod, fn, label)

for name, cls, p in cases:
    for rep in range(2):
        PHASES.clear()
        s = cls(dict(p))
        t0 = time.perf_counter()
        res = s.solve()
        wall = time.perf_counter() - t0
    st = getattr(s, "last_loop_stats", {})
    print(f"{name}: solve() {wall:.2f} s, loop {st.get('loop_ms', float('nan')) / 1e3:.3f} s for {st.get('steps_run')} steps on "
          f"{st.get('voxels')} voxels, success={res.get('success')}, final_volume={res.get('final_volume')}; phases "
          + ", ".join(f"{k} {v:.3f}" for k, v in PHASES.items()), flush=True)
