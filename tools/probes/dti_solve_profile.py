"""cProfile of one FK_DTI_Solver.solve() at the notebook's parameters (where the host time of a whole solve goes)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tumorgrowthtoolkit_b200 import synthetic
from tumorgrowthtoolkit_b200.FK_DTI import FK_DTI_Solver as DTI
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
wm, gm = synthetic.tissue_from_labels(np.load(os.path.join(root, "tests", "golden", "sri_tissue_labels.npz"))["seg"])
T = synthetic.dti_tensor_field(wm, gm)
p = dict(Dw=1.0, rho=0.2, init_scale=0.1, diffusionEllipsoidScaling=1, stopping_volume=15000, stopping_time=1000,
         diffusionTensors=T, resolution_factor=1.0, NxT1_pct=0.49, NyT1_pct=0.30, NzT1_pct=0.7)
DTI(dict(p)).solve()
t0 = time.perf_counter(); DTI(dict(p)).solve(); print("solve", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); DTI(dict(p)).solve(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
