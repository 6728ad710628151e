"""Small all-model run for compute-sanitizer (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tumorgrowthtoolkit_b200 import _loop
rng = np.random.default_rng(0)
for shape in ((19, 21, 37), (5, 40, 70)):
    wm = rng.uniform(0, 1, shape) * (rng.uniform(0, 1, shape) < 0.8); gm = rng.uniform(0, 1, shape) * (1 - wm)
    u0 = rng.uniform(0, 1, shape); N0 = 0.3 * u0[::-1]; S0 = rng.uniform(0.2, 1, shape)
    rgb = rng.uniform(0, 3, shape + (3,))
    for prec in ("fp64", "fp32"):
        _loop.fk_loop(u0, wm, gm, 0.1, 1.0, 10.0, 0.2, 0.02, 1.0, 1.0, 1.0, 7, precision=prec)
        _loop.fk_loop(u0, wm, gm, 0.1, 1.0, 10.0, 0.2, 0.02, 1.0, 1.0, 1.0, 7, precision=prec, stopping_volume=1e9)
        _loop.dti_loop(u0, rgb, 1.0, 0.2, 0.005, 1.0, 1.0, 1.0, 7, precision=prec)
        _loop.fk2c_loop(u0, N0, S0, wm, gm, 0.1, 0.9, 1.0, 10.0, 1.3, 0.2, 0.35, 0.5, 0.05, 0.02, 1.0, 1.0, 1.0, 7, precision=prec)
        _loop.fk2c_loop(u0, N0, S0, wm, gm, 0.1, 0.9, 1.0, 10.0, 1.3, 0.2, 0.35, 0.5, 0.05, 0.02, 1.0, 1.0, 1.0, 7, precision=prec,
                        record_steps=[0, 3, 6], stopping_volume=1e9)
print("sanitize case done")
