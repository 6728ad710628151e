"""Per-step device time of the step kernels vs box size (fixed overhead + throughput)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tumorgrowthtoolkit_b200 import _native as nat

def run(model, dtype, shape, nsteps=2000):
    rng = np.random.default_rng(0)
    wm = rng.uniform(0.2, 0.7, shape); gm = 1 - wm
    u0 = rng.uniform(0, 0.2, shape)
    if model == nat.MODEL_FK_2C:
        p = nat.make_params(model, dtype, shape, (1, 1, 1), 0.01, rho=0.1, Dw=1.0, ratio=10, th_matter=0.1, th_necro=0.9,
                            Ds=1.3, lambda_np=0.35, sigma_np=0.5, lambda_s=0.05)
    else:
        p = nat.make_params(model, dtype, shape, (1, 1, 1), 0.01, rho=0.1, Dw=1.0, ratio=10, th_matter=0.1)
    with nat.Sim(p) as sim:
        sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm); sim.upload(nat.FIELD_U, u0)
        if model == nat.MODEL_FK_2C:
            sim.upload(nat.FIELD_N, np.zeros(shape)); sim.upload(nat.FIELD_S, np.ones(shape))
        sim.prepare()
        sim.run(200); sim.sync()
        sim.run(nsteps); st = sim.sync()
        return st.loop_ms * 1e3 / nsteps

for shape in [(8, 8, 8), (32, 32, 32), (64, 64, 64), (96, 112, 96), (145, 178, 139), (200, 240, 200), (288, 354, 277)]:
    n = int(np.prod(shape))
    out = []
    for model, name in ((nat.MODEL_FK, "fk"), (nat.MODEL_FK_2C, "2c")):
        for dt, dn in ((nat.F64, "f64"), (nat.F32, "f32")):
            us = run(model, dt, shape, 400 if n > 5e6 else 2000)
            out.append(f"{name}/{dn} {us:7.1f}us {n / us / 1e3:6.1f}GLUPS")
    print(shape, n, " | ".join(out), flush=True)
