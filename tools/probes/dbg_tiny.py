import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from test_gpu_sparse import _env, _args
from tumorgrowthtoolkit_b200 import _loop
from oracle import oracle as orc
shape = tuple(int(v) for v in sys.argv[1].split(","))
rng = np.random.default_rng(7)
X, Y, Z = shape
wm = rng.uniform(0.2, 0.8, shape); gm = 1.0 - wm
dead = np.zeros(shape, bool); dead[: X // 3] = True
dead[:, :, Z // 2:] |= rng.uniform(size=(X, Y, Z - Z // 2)) < 0.3
wm[dead] = 0.0; gm[dead] = 0.0
u0 = np.zeros(shape); u0[X // 2, Y // 2, Z // 4] = 0.8; u0[X // 2, Y // 2, (Z // 4 + 1) % Z] = 0.4
S0 = np.where(wm + gm >= 0.1, 1.0, 0.0)
for nsteps in (1, 2, 5, 50):
    out = {}
    for sparse in (1, 0):
        with _env(TGTK_SPARSE=sparse, TGTK_VARIANT=4):
            out[sparse] = _loop.fk2c_loop(*_args(wm, gm, u0, S0, nsteps))
    ref = orc.fk2c_loop(*_args(wm, gm, u0, S0, nsteps))
    for k in "PNS":
        a, b, r = out[1]["state"][k], out[0]["state"][k], ref["state"][k]
        d = np.argwhere(a != b)
        print(nsteps, k, "sparse!=dense:", len(d), d[:4].tolist(), "| sparse vs oracle", np.abs(a - r).max(), "dense vs oracle", np.abs(b - r).max(), "items", out[1]["sparse_items"])
