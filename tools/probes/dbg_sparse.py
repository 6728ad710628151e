import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from test_gpu_sparse import _problem, _env
from tumorgrowthtoolkit_b200 import _native as nat
shape = (40, 48, 70)
wm, gm, u0, S0 = _problem(shape)
for model in ("fk", "dti"):
    res = {}
    for sparse in (1, 0):
        with _env(TGTK_SPARSE=sparse, TGTK_VARIANT=4, TGTK_TB=0):
            if model == "fk":
                p = nat.make_params(nat.MODEL_FK, nat.F64, shape, (1, 1, 1), 0.03, rho=0.2, Dw=1.0, ratio=10.0, th_matter=0.1)
                sim = nat.Sim(p, 0)
                sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm); sim.upload(nat.FIELD_U, u0)
            else:
                p = nat.make_params(nat.MODEL_DTI, nat.F64, shape, (1, 1, 1), 0.03, rho=0.2, Dw=1.0, small_volume=1e-6)
                sim = nat.Sim(p, 0)
                rng = np.random.default_rng(5)
                rgb = rng.uniform(0.05, 1.0, shape + (3,)) * (wm + gm >= 0.1)[..., None]
                for a in range(3):
                    sim.upload(nat.FIELD_COEF0 + a, rgb[..., a])
                sim.upload(nat.FIELD_U, u0 * (wm + gm >= 0.1))
        sim.prepare()
        sim.run(40)
        st = sim.sync()
        res[sparse] = (sim.download(nat.FIELD_U), sim.volumes(0, 40), st.steps_run, st.stop_reason, st.last_volume, sim.sparse_info(), sim.tuning())
        sim.close()
    a, b = res[1], res[0]
    print(model, "state equal", np.array_equal(a[0], b[0]), "maxdiff", np.abs(a[0] - b[0]).max(), "steps", a[2], b[2], "stop", a[3], b[3], "last", a[4], b[4])
    print("  vols sparse", a[1][:3], a[1][-2:], "dense", b[1][:3], b[1][-2:])
    print("  info", a[5], a[6])
    d = np.argwhere(a[0] != b[0])
    if len(d):
        print("  first diffs", d[:5], "count", len(d))
