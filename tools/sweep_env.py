"""A/B sweep of library tunables (environment variables read at tgtk_sim_create) in one process:
per-step device time and HBM-roofline fraction of the step kernels on a given box.

    python tools/sweep_env.py TGTK_L2PF=0,1,2,3 [--shape 145,178,139] [--models 2c,fk,dti] [--dtypes f64,f32]
"""
import argparse, itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tumorgrowthtoolkit_b200 import _native as nat

WORDS = {"fk": 4, "dti": 5, "2c": 8, "dtifull": 8}
MODELS = {"fk": nat.MODEL_FK, "2c": nat.MODEL_FK_2C, "dti": nat.MODEL_DTI, "dtifull": nat.MODEL_DTI_FULL}


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return json.load(open(p))["hbm_gbs"]
    except Exception:
        return 6561.9


def run(model, dtype, shape, nsteps):
    rng = np.random.default_rng(0)
    wm = rng.uniform(0.2, 0.7, shape); gm = 1 - wm
    u0 = rng.uniform(0, 0.2, shape)
    kw = dict(rho=0.1, Dw=1.0, ratio=10, th_matter=0.1)
    if model == "2c":
        kw.update(th_necro=0.9, Ds=1.3, lambda_np=0.35, sigma_np=0.5, lambda_s=0.05)
    p = nat.make_params(MODELS[model], dtype, shape, (1, 1, 1), 0.01, **kw)
    with nat.Sim(p) as sim:
        if model in ("dti", "dtifull"):
            for c in range(3):
                sim.upload(nat.FIELD_COEF0 + c, rng.uniform(0.1, 1.0, shape))
            if model == "dtifull":
                for c in range(3, 6):
                    sim.upload(nat.FIELD_COEF0 + c, rng.uniform(-0.05, 0.05, shape))
        else:
            sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm)
        sim.upload(nat.FIELD_U, u0)
        if model == "2c":
            sim.upload(nat.FIELD_N, np.zeros(shape)); sim.upload(nat.FIELD_S, np.ones(shape))
        sim.prepare()
        sim.run(100); sim.sync()
        best = 1e30
        for _ in range(3):
            sim.run(nsteps); st = sim.sync()
            best = min(best, st.loop_ms * 1e3 / nsteps)
        return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("settings", nargs="*", help="NAME=v1,v2,... (cartesian product)")
    ap.add_argument("--shape", default="145,178,139")
    ap.add_argument("--models", default="2c,fk,dti")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--steps", type=int, default=600)
    a = ap.parse_args()
    shape = tuple(int(v) for v in a.shape.split(","))
    n = int(np.prod(shape))
    names = [s.split("=")[0] for s in a.settings]
    values = [s.split("=")[1].split(",") for s in a.settings]
    peak = peak_gbs()
    for combo in itertools.product(*values):
        for k, v in zip(names, combo):
            os.environ[k] = v
        out = []
        for m in a.models.split(","):
            for d in a.dtypes.split(","):
                us = run(m, nat.F64 if d == "f64" else nat.F32, shape, a.steps)
                frac = n * WORDS[m] * (8 if d == "f64" else 4) / (us * 1e-6) / 1e9 / peak
                out.append(f"{m}/{d} {us:6.2f}us {frac:.3f}")
        print(" ".join(f"{k}={v}" for k, v in zip(names, combo)), "|", " | ".join(out), flush=True)


if __name__ == "__main__":
    main()
