"""What one rank of a decomposed run costs without any other rank: the ordinary periodic loop on a (planes, Y, Z) box
against the fused slab driver closing its ring on the rank itself (same planes + 2 ghosts, in-kernel pushes, epoch
words, chunk commit), per time step.  python tools/slab_probe.py [--planes 36] [--yz 354,277] [--model dti|fk2c] [--steps 400]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--planes", type=int, default=36)
    ap.add_argument("--yz", default="354,277")
    ap.add_argument("--model", default="dti")
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--precision", default="fp64")
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from tumorgrowthtoolkit_b200 import _native as nat, slab
    from tumorgrowthtoolkit_b200._loop import dtype_code
    Y, Z = (int(v) for v in a.yz.split(","))
    shape = (a.planes, Y, Z)
    rng = np.random.default_rng(0)
    u0 = rng.uniform(0, 0.2, shape)
    out = {"shape": shape, "model": a.model, "steps": a.steps}
    if a.model == "dti":
        rgb = rng.uniform(0.1, 1.0, shape + (3,))
        p = nat.make_params(nat.MODEL_DTI, dtype_code(a.precision), shape, (1, 1, 1), 0.01, rho=0.1, Dw=1.0)
        sim = nat.Sim(p)
        for c in range(3):
            sim.upload(nat.FIELD_COEF0 + c, rgb[..., c])
        sim.upload(nat.FIELD_U, u0)
        sess = lambda: slab.session_dti(u0, rgb, 1.0, 0.1, 0.01, 1, 1, 1, precision=a.precision)
    else:
        wm = rng.uniform(0.2, 0.7, shape); gm = 1 - wm
        p = nat.make_params(nat.MODEL_FK_2C, dtype_code(a.precision), shape, (1, 1, 1), 0.01, rho=0.1, Dw=1.0, ratio=10, th_matter=0.1,
                            th_necro=0.9, Ds=1.3, lambda_np=0.35, sigma_np=0.5, lambda_s=0.05)
        sim = nat.Sim(p)
        sim.upload(nat.FIELD_WM, wm); sim.upload(nat.FIELD_GM, gm)
        sim.upload(nat.FIELD_P, u0); sim.upload(nat.FIELD_N, np.zeros(shape)); sim.upload(nat.FIELD_S, np.ones(shape))
        sess = lambda: slab.session_fk2c(u0, np.zeros(shape), np.ones(shape), wm, gm, 0.1, 0.9, 1.0, 10, 1.3, 0.1, 0.35, 0.5, 0.05, 0.01, 1, 1, 1,
                                         precision=a.precision)
    sim.prepare()
    out["plain_tuning"] = sim.tuning()
    sim.run(100); sim.sync()
    best = 1e30
    for _ in range(3):
        sim.run(a.steps); best = min(best, sim.sync().loop_ms)
    out["plain_us_per_step"] = 1e3 * best / a.steps
    sim.close()
    s = sess()
    out["slab_tuning"] = s.sim.tuning()
    s.run(100); s.sync_ms()
    best = 1e30
    for _ in range(3):
        s.reset(); s.run(a.steps); best = min(best, s.sync_ms())
    out["slab_self_ring_us_per_step"] = 1e3 * best / a.steps
    s.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
