"""A/B sweep of library tunables on the BENCHMARK problem (the cropped SRI atlas box of bench.py, not a random all-tissue
box): device time per step of whole loops, per setting, interleaved over several rounds so that clock drift hits all alike.

    python tools/sweep_atlas.py TGTK_SPARSE=0,1 TGTK_SP_WAVES=1,2,4 [--model fk2c] [--precision fp64] [--steps 400] [--rounds 3]
"""
import argparse, itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

ap = argparse.ArgumentParser()
ap.add_argument("settings", nargs="*")
ap.add_argument("--model", default="fk2c")
ap.add_argument("--precision", default="fp64")
ap.add_argument("--res", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--planes", default="", help="a:b -- only these planes of axis 0 (what one rank of a slab run owns), as a periodic box")
args = ap.parse_args()
keys = [s.split("=")[0] for s in args.settings]
vals = [s.split("=")[1].split(",") for s in args.settings]
pb = bench.build_problem(args.model, args.res)
if args.planes:
    a, b = (int(v) for v in args.planes.split(":"))
    cut = lambda v: None if v is None else np.ascontiguousarray(v[a:b])
    pb["wm"], pb["gm"] = cut(pb["wm"]), cut(pb["gm"])
    pb["fields"] = {k: cut(v) for k, v in pb["fields"].items()}
    pb["shape"] = next(iter(pb["fields"].values())).shape[:3]
    pb["voxels"] = int(np.prod(pb["shape"]))
print(f"# {args.model} {args.precision} box {pb['shape']} {pb['voxels']} voxels, {args.steps} steps per loop", flush=True)
combos = list(itertools.product(*vals)) or [()]
sims = []
for combo in combos:
    for k, v in zip(keys, combo):
        os.environ[k] = v
    sim, _ = bench.make_sim(pb, args.precision, 0)
    sim.run(64); sim.sync(); sim.reset()
    sims.append(sim)
    for k in keys:
        os.environ.pop(k, None)
times = [[] for _ in combos]
for r in range(args.rounds):
    for i, sim in enumerate(sims):
        sim.run(args.steps)
        st = sim.sync()
        times[i].append(st.loop_ms * 1e3 / args.steps)
        sim.reset()
for combo, sim, t in zip(combos, sims, times):
    sp = sim.sparse_info()
    tun = sim.tuning()
    us = float(np.median(t))
    print(" ".join(f"{k}={v}" for k, v in zip(keys, combo)) or "(default)",
          f"| {us:7.2f} us/step (min {min(t):.2f})  {pb['voxels'] / us / 1e3:7.1f} GLUPS | items {sp['items']} live {sp['live_voxels'] / sp['voxels']:.3f}"
          f" variant {tun['variant']} grid {tun['grid']}", flush=True)
    sim.close()
