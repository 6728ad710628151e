"""Where the time of a sparse launch goes: per work item the block's start, prologue and loop times and its SM
(tgtk_sim_sparse_trace), summarised.   python tools/sparse_trace.py [--model fk2c] [--precision fp64]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="fk2c")
ap.add_argument("--precision", default="fp64")
args = ap.parse_args()
pb = bench.build_problem(args.model, 1.0)
sim, _ = bench.make_sim(pb, args.precision, 0)
sim.run(200); sim.sync()
for rep in range(2):
    t, planes = sim.sparse_trace()
n = len(t)
start, pro, end, sm = t[:, 0] / 1e3, t[:, 1] / 1e3, t[:, 2] / 1e3, t[:, 3]
npl = planes[:, 1] - planes[:, 0]
print(f"{n} items, planes per item: mean {npl.mean():.1f} min {npl.min()} max {npl.max()}; launch span {end.max():.1f} us")
print(f"prologue: mean {np.mean(pro - start):.2f} us, p90 {np.percentile(pro - start, 90):.2f}")
loop = end - pro
print(f"loop: mean {loop.mean():.2f} us; per plane: mean {np.mean(loop / npl):.3f} us, p10 {np.percentile(loop / npl, 10):.3f}, p90 {np.percentile(loop / npl, 90):.3f}")
print(f"block start: first wave (< 1 us) {np.sum(start < 1.0)}, p50 {np.percentile(start, 50):.1f}, p90 {np.percentile(start, 90):.1f}, max {start.max():.1f} us")
print(f"block end: p10 {np.percentile(end, 10):.1f} p50 {np.percentile(end, 50):.1f} p90 {np.percentile(end, 90):.1f} max {end.max():.1f} us")
busy = np.zeros(int(sm.max()) + 1)
last = np.zeros(int(sm.max()) + 1)
for i in range(n):
    busy[int(sm[i])] += end[i] - start[i]
    last[int(sm[i])] = max(last[int(sm[i])], end[i])
print(f"per SM: block-time sum mean {busy.mean():.1f} us (min {busy.min():.1f}, max {busy.max():.1f}); last block ends mean {last.mean():.1f} (min {last.min():.1f}, max {last.max():.1f})")
order = np.argsort(start)
for i in list(order[:3]) + list(order[n // 2:n // 2 + 3]) + list(order[-3:]):
    print(f"  item {i}: SM {sm[i]} planes {planes[i, 0]}..{planes[i, 1]} start {start[i]:.2f} prologue {pro[i] - start[i]:.2f} loop {loop[i]:.2f} ({loop[i] / npl[i]:.2f}/plane) end {end[i]:.2f}")
sim.close()
