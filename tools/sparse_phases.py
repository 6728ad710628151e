"""Cycles per phase of a plane step of fk2c_step_sp, per warp, from a -DTGTK_SP_PHASES build of the library:
    nvcc ... -DTGTK_SP_PHASES -shared -o build_ab/libtgtk_phases.so tumorgrowthtoolkit_b200/csrc/tgtk_api.cu
    TGTK_LIB_PATH=build_ab/libtgtk_phases.so python tools/sparse_phases.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

pb = bench.build_problem("fk2c", 1.0)
sim, _ = bench.make_sim(pb, "fp64", 0)
sim.run(200); sim.sync()
out = os.path.join(tempfile.mkdtemp(), "phases.bin")
os.environ["TGTK_SP_PHASES_OUT"] = out
for _ in range(2):
    t, planes = sim.sparse_trace()
ph = np.fromfile(out, dtype=np.uint64).reshape(len(t), 4, 8).astype(np.float64)
n = ph[..., 6]
live = n > 0
names = ["top: publish, issue loads, prefetch, H / reaction", "barrier", "shared-memory reads + in-plane faces", "wait for plane x+1, x faces",
         "update P, N, S", "stores + volume sums"]
tot = ph[..., :6].sum(axis=(0, 1))
steps = n.sum()
print(f"{len(t)} items, {int(steps)} live warp-steps; cycles per live warp-step: {tot.sum() / steps:.0f}")
for i, nm in enumerate(names):
    per = ph[..., i][live] / n[live]
    print(f"  {nm:55s} {tot[i] / steps:7.0f} cycles ({100 * tot[i] / tot.sum():4.1f} %)  p10 {np.percentile(per, 10):6.0f} p90 {np.percentile(per, 90):6.0f}")
loop_ns = (t[:, 2] - t[:, 1]).astype(float)
npl = (planes[:, 1] - planes[:, 0]).astype(float)
print(f"loop time per plane (all planes of an item, dead ones included): {np.mean(loop_ns / npl):.0f} ns")
sim.close()
