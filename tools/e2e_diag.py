import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
pb = bench.build_problem('fk2c')
from tumorgrowthtoolkit_b200 import _native as nat
host = dict(wm=pb['wm'], gm=pb['gm'], **pb['fields'])
pinned = {}
for k, v in host.items():
    t = torch.empty(v.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = v; pinned[k] = t.numpy()
outs = [torch.empty(pb['shape'], dtype=torch.float64, pin_memory=True).numpy() for _ in range(3)]
for it in range(4):
    t0 = time.perf_counter()
    sim, of = bench.make_sim(pb, 'fp64', 0, pinned=pinned)
    t1 = time.perf_counter()
    sim.run(pb['nsteps']); t2 = time.perf_counter()
    st = sim.sync(); t3 = time.perf_counter()
    for f, o in zip(of, outs): sim.download(f, o)
    t4 = time.perf_counter()
    sim.close(); t5 = time.perf_counter()
    print(f"iter {it}: make_sim {t1-t0:.3f} enqueue {t2-t1:.3f} sync {t3-t2:.3f} download {t4-t3:.3f} close {t5-t4:.3f} loop_ms {st.loop_ms:.1f}")
