#!/usr/bin/env python
"""Headline benchmark: voxel-updates/s (GLUPS) of the explicit time-stepping hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--model fk2c|fk|dti] [--precision fp64|fp32]

Workload (BASELINE.json configs[1]): FK_2c (proliferative / necrotic / nutrient) on the SRI
atlas 240x240x155 with the FK_2c_example.ipynb parameters, at resolution_factor 1.0 so the
working set (8 float64 arrays of the 145x178x139 cropped box = 230 MB) exceeds the 126 MB L2.
One bench "step" = one complete reference time loop (1340 explicit Euler steps at those
parameters, FK_2c.py:200) = 4.8e9 voxel-updates.

  value   device-resident throughput: state and tissue maps already in HBM, the loop's
          launches timed with CUDA events on the launching stream (tgtk_status.loop_ms).
  e2e     the same loop through the host-buffer C ABI: per step sim create, H2D of the five
          float64 input arrays from pinned host memory, coefficient set-up, the loop, D2H of
          P, N, S and the volume, sim destroy; wall clock around it.
  N > 1   one process per GPU, every rank steps its own patient (ensemble sharding, no
          collective on the data path); value = all ranks' voxel-updates / max-over-ranks time.

--impl reference times the CPU restatement of the reference loop (oracle/, OpenMP over all
host cores; the reference itself is single-threaded numpy and cannot travel to the GPU box)
on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {  # compulsory words per voxel-update x sizeof
    ("fk", "fp64"): 32, ("fk", "fp32"): 16, ("dti", "fp64"): 40, ("dti", "fp32"): 20,
    ("fk2c", "fp64"): 64, ("fk2c", "fp32"): 32,
}
FK2C = dict(Dw=0.9, rho=0.14, lambda_np=0.35, sigma_np=0.5, D_s=1.3, lambda_s=0.05, ratio=100.0,
            th_matter=0.1, th_necro=0.9, Nt_multiplier=8, T=100.0)      # FK_2c_example.ipynb:106-123
FK = dict(Dw=1.0, rho=0.10, ratio=10.0, th_matter=0.1, T=100.0)           # FK_exampleAtlas.ipynb:27-39
DTI = dict(Dw=1.0, rho=0.2, T=100.0)                                      # FK_DTI_example.ipynb:134-151


def atlas_inputs():
    """Cropped res-1.0 box of the SRI tissue atlas (uint8 labels under tests/golden/), or of
    the synthetic phantom of the same 240x240x155 shape when the fixture is absent."""
    from tumorgrowthtoolkit_b200 import _host, synthetic
    path = os.path.join(ROOT, "tests", "golden", "sri_tissue_labels.npz")
    if os.path.exists(path):
        wm, gm = synthetic.tissue_from_labels(np.load(path)["seg"])
        source = "SRI tissue atlas 240x240x155 (wm=label 3, gm=label 2)"
    else:
        wm, gm = synthetic.brain_phantom((240, 240, 155))
        source = "synthetic phantom 240x240x155"
    grid = _host.LowResGrid(wm.shape, wm.shape, 1.0, 1.0, 1.0, 1.0, (0.3, 0.7, 0.5))
    box = _host.bounding_box(gm + wm >= 0.5, 2, wm.shape)
    u0 = _host.seed_gaussian(grid, 1.0)
    c = lambda a: np.ascontiguousarray(_host.crop(a, box))
    return c(wm), c(gm), c(u0), source


def build_problem(model):
    wm, gm, u0, source = atlas_inputs()
    h = (1.0, 1.0, 1.0)
    if model == "fk2c":
        q = FK2C
        nt = q["T"] * max(q["Dw"], q["D_s"]) / min(h) ** 2 * q["Nt_multiplier"] + 300   # FK_2c.py:200
        fields = dict(P=u0, N=np.zeros_like(u0), S=np.where(wm + gm >= q["th_matter"], 1.0, 0.0))
    elif model == "fk":
        q = FK
        nt = max(q["T"] * q["Dw"] / min(h) ** 2 * 8 + 100, q["T"] * q["rho"] * 1.1)     # FK.py:185
        fields = dict(u=u0)
    else:
        q = DTI
        rgb = np.stack([wm + 0.5 * gm, 0.7 * wm + 0.5 * gm, 0.4 * wm + 0.5 * gm], axis=-1)
        rgb /= rgb[rgb.max(axis=-1) > 0].mean()
        nt = max(q["T"] * q["Dw"] * rgb.max() / min(h) ** 2 * 8 + 100, q["T"] * q["rho"] * 1.1)  # FK_DTI.py:156
        fields = dict(u=u0, rgb=rgb)
    nsteps = int(np.ceil(nt))
    return dict(model=model, q=q, wm=wm, gm=gm, fields=fields, h=h, dt=q["T"] / nt, nsteps=nsteps,
                shape=wm.shape, voxels=int(wm.size), source=source)


# ------------------------------------------------------------------------------------------
def make_sim(pb, precision, device, pinned=None):
    from tumorgrowthtoolkit_b200 import _loop, _native as nat
    q, h = pb["q"], pb["h"]
    dt_code = _loop.dtype_code(precision)
    src = pinned if pinned is not None else dict(wm=pb["wm"], gm=pb["gm"], **pb["fields"])
    pin = pinned is not None
    if pb["model"] == "fk2c":
        p = nat.make_params(nat.MODEL_FK_2C, dt_code, pb["shape"], h, pb["dt"], rho=q["rho"], Dw=q["Dw"],
                            ratio=q["ratio"], th_matter=q["th_matter"], th_necro=q["th_necro"], Ds=q["D_s"],
                            lambda_np=q["lambda_np"], sigma_np=q["sigma_np"], lambda_s=q["lambda_s"])
        sim = nat.Sim(p, device)
        sim.upload(nat.FIELD_WM, src["wm"], pin)
        sim.upload(nat.FIELD_GM, src["gm"], pin)
        for f, k in ((nat.FIELD_P, "P"), (nat.FIELD_N, "N"), (nat.FIELD_S, "S")):
            sim.upload(f, src[k], pin)
        out_fields = (nat.FIELD_P, nat.FIELD_N, nat.FIELD_S)
    elif pb["model"] == "fk":
        p = nat.make_params(nat.MODEL_FK, dt_code, pb["shape"], h, pb["dt"], rho=q["rho"], Dw=q["Dw"],
                            ratio=q["ratio"], th_matter=q["th_matter"])
        sim = nat.Sim(p, device)
        sim.upload(nat.FIELD_WM, src["wm"], pin)
        sim.upload(nat.FIELD_GM, src["gm"], pin)
        sim.upload(nat.FIELD_U, src["u"], pin)
        out_fields = (nat.FIELD_U,)
    else:
        p = nat.make_params(nat.MODEL_DTI, dt_code, pb["shape"], h, pb["dt"], rho=q["rho"], Dw=q["Dw"],
                            small_volume=1e-6)
        sim = nat.Sim(p, device)
        for a in range(3):
            sim.upload(nat.FIELD_COEF0 + a, src["c%d" % a] if pin else pb["fields"]["rgb"][..., a] * q["Dw"], pin)
        sim.upload(nat.FIELD_U, src["u"], pin)
        out_fields = (nat.FIELD_U,)
    sim.prepare()
    return sim, out_fields


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()            # the first sample takes nvidia-smi ~0.3 s to produce
            while not self.lines and time.perf_counter() - t0 < 3.0:
                time.sleep(0.02)
            self.lines.clear()                  # keep only samples taken while the region runs
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(model, precision):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(f"{model}_{precision}")
    return None


# ------------------------------------------------------------------------------------------
def cpu_sample(pb, seconds_target=12.0, max_steps=2000):
    """Times the oracle's loop body (step + volume) on the host cores for a bounded number
    of time steps of the SAME cropped box; returns (GLUPS, threads, steps, seconds)."""
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    q, (dx, dy, dz), dt = pb["q"], pb["h"], pb["dt"]
    f = pb["fields"]
    if pb["model"] == "fk2c":
        st = orc._Fk2cTissueStepper(pb["wm"], pb["gm"], q["th_matter"], q["th_necro"], q["Dw"], q["ratio"], q["D_s"])
        state = (f["P"].copy(), f["N"].copy(), f["S"].copy())

        def one(state):
            P, N, S = st.step(*state, q["rho"], q["lambda_np"], q["sigma_np"], q["lambda_s"], dt, dx, dy, dz)
            _ = dx * dy * dz * orc.lib().tgo_sum(orc._p(P), P.size) + orc.lib().tgo_sum(orc._p(N), N.size)
            return (P, N, S)
    elif pb["model"] == "fk":
        Dm = orc.faces(pb["wm"], pb["gm"], q["th_matter"], q["Dw"], q["ratio"])
        state = f["u"].copy()

        def one(u):
            u = orc.fk_step_faces(u, Dm, q["rho"], dt, dx, dy, dz)
            _ = dx * dy * dz * orc.lib().tgo_sum(orc._p(u), u.size)
            return u
    else:
        D = {}
        for a, n in enumerate("xyz"):
            D["D_minus_" + n] = np.ascontiguousarray(f["rgb"][..., a] * q["Dw"])
            D["D_plus_" + n] = D["D_minus_" + n]
        state = f["u"].copy()

        def one(u):
            u = orc.fk_step(u, D, q["rho"], dt, dx, dy, dz)
            _ = dx * dy * dz * orc.lib().tgo_sum(orc._p(u), u.size)
            return u
    state = one(state)                       # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        state = one(state)
        n += 1
        if time.perf_counter() - t0 >= seconds_target:
            break
    sec = time.perf_counter() - t0
    return pb["voxels"] * n / sec / 1e9, threads, n, sec


def run_reference(args, pb, rank):
    if rank != 0:
        return
    per_step_target = max(2.0, min(15.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(pb, seconds_target=per_step_target / 4, max_steps=2000)
    vals, nsum, ssum = [], 0, 0.0
    for _ in range(args.steps):
        g, threads, n, sec = cpu_sample(pb, seconds_target=per_step_target, max_steps=2000)
        vals.append(g)
        nsum += n
        ssum += sec
    value = pb["voxels"] * nsum / ssum / 1e9
    sample = (f"{nsum} time steps of the {pb['shape'][0]}x{pb['shape'][1]}x{pb['shape'][2]} box over {args.steps} "
              f"bench steps ({ssum:.1f} s); per-voxel C restatement of the numpy loop, OpenMP")
    line = {
        "impl": "reference", "metric": "voxel-updates/s", "value": value, "unit": "GLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * ssum / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(pb, "fp64", args.gpus),
        "cpu_baseline": {"value": value, "unit": "GLUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(pb, precision, gpus):
    X, Y, Z = pb["shape"]
    return {"workload": f"{pb['model']} on {pb['source']}, notebook parameters, resolution_factor 1.0, cropped box "
                        f"{X}x{Y}x{Z} = {pb['voxels']} voxels, {pb['nsteps']} time steps per bench step",
            "precision": precision, "time_steps_per_step": pb["nsteps"], "voxels": pb["voxels"],
            "l2": "working set larger than L2 (no flush needed)" if pb["voxels"] * ALGO_BYTES[(pb["model"], precision)] > 126e6
                  else "working set fits L2; L2 flushed between bench steps by the state reset + a 256 MB memset",
            "parallelism": f"ensemble x{gpus} (one independent patient per GPU, no collective)"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="fk2c", choices=["fk2c", "fk", "dti"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    pb = build_problem(args.model)
    if args.impl == "reference":
        run_reference(args, pb, rank)
        return

    import torch
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    device = local_rank
    torch.cuda.set_device(device)
    from tumorgrowthtoolkit_b200 import _native as nat
    nat.lib()                                  # fails loudly if the CUDA library is missing

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------
    sim, out_fields = make_sim(pb, args.precision, device)
    tuning = sim.tuning()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{device}")
    launches0 = sim.sync().kernel_launches

    def one_step():
        sim.reset()
        flush.zero_()
        torch.cuda.synchronize()
        sim.run(pb["nsteps"])
        st = sim.sync()
        assert st.steps_run == pb["nsteps"], (st.steps_run, st.stop_reason)
        return st.loop_ms

    for _ in range(args.warmup):
        one_step()
    barrier()
    with ClockSampler(device) as clocks:
        t0 = time.perf_counter()
        dev_ms = sum(one_step() for _ in range(args.steps))
        barrier()
        wall = time.perf_counter() - t0
        if wall < 0.6:                          # too short for a 200 ms sampler: keep the GPU under the same
            t1 = time.perf_counter()            # load (untimed) until a few samples exist
            while time.perf_counter() - t1 < 0.8:
                one_step()
    st = sim.sync()
    launches = int(st.kernel_launches - launches0) // (args.warmup + args.steps) * args.steps
    final_volume = st.last_volume
    sim.close()

    # ---- end-to-end arm: host buffers through the C ABI ------------------------------------
    names = ["wm", "gm"] + ([k for k in pb["fields"] if k != "rgb"])
    host = {k: (pb[k] if k in ("wm", "gm") else pb["fields"][k]) for k in names}
    if pb["model"] == "dti":
        del host["wm"], host["gm"]
        for a in range(3):
            host["c%d" % a] = pb["fields"]["rgb"][..., a] * pb["q"]["Dw"]
    pinned = {}
    for k, v in host.items():
        tns = torch.empty(v.shape, dtype=torch.float64, pin_memory=True)
        tns.numpy()[...] = v
        pinned[k] = tns.numpy()
    outs = [torch.empty(pb["shape"], dtype=torch.float64, pin_memory=True).numpy() for _ in out_fields]
    h2d = sum(v.nbytes for v in pinned.values())
    d2h = sum(o.nbytes for o in outs) + 8

    def e2e_step():
        s2, of = make_sim(pb, args.precision, device, pinned=pinned)
        s2.run(pb["nsteps"])
        stt = s2.sync()
        for f, o in zip(of, outs):
            s2.download(f, o)
        s2.close()
        return stt.last_volume

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_vol = e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t0

    # ---- reduce over ranks -----------------------------------------------------------------
    tmax = torch.tensor([dev_ms, e2e_wall, wall], dtype=torch.float64, device=f"cuda:{device}")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_wall_max, wall_max = (float(x) for x in tmax.tolist())
    updates = float(pb["voxels"]) * pb["nsteps"] * args.steps * world
    value = updates / (dev_ms_max * 1e-3) / 1e9
    e2e_value = updates / e2e_wall_max / 1e9

    if rank == 0:
        peak, peak_src = measured_peak()
        bpu = ALGO_BYTES[(pb["model"], args.precision)]
        launch_ms = dev_ms / (pb["nsteps"] * args.steps)          # rank 0's average launch duration
        achieved = pb["voxels"] * bpu / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": "voxel-updates/s", "value": value, "unit": "GLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.precision == "fp64" else "f32",
            "data": "synthetic", "config": workload_config(pb, args.precision, world),
            "e2e": {"value": e2e_value, "unit": "GLUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_wall_max / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(pb["model"], args.precision), "peak_source": peak_src,
                         "algorithmic_bytes_per_voxel_update": bpu, "kernel_us": 1e3 * launch_ms,
                         "kernel": {"fk2c": "fk2c_step", "fk": "fk_step", "dti": "dti_step"}[pb["model"]] + "_v%d" % tuning["variant"],
                         "launch": tuning},
            "clocks": clocks.summary(),
            "wall_s_timed_region": wall_max, "final_volume": final_volume, "e2e_final_volume": e2e_vol,
        }
        if world == 1 and not args.no_cpu_baseline:
            g, threads, n, sec = cpu_sample(pb, seconds_target=12.0)
            line["cpu_baseline"] = {"value": g, "unit": "GLUPS", "cores": threads, "kind": "port",
                                    "sample": f"{n} time steps of the same {pb['shape']} box in {sec:.1f} s "
                                              "(oracle C restatement, OpenMP over all host threads)"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
