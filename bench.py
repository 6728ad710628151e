#!/usr/bin/env python
"""Headline benchmark: voxel-updates/s (GLUPS) of the explicit time-stepping hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--model fk2c|fk|dti] [--precision fp64|fp32|mixed]

N = 1 (BASELINE.json configs[1]): FK_2c (proliferative / necrotic / nutrient) on the SRI atlas 240x240x155 with
the FK_2c_example.ipynb parameters, at resolution_factor 1.0 so the working set (8 float64 arrays of the
145x178x139 cropped box = 230 MB) exceeds the 126 MB L2.  One bench "step" = one complete reference time loop
(1340 explicit Euler steps, FK_2c.py:200) = 4.8e9 voxel-updates.

  value    device-resident throughput: state and tissue maps already in HBM, the loop's launches timed with CUDA
           events on the launching stream (tgtk_status.loop_ms).
  e2e      the same loop through the host-buffer C ABI: per step sim create, H2D of the five float64 input arrays
           from pinned host memory, coefficient set-up, the loop, D2H of P, N, S and the volume, sim destroy.
  parity   the CPU oracle steps the SAME box for the same number of time steps (all 1340 when the host finishes
           them inside the time budget) and the final fields are compared with the GPU's: max |diff| per field
           and the Dice of the thresholded segmentation.
  ensemble BASELINE.json configs[3]: 1024 random FK_2c parameter sets (run_gen_FK_2c.py:5-26,41, seed 42000) at
           resolution_factor 0.65 sharded over the ranks, full-resolution outputs on the host, in patients/hour.

N > 1 (BASELINE.json configs[4], one process per GPU): FK_DTI on the 2x up-sampled atlas (480x480x310, cropped box
about 288x354x277 = 28 M voxels), axis 0 slab-decomposed over the N GPUs, boundary planes exchanged over NVLink
every time step.  One bench step = 400 time steps of the decomposed loop.  value = voxel-updates of the WHOLE grid
per second ("scaling": "strong"); the `slab` object carries the speed-up over the undecomposed loop on one GPU
and the max |diff| against it; `ensemble` as above.  N = 1 keeps configs[1] so that it agrees with BENCH.

--impl reference times the reference's CPU implementation of the same workload on the box's host cores: the
oracle's C restatement of the numpy loop (oracle/, OpenMP over every core the process may use; `kind: "port"`)
on a bounded sample, and -- when baseline/_ref holds the pip-installed reference -- the literal numpy methods
for a few steps on one core beside it (`cpu_baseline.numpy`).
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
    ("fk2c", "fp64"): 64, ("fk2c", "fp32"): 32, ("fk2c", "mixed"): 48,
}
FK2C = dict(Dw=0.9, rho=0.14, lambda_np=0.35, sigma_np=0.5, D_s=1.3, lambda_s=0.05, ratio=100.0,
            th_matter=0.1, th_necro=0.9, Nt_multiplier=8, T=100.0)      # FK_2c_example.ipynb:106-123
FK = dict(Dw=1.0, rho=0.10, ratio=10.0, th_matter=0.1, T=100.0)           # FK_exampleAtlas.ipynb:27-39
DTI = dict(Dw=1.0, rho=0.2, T=100.0)                                      # FK_DTI_example.ipynb:134-151, fixed-step variant
SEED_PCT = (0.3, 0.7, 0.5)                                                # FK_exampleAtlas.ipynb / FK_2c_example.ipynb
SEED_PCT_DTI = (0.49, 0.30, 0.7)                                          # FK_DTI_example.ipynb


def atlas_labels():
    path = os.path.join(ROOT, "tests", "golden", "sri_tissue_labels.npz")
    return np.load(path)["seg"] if os.path.exists(path) else None


def _scipy_zoom(a, factor, device=0):
    from scipy.ndimage import zoom
    return zoom(a, factor, order=1)


def build_problem(model, res=1.0, on_gpu=False, device=0):
    """Cropped low-resolution arrays the time loop runs on, built the way the solvers' solve() builds them
    (resample, seed, crop, step-count law).  res 1.0: the zoom is the identity.  on_gpu: resample with the
    library's zoom kernel (identical to scipy to 1e-14) instead of scipy -- the reference arm never touches a GPU."""
    from tumorgrowthtoolkit_b200 import _host, synthetic
    seg = atlas_labels()
    if seg is not None:
        wm, gm = synthetic.tissue_from_labels(seg)
        source = "SRI tissue atlas 240x240x155 (wm=label 3, gm=label 2)"
    else:
        wm, gm = synthetic.brain_phantom((240, 240, 155))
        seg = np.where(wm > 0.5, 3, np.where(gm > 0.5, 2, 0))
        source = "synthetic phantom 240x240x155"
    zoomer = (lambda a, f: _host.resample_linear(a, f, device=device)) if on_gpu else _scipy_zoom
    rs = (lambda a: a) if res == 1.0 else (lambda a: zoomer(a, res))
    c = lambda a, box: np.ascontiguousarray(_host.crop(a, box))
    if model == "dti":
        # SURVEY.md 8(d) input 3: synthetic lower-6 tensor volume -> get_tensor_from_lower6 -> makeXYZ_rgb_from_tensor
        from tumorgrowthtoolkit_b200.FK_DTI import tools
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):              # tools.py:95 prints the shape
            tensors = tools.get_tensor_from_lower6(synthetic.dti_lower6_field(seg))
        rgb = tools.makeXYZ_rgb_from_tensor(tensors)
        del tensors
        q = DTI
        rgb_max = float(np.max(rgb))                                  # FK_DTI.py:156: max over the FULL-resolution rgb
        pct = synthetic.seed_inside(seg == 3, SEED_PCT_DTI)
        rgb_lo = rgb if res == 1.0 else np.stack([zoomer(np.ascontiguousarray(rgb[..., a]), res) for a in range(3)], axis=-1)
        grid = _host.LowResGrid(rgb_lo.shape, rgb.shape, res, 1.0, 1.0, 1.0, pct)
        box = _host.bounding_box(np.max(rgb_lo, axis=-1) > 0, 2, rgb_lo.shape[:3])      # FK_DTI.py:169-171
        u0 = c(_host.seed_gaussian(grid, 1.0), box)
        h = grid.h
        nt = max(q["T"] * q["Dw"] * rgb_max / min(h) ** 2 * 8 + 100, q["T"] * q["rho"] * 1.1)
        fields = dict(u=u0, rgb=c(rgb_lo, box))
        cw = cg = None
        source = "synthetic FSL_HCP1065-shaped tensor volume (240,240,155,6) on the " + source
    else:
        wml, gml = rs(wm), rs(gm)
        grid = _host.LowResGrid(wml.shape, wm.shape, res, 1.0, 1.0, 1.0, SEED_PCT)
        box = _host.bounding_box(gml + wml >= 0.5, 2, wml.shape)
        cw, cg = c(wml, box), c(gml, box)
        u0 = c(_host.seed_gaussian(grid, 1.0), box)
        h = grid.h
        if model == "fk2c":
            q = FK2C
            nt = q["T"] * max(q["Dw"], q["D_s"]) / min(h) ** 2 * q["Nt_multiplier"] + 300       # FK_2c.py:200
            fields = dict(P=u0, N=np.zeros_like(u0), S=np.where(cw + cg >= q["th_matter"], 1.0, 0.0))
        else:
            q = FK
            nt = max(q["T"] * q["Dw"] / min(h) ** 2 * 8 + 100, q["T"] * q["rho"] * 1.1)         # FK.py:185
            fields = dict(u=u0)
    nsteps = int(np.ceil(nt))
    return dict(model=model, q=q, wm=cw, gm=cg, fields=fields, h=h, dt=q["T"] / nt, nsteps=nsteps,
                shape=u0.shape, voxels=int(u0.size), source=source, res=res)


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


def host_threads():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is not a property of the box)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------
class OracleStepper:
    """The oracle's loop body (step + volume, FK.py:212-215 / FK_2c.py:229-235 / FK_DTI.py:196-198) on the
    problem's cropped box, OpenMP over `threads` cores."""

    def __init__(self, pb, threads=None):
        from oracle import oracle as orc
        orc.build()
        self.orc, self.pb = orc, pb
        self.threads = threads or host_threads()
        orc.set_threads(self.threads)
        q, f = pb["q"], pb["fields"]
        if pb["model"] == "fk2c":
            self.st = orc._Fk2cTissueStepper(pb["wm"], pb["gm"], q["th_matter"], q["th_necro"], q["Dw"], q["ratio"], q["D_s"])
            self.state = (f["P"].copy(), f["N"].copy(), f["S"].copy())
        elif pb["model"] == "fk":
            self.Dm = orc.faces(pb["wm"], pb["gm"], q["th_matter"], q["Dw"], q["ratio"])
            self.state = f["u"].copy()
        else:
            self.D = {}
            for a, n in enumerate("xyz"):
                self.D["D_minus_" + n] = np.ascontiguousarray(f["rgb"][..., a] * q["Dw"])
                self.D["D_plus_" + n] = self.D["D_minus_" + n]
            self.state = f["u"].copy()
        self.volume = None

    def step(self):
        orc, pb, q = self.orc, self.pb, self.pb["q"]
        (dx, dy, dz), dt = pb["h"], pb["dt"]
        if pb["model"] == "fk2c":
            P, N, S = self.st.step(*self.state, q["rho"], q["lambda_np"], q["sigma_np"], q["lambda_s"], dt, dx, dy, dz)
            self.volume = dx * dy * dz * orc.lib().tgo_sum(orc._p(P), P.size) + orc.lib().tgo_sum(orc._p(N), N.size)
            self.state = (P, N, S)
        else:
            u = (orc.fk_step_faces(self.state, self.Dm, q["rho"], dt, dx, dy, dz) if pb["model"] == "fk"
                 else orc.fk_step(self.state, self.D, q["rho"], dt, dx, dy, dz))
            self.volume = dx * dy * dz * orc.lib().tgo_sum(orc._p(u), u.size)
            self.state = u

    def fields(self):
        return dict(zip("PNS", self.state)) if self.pb["model"] == "fk2c" else {"u": self.state}


def cpu_sample(pb, seconds_target=12.0, max_steps=None, threads=None):
    """Times the oracle's loop body on the host cores from the problem's initial state for at most `max_steps`
    (default: the full loop) time steps or `seconds_target` seconds; returns (GLUPS, threads, steps, seconds, stepper)."""
    st = OracleStepper(pb, threads)
    max_steps = pb["nsteps"] if max_steps is None else max_steps
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        st.step()
        n += 1
        if time.perf_counter() - t0 >= seconds_target:
            break
    sec = time.perf_counter() - t0
    return pb["voxels"] * n / sec / 1e9, st.threads, n, sec, st


def segmentation_map(P, N, th_necro_n=0.12, th_enhancing_p=0.4, th_edema_p=0.15):
    """create_segmentation_map, synthetic_gens/generatePatient_FK_2c.py:67-83 (argparse defaults :176-178)."""
    seg = np.zeros(P.shape, dtype=np.int8)
    seg[(P >= th_edema_p) & (P < th_enhancing_p)] = 3
    seg[P >= th_enhancing_p] = 1
    seg[N > th_necro_n] = 4
    return seg


def dice_of(got, ref, model):
    if model == "fk2c":
        a, b = segmentation_map(got["P"], got["N"]), segmentation_map(ref["P"], ref["N"])
        labels = (1, 3, 4)
    else:
        a, b = (got["u"] >= 0.5).astype(np.int8), (ref["u"] >= 0.5).astype(np.int8)
        labels = (1,)
    out = []
    for lab in labels:
        x, y = a == lab, b == lab
        den = int(x.sum()) + int(y.sum())
        out.append(1.0 if den == 0 else 2.0 * int((x & y).sum()) / den)
    return min(out)


def parity_vs_oracle(pb, sim, out_fields, stepper, n):
    """GPU state after exactly n steps from the initial state against the oracle's after the same n steps."""
    sim.reset()
    sim.run(n)
    st = sim.sync()
    names = "PNS" if pb["model"] == "fk2c" else "u"
    got = {k: sim.download(f) for k, f in zip(names, out_fields)}
    ref = stepper.fields()
    per_field = {k: float(np.max(np.abs(got[k] - ref[k]))) for k in names}
    return {"steps_compared": n, "of_steps": pb["nsteps"], "max_abs": max(per_field.values()), "max_abs_per_field": per_field,
            "dice": dice_of(got, ref, pb["model"]),
            "volume_rel_diff": abs(st.last_volume - stepper.volume) / max(abs(stepper.volume), 1e-300),
            "against": "oracle (C restatement of the numpy loop, pinned on reference-generated vectors)"}


def numpy_reference_sample(pb, steps=3):
    """The literal reference (pip-installed into baseline/_ref from /root/reference, unmodified): its own numpy
    update methods on the problem's cropped arrays for a few time steps on ONE core (numpy's elementwise ops are
    single-threaded), BASELINE.md 4 step 2.  None when baseline/_ref is absent."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "TumorGrowthToolkit")):
        return None
    from unittest.mock import MagicMock
    for m in ("matplotlib", "matplotlib.pyplot", "nibabel", "dipy", "dipy.core", "dipy.core.gradients", "dipy.data", "dipy.io",
              "dipy.io.image", "dipy.reconst", "dipy.reconst.dti", "dipy.segment", "dipy.segment.mask"):
        sys.modules.setdefault(m, MagicMock())
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    q, (dx, dy, dz), dt, f = pb["q"], pb["h"], pb["dt"], pb["fields"]
    if pb["model"] == "fk2c":
        from TumorGrowthToolkit.FK_2c import Solver
        s = Solver({})
        P, N, S = f["P"].copy(), f["N"].copy(), f["S"].copy()
        D = s.get_D_with_necro(pb["wm"], pb["gm"], q["th_matter"], q["Dw"], q["ratio"], P, N, q["th_necro"])
        Ds = s.get_D(pb["wm"], pb["gm"], q["th_matter"], q["D_s"], 1)
        t0 = time.perf_counter()
        for _ in range(steps):                                               # FK_2c.py:229-235
            upd = s.FK_2c_update(P, N, S, D, q["rho"], q["lambda_np"], q["sigma_np"], Ds, q["lambda_s"], dt, dx, dy, dz)
            D = s.get_D_with_necro(pb["wm"], pb["gm"], q["th_matter"], q["Dw"], q["ratio"], P, upd["N"], q["th_necro"])
            _ = dx * dy * dz * np.sum(upd["P"]) + np.sum(upd["N"])
        sec = time.perf_counter() - t0
        what = "FK_2c_update + get_D_with_necro + np.sum"
    else:
        if pb["model"] == "fk":
            from TumorGrowthToolkit.FK import Solver
            s = Solver({})
            D = s.get_D(pb["wm"], pb["gm"], q["th_matter"], q["Dw"], q["ratio"])
        else:
            from TumorGrowthToolkit.FK_DTI import FK_DTI_Solver
            s = FK_DTI_Solver({})
            D = s.get_D_from_DTI(f["rgb"], q["Dw"])
        A = f["u"].copy()
        t0 = time.perf_counter()
        for _ in range(steps):                                               # FK.py:213-215
            A = s.FK_update(A, D, q["rho"], dt, dx, dy, dz)
            _ = dx * dy * dz * np.sum(A)
        sec = time.perf_counter() - t0
        what = "FK_update + np.sum"
    return {"value": pb["voxels"] * steps / sec / 1e9, "unit": "GLUPS", "cores": 1, "steps": steps, "seconds": sec,
            "what": f"unmodified reference from baseline/_ref: {what} on the same cropped box"}


def workload_config(pb, precision, gpus, slab=None):
    X, Y, Z = pb["shape"]
    cfg = {"workload": f"{pb['model']} on {pb['source']}, notebook parameters, resolution_factor {pb['res']}, cropped box "
                       f"{X}x{Y}x{Z} = {pb['voxels']} voxels, {pb['nsteps'] if slab is None else slab['steps']} time steps per bench step",
           "precision": precision, "time_steps_per_step": pb["nsteps"] if slab is None else slab["steps"], "voxels": pb["voxels"],
           "l2": "working set larger than L2 (no flush needed)" if pb["voxels"] * ALGO_BYTES[(pb["model"], precision)] / max(1, gpus if slab else 1) > 126e6
                 else "working set fits L2; L2 flushed between bench steps by the state reset + a 256 MB memset"}
    if slab is None:
        cfg["parallelism"] = f"ensemble x{gpus} (one independent patient per GPU, no collective)"
    else:
        cfg["parallelism"] = (f"slab x{gpus}: axis 0 split into {gpus} slabs, one process per GPU, driver '{slab['driver']}', halo "
                              f"{slab['halo']}, boundary planes exchanged over NVLink")
    return cfg


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world, emit):
    """Reference arm: rank 0 alone, every host core the process may use."""
    if rank != 0:
        return
    slab_cfg = world > 1
    model = "dti" if slab_cfg else args.model
    pb = build_problem(model, res=2.0 if slab_cfg else 1.0, on_gpu=False)
    threads = host_threads()
    budget = 150.0                                   # seconds for the whole arm on a box of unknown speed
    per_step_target = max(2.0, min(15.0, budget / max(1, args.steps + args.warmup)))
    cap = args.slab_steps if slab_cfg else pb["nsteps"]
    for _ in range(args.warmup):
        cpu_sample(pb, seconds_target=per_step_target / 4, max_steps=cap, threads=threads)
    nsum, ssum = 0, 0.0
    for _ in range(args.steps):
        g, threads, n, sec, _ = cpu_sample(pb, seconds_target=per_step_target, max_steps=cap, threads=threads)
        nsum += n
        ssum += sec
    value = pb["voxels"] * nsum / ssum / 1e9
    sample = (f"{nsum} time steps of the {pb['shape'][0]}x{pb['shape'][1]}x{pb['shape'][2]} box over {args.steps} "
              f"bench steps ({ssum:.1f} s); per-voxel C restatement of the numpy loop, OpenMP on {threads} threads")
    cpu = {"value": value, "unit": "GLUPS", "cores": threads, "kind": "port", "sample": sample}
    try:
        cpu["numpy"] = numpy_reference_sample(pb, steps=3 if pb["voxels"] < 8e6 else 1)
    except Exception as e:                            # the literal reference is a second, informative leg
        cpu["numpy"] = {"unavailable": f"{type(e).__name__}: {e}"}
    slab = dict(steps=args.slab_steps, driver=args.slab_driver, halo=args.slab_halo) if slab_cfg else None
    line = {
        "impl": "reference", "metric": "voxel-updates/s", "value": value, "unit": "GLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * ssum / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong" if slab_cfg else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(pb, "fp64", args.gpus, slab),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
def ensemble_leg(args, rank, world, device, dist):
    """BASELINE.json configs[3]: args.patients parameter sets sharded over the ranks, full-resolution outputs."""
    import torch
    from tumorgrowthtoolkit_b200 import ensemble, synthetic
    seg = atlas_labels()
    wm, gm = synthetic.tissue_from_labels(seg) if seg is not None else synthetic.brain_phantom((240, 240, 155))
    params = ensemble.draw_ensemble(args.patients)
    mine = [params[i] for i in ensemble.shard(params, rank, world)]
    runner = ensemble.EnsembleRunner(wm, gm, resolution_factor=0.65, device=device, precision=args.precision if args.precision != "mixed" else "fp64",
                                     concurrency=4, full_resolution=True)
    runner.run(mine[:4], keep_results=False)          # warm-up: kernels, pools, page-locked ring
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    _, stats = runner.run(mine, keep_results=False)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall, float(stats["voxel_updates"]), float(len(mine))], dtype=torch.float64, device=f"cuda:{device}")
    if dist is not None:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        wall = float(tm[0])
    updates, patients = float(t[1]), int(round(float(t[2])))
    return {"workload": f"{args.patients} FK_2c parameter sets (run_gen_FK_2c.py:5-26, seed 42000), resolution_factor 0.65, cropped box "
                        f"{runner.cw.shape[0]}x{runner.cw.shape[1]}x{runner.cw.shape[2]}, Nt_multiplier 10; a patient = P, N, S restored and zoomed "
                        "to 240x240x155 plus the label map, on the host (FET / Gibbs noise / NIfTI excluded)",
            "patients": patients, "n_gpus": world, "wall_s_max_over_ranks": wall, "patients_per_hour": 3600.0 * patients / wall,
            "glups": updates / wall / 1e9, "voxel_updates": updates, "full_resolution_outputs": True,
            "sharding": "sorted by step count, dealt round-robin; no data-path collective"}


def slab_session(args, pb, device):
    from tumorgrowthtoolkit_b200 import slab
    q, h, f = pb["q"], pb["h"], pb["fields"]
    kw = dict(halo=args.slab_halo, precision=args.precision, device=device, driver=args.slab_driver)
    if pb["model"] == "dti":
        return slab.session_dti(f["u"], f["rgb"], q["Dw"], q["rho"], pb["dt"], *h, **kw)
    return slab.session_fk2c(f["P"], f["N"], f["S"], pb["wm"], pb["gm"], q["th_matter"], q["th_necro"], q["Dw"], q["ratio"], q["D_s"],
                             q["rho"], q["lambda_np"], q["sigma_np"], q["lambda_s"], pb["dt"], *h, **kw)


def slab_leg(args, rank, world, device, dist, pb, steps, warmup, with_e2e=True):
    """BASELINE.json configs[4] for one model: (value GLUPS, ms per bench step, slab object, e2e object, launches, clocks, wall)."""
    import torch
    from tumorgrowthtoolkit_b200 import slab, _native as nat
    q, h, f = pb["q"], pb["h"], pb["fields"]
    nsteps = args.slab_steps
    names = "PNS" if pb["model"] == "fk2c" else ("u",)
    sess = slab_session(args, pb, device)

    def one_step():
        sess.reset()
        sess.run(nsteps)
        return sess.sync_ms()

    for _ in range(warmup):
        one_step()
    launches0 = sess.launches()
    with ClockSampler(device) as clocks:
        t0 = time.perf_counter()
        ms = sum(one_step() for _ in range(steps))
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        wall = time.perf_counter() - t0
        launches = sess.launches() - launches0
        extra = int(np.ceil(max(0.0, 0.8 - wall) / max(wall / steps, 1e-3)))      # same on every rank: wall was barrier-aligned ...
        if dist is not None:                                                       # ... but make sure of it
            te = torch.tensor([extra], device=f"cuda:{device}")
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            extra = int(te[0])
        for _ in range(extra):                  # too short for a 200 ms sampler: keep the same load on (untimed) until samples exist
            one_step()
    # parity: 13 steps of the decomposed loop against the undecomposed loop on rank 0's GPU, every field, every voxel
    sess.reset()
    sess.run(13)
    got = sess.gather()
    got = got if isinstance(got, dict) else {"u": got}
    vol13 = sess.volume()
    sess.close()
    t = torch.tensor([ms, wall, float(launches)], dtype=torch.float64, device=f"cuda:{device}")
    if dist is not None:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, wall = float(tm[0]), float(tm[1])
    launches = int(round(float(t[2])))
    updates = float(pb["voxels"]) * nsteps * steps
    value = updates / (ms * 1e-3) / 1e9
    info = {"workload": f"{pb['model']} {int(240 * pb['res'])}x{int(240 * pb['res'])}x{int(155 * pb['res'])} -> cropped "
                        f"{pb['shape'][0]}x{pb['shape'][1]}x{pb['shape'][2]}, {nsteps} time steps", "driver": args.slab_driver,
            "halo": args.slab_halo, "glups": value, "loop_ms_max_over_ranks": ms / steps, "us_per_time_step": 1e3 * ms / steps / nsteps,
            "planes_per_rank": [int(v) for v in sess.plan.sizes]}
    if rank == 0:
        # the same grid undecomposed on one GPU: speed-up denominator and parity reference
        sim, out_fields = make_sim(pb, args.precision, device)
        tun = sim.tuning()
        sim.run(13)
        st13 = sim.sync()
        ref = {k: sim.download(fld) for k, fld in zip(names, out_fields)}
        info["max_abs_diff_vs_undecomposed"] = max(float(np.max(np.abs(got[k] - ref[k]))) for k in names)
        info["parity_steps"] = 13
        info["volume_rel_diff"] = abs(vol13 - st13.last_volume) / abs(st13.last_volume)
        best = 1e30
        for _ in range(max(2, warmup)):
            sim.reset()
            sim.run(nsteps)
            best = min(best, sim.sync().loop_ms)
        sim.close()
        info["one_gpu_loop_ms"] = best
        info["one_gpu_glups"] = pb["voxels"] * nsteps / (best * 1e-3) / 1e9
        info["one_gpu_steps_per_launch"] = tun["steps_per_launch"]
        info["speedup_vs_1gpu"] = best / (ms / steps)
        info["efficiency_vs_1gpu"] = info["speedup_vs_1gpu"] / world
    if dist is not None:
        dist.barrier()
    e2e = None
    if with_e2e:
        # end to end: the public slab API on host arrays (create, upload of the extended slabs, the whole-length loop, download + gather)
        t0 = time.perf_counter()
        res = slab.dti_loop(f["u"], f["rgb"], q["Dw"], q["rho"], pb["dt"], *h, pb["nsteps"], halo=args.slab_halo, precision=args.precision,
                            device=device, driver=args.slab_driver)
        torch.cuda.synchronize()
        e2e_wall = time.perf_counter() - t0
        t = torch.tensor([e2e_wall], dtype=torch.float64, device=f"cuda:{device}")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_wall = float(t[0])
        plane_bytes = pb["shape"][1] * pb["shape"][2] * 8
        e2e = {"value": pb["voxels"] * pb["nsteps"] / e2e_wall / 1e9, "unit": "GLUPS",
               "h2d_bytes_per_step": int(4 * res["plan"].local_planes * plane_bytes * world), "d2h_bytes_per_step": int(pb["voxels"] * 8),
               "ms_per_step": 1e3 * e2e_wall, "time_steps": pb["nsteps"],
               "what": "slab.dti_loop on host arrays: sim create, H2D of every rank's extended slab (u + 3 coefficient arrays), the "
                       "full-length decomposed loop, D2H of the owned planes, all-gather"}
    return value, ms / steps, info, e2e, launches, clocks.summary(), wall


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
    ap.add_argument("--patients", type=int, default=1024, help="ensemble leg (0: skip)")
    ap.add_argument("--slab-steps", type=int, default=400)
    ap.add_argument("--slab-halo", type=int, default=1)
    ap.add_argument("--slab-driver", default="fused", choices=["fused", "p2p", "native", "torch"])
    ap.add_argument("--slab-res", type=float, default=2.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # the contract is ONE JSON line on stdout: libraries that print there (NCCL announces its version when a communicator
    # is created) are sent to stderr for the whole run, the line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(line):
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
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

    peak, peak_src = measured_peak()

    if world > 1:
        # ---- N > 1: the slab-decomposed grid is the headline, the ensemble rides along -------------------------
        if args.slab_driver != "fused" and args.slab_halo == 1:
            args.slab_halo = 4
        pb = build_problem("dti", res=args.slab_res, on_gpu=True, device=device)
        value, ms_step, info, e2e, launches, clocks, wall = slab_leg(args, rank, world, device, dist, pb, args.steps, args.warmup)
        # the same decomposition for FK_2c (P, N and S planes exchanged), a shorter measurement beside the headline
        pb2 = build_problem("fk2c", res=args.slab_res, on_gpu=True, device=device)
        _, _, info2, _, _, _, _ = slab_leg(args, rank, world, device, dist, pb2, max(1, args.steps // 2), 1, with_e2e=False)
        del pb2
        ens = ensemble_leg(args, rank, world, device, dist) if args.patients > 0 else None
        if rank == 0:
            bpu = ALGO_BYTES[("dti", args.precision)]
            per_rank_vox = pb["voxels"] / world
            kernel_us = 1e3 * ms_step / args.slab_steps
            achieved = per_rank_vox * bpu / (kernel_us * 1e-6) / 1e9
            line = {
                "metric": "voxel-updates/s", "value": value, "unit": "GLUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic",
                "config": workload_config(pb, args.precision, world, dict(steps=args.slab_steps, driver=args.slab_driver, halo=args.slab_halo)),
                "e2e": e2e, "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                             "peak_source": peak_src, "algorithmic_bytes_per_voxel_update": bpu, "kernel_us": kernel_us,
                             "kernel": "dti_step_sp (sparse launch over the rank's owned planes; dti_step_v4 where a slab has no dead space) + in-kernel halo push"
                                       if args.slab_driver == "fused" else "dti_step",
                             "note": "per GPU: one rank's owned planes x algorithmic bytes / the step's share of the decomposed loop "
                                     "(exchange and synchronisation included)"},
                "clocks": clocks, "wall_s_timed_region": wall, "slab": info, "slab_fk2c": info2, "ensemble": ens,
                "note": "N = 1 of this bench is BASELINE.json configs[1] (FK_2c, one GPU); N > 1 is configs[4] (this line) + configs[3] "
                        "(`ensemble`): compare N > 1 values with slab.one_gpu_glups, not with the N = 1 line",
            }
            emit(line)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- N = 1: device-resident arm ------------------------------------------------------------------------
    pb = build_problem(args.model)
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
        launches = int(sim.sync().kernel_launches - launches0) * args.steps // (args.warmup + args.steps)
        if wall < 0.6:                          # too short for a 200 ms sampler: keep the GPU under the same
            t1 = time.perf_counter()            # load (untimed) until a few samples exist
            while time.perf_counter() - t1 < 0.8:
                one_step()
    st = sim.sync()
    final_volume = st.last_volume
    tuning = sim.tuning()               # (after the loop: FK_DTI fp32 picks its kernel family with the sparse mask, at the first run)
    sparse = sim.sparse_info()          # sparse launches: tiles outside the tissue are never loaded / stored (include/tgtk_b200.h)
    # the same loop with the exact per-warp shortcuts for tumour-free / background warps switched off (reported beside the headline)
    dense = None
    if pb["model"] in ("fk2c", "fk"):
        os.environ["TGTK_CLOSED_SKIP"] = "0"
        os.environ["TGTK_SPARSE"] = "0"
        try:
            sim_d, _ = make_sim(pb, args.precision, device)
        finally:
            os.environ.pop("TGTK_CLOSED_SKIP", None)
            os.environ.pop("TGTK_SPARSE", None)
        ms_d = []
        for _ in range(3):
            sim_d.reset()
            flush.zero_()
            torch.cuda.synchronize()
            sim_d.run(pb["nsteps"])
            ms_d.append(sim_d.sync().loop_ms)
        sim_d.close()
        dense = min(ms_d[1:])

    # ---- end-to-end arm: host buffers through the C ABI ------------------------------------
    names = (["wm", "gm"] if pb["model"] != "dti" else []) + ([k for k in pb["fields"] if k != "rgb"])
    host = {k: (pb[k] if k in ("wm", "gm") else pb["fields"][k]) for k in names}
    if pb["model"] == "dti":
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

    updates = float(pb["voxels"]) * pb["nsteps"] * args.steps
    value = updates / (dev_ms * 1e-3) / 1e9
    e2e_value = updates / e2e_wall / 1e9
    bpu = ALGO_BYTES[(pb["model"], args.precision)]
    launch_ms = dev_ms / (pb["nsteps"] * args.steps)          # average launch duration
    # bytes a launch has to move: with sparse launches only the voxels of live tiles are read and written
    # `achieved` is the contract's figure: SURVEY.md 8(d)'s bytes per voxel update x the voxels one launch updates / its duration.
    # With sparse launches the voxels of dead tiles are updated without being touched, so the bytes actually moved are fewer:
    # reported beside it (roofline.sparse), together with the ncu-measured DRAM traffic (roofline.traffic)
    achieved = pb["voxels"] * bpu / (launch_ms * 1e-3) / 1e9
    moved = (sparse["live_voxels"] if sparse["items"] > 0 else pb["voxels"]) * bpu / (launch_ms * 1e-3) / 1e9
    line = {
        "metric": "voxel-updates/s", "value": value, "unit": "GLUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"fp64": "f64", "fp32": "f32", "mixed": "f64 (P, N) + f32 (S, coefficients)"}[args.precision],
        "data": "synthetic", "config": workload_config(pb, args.precision, world),
        "e2e": {"value": e2e_value, "unit": "GLUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_wall / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(pb["model"], args.precision + ("_sparse" if sparse["items"] > 0 else "")), "peak_source": peak_src,
                     "sparse": None if sparse["items"] == 0 else {
                         "what": "2 x 32 voxel tiles that lie outside the tissue and hold no tumour are fixed points of the update: never "
                                 "loaded, computed or stored; the grid is a work list over the live tiles (results bitwise unchanged). "
                                 "`achieved` / `frac` count the algorithmic bytes of every voxel the launch updates; moved_gbs / moved_frac "
                                 "count the live tiles only -- the bytes the launch really has to move (the kernel is then bound by "
                                 "instruction issue and latency, not by HBM: DESIGN.md 4)",
                         "work_items": sparse["items"], "live_fraction": sparse["live_voxels"] / sparse["voxels"],
                         "moved_gbs": moved, "moved_frac": moved / peak},
                     "algorithmic_bytes_per_voxel_update": bpu, "kernel_us": 1e3 * launch_ms,
                     "kernel": {"fk2c": "fk2c_step", "fk": "fk_step", "dti": "dti_step"}[pb["model"]]
                               + ("_sp (sparse launch of the two-voxel kernel)" if sparse["items"] > 0 else "_v%d" % tuning["variant"]),
                     "launch": tuning,
                     "shortcuts": None if dense is None else {
                         "what": "the dense launches with every voxel loaded, computed and stored (TGTK_SPARSE=0 TGTK_CLOSED_SKIP=0): no "
                                 "sparse work list, no per-warp shortcut for tumour-free / background warps",
                         "kernel_us_without": 1e3 * dense / pb["nsteps"],
                         "frac_without": pb["voxels"] * bpu / (dense / pb["nsteps"] * 1e-3) / 1e9 / peak}},
        "clocks": clocks.summary(),
        "wall_s_timed_region": wall, "final_volume": final_volume, "e2e_final_volume": e2e_vol,
    }
    if not args.no_cpu_baseline:
        # the oracle steps the same box from the same initial state: all nsteps when they fit the time budget
        g, threads, n, sec, stepper = cpu_sample(pb, seconds_target=40.0)
        line["cpu_baseline"] = {"value": g, "unit": "GLUPS", "cores": threads, "kind": "port",
                                "sample": f"{n} of {pb['nsteps']} time steps of the same {pb['shape']} box in {sec:.1f} s "
                                          f"(oracle C restatement, OpenMP on {threads} threads)"}
        line["parity"] = parity_vs_oracle(pb, sim, out_fields, stepper, n)
        try:
            line["cpu_baseline"]["numpy"] = numpy_reference_sample(pb, steps=3)
        except Exception as e:
            line["cpu_baseline"]["numpy"] = {"unavailable": f"{type(e).__name__}: {e}"}
    sim.close()
    del flush
    if args.patients > 0 and args.model == "fk2c":
        line["ensemble"] = ensemble_leg(args, rank, world, device, None)
    emit(line)


if __name__ == "__main__":
    main()
