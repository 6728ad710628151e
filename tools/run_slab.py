"""Slab-decomposed FK_DTI / FK / FK_2c time loop on the resampled atlas (BASELINE.json configs[4]: --factor 2.0).
Launch with torchrun (one rank per GPU) or plain python (one GPU, ring on itself).

    python -m torch.distributed.run --nproc-per-node 8 tools/run_slab.py --model dti --steps 400
    ... --check            compare every field with the undecomposed loop on rank 0 (small --steps)
    ... --stop-fraction f  stop criterion: stopping_volume = volume reached after f * steps steps (fused driver)
    ... --solve            instead: FK_DTI_Solver({... 'devices': 'slab'}).solve() against the one-GPU solve()
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dti", choices=["dti", "fk", "fk2c"])
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--halo", type=int, default=0, help="0: 1 for the fused driver, 4 otherwise")
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--factor", type=float, default=2.0)
    ap.add_argument("--driver", default="fused", choices=["fused", "native", "torch", "p2p"])
    ap.add_argument("--check", action="store_true", help="compare with the undecomposed loop on rank 0 (small --steps)")
    ap.add_argument("--stop-fraction", type=float, default=0.0)
    ap.add_argument("--solve", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    halo = args.halo or (1 if args.driver == "fused" else 4)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    import bench
    from tumorgrowthtoolkit_b200 import _loop, slab

    if args.solve:
        return solve_mode(args, rank, world, local, dist)

    pb = bench.build_problem(args.model, res=args.factor, on_gpu=True, device=local)
    q, h, f = pb["q"], pb["h"], pb["fields"]
    extra = {}
    if args.model == "dti":
        call = lambda mod, n, **kw: mod.dti_loop(f["u"], f["rgb"], q["Dw"], q["rho"], pb["dt"], *h, n, precision=args.precision, device=local, **kw)
    elif args.model == "fk":
        call = lambda mod, n, **kw: mod.fk_loop(f["u"], pb["wm"], pb["gm"], q["th_matter"], q["Dw"], q["ratio"], q["rho"], pb["dt"], *h, n,
                                                precision=args.precision, device=local, **kw)
    else:
        call = lambda mod, n, **kw: mod.fk2c_loop(f["P"], f["N"], f["S"], pb["wm"], pb["gm"], q["th_matter"], q["th_necro"], q["Dw"], q["ratio"],
                                                  q["D_s"], q["rho"], q["lambda_np"], q["sigma_np"], q["lambda_s"], pb["dt"], *h, n,
                                                  precision=args.precision, device=local, **kw)
    if args.stop_fraction > 0:
        probe = call(slab, max(1, int(args.stop_fraction * args.steps)), halo=halo, driver=args.driver)
        # just below the volume the probe reached: the decomposed and the undecomposed run sum their partial volumes in
        # different orders (last-bit differences), a threshold EXACTLY on a reached volume would let them stop one step apart
        extra["stopping_volume"] = probe["volume"] * (1.0 - 1e-9)

    def timed():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        res = call(slab, args.steps, halo=halo, driver=args.driver, **extra)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        return res, time.perf_counter() - t0

    for _ in range(max(1, args.reps) - 1):
        timed()                   # warm-up (module load, NCCL channels, memory pools)
    res, wall = timed()
    loop_ms = res["loop_ms"]
    if dist is not None:
        t = torch.tensor([wall, loop_ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, loop_ms = float(t[0]), float(t[1])
    out = {"workload": f"{args.model} slab-decomposed on the {args.factor}x resampled atlas", "box": list(pb["shape"]),
           "voxels": pb["voxels"], "steps": args.steps, "steps_run": res["steps_run"], "stop": res["stop"], "halo": halo, "driver": args.driver,
           "n_gpus": world, "precision": args.precision, "wall_s_incl_setup_and_gather": wall, "loop_ms_max_over_ranks": loop_ms,
           "glups_loop": pb["voxels"] * res["steps_run"] / (loop_ms * 1e-3) / 1e9 if loop_ms > 0 else None,
           "us_per_step": 1e3 * loop_ms / max(1, res["steps_run"]), "final_volume": res["volume"]}
    if args.check and rank == 0:
        os.environ["TGTK_VARIANT"] = "4"      # the two-voxel kernels the fused driver steps (fp32 FK_DTI defaults to z pairs)
        ref = call(_loop, args.steps, **extra)
        os.environ.pop("TGTK_VARIANT", None)
        names = "PNS" if args.model == "fk2c" else ("u",)
        a = res["state"] if args.model == "fk2c" else {"u": res["state"]}
        b = ref["state"] if args.model == "fk2c" else {"u": ref["state"]}
        out["max_abs_diff_vs_undecomposed"] = max(float(np.max(np.abs(a[k] - b[k]))) for k in names)
        out["volume_rel_diff"] = abs(res["volume"] - ref["volume"]) / abs(ref["volume"])
        out["same_stop"] = bool(res["steps_run"] == ref["steps_run"] and res["stop"] == ref["stop"])
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def solve_mode(args, rank, world, local, dist):
    """FK_DTI notebook configuration through the solver API on `world` GPUs against one GPU: identical result dict."""
    import bench
    from tumorgrowthtoolkit_b200 import synthetic
    from tumorgrowthtoolkit_b200.FK_DTI import FK_DTI_Solver, tools
    import contextlib, io
    seg = bench.atlas_labels()
    with contextlib.redirect_stdout(io.StringIO()):
        tensors = tools.get_tensor_from_lower6(synthetic.dti_lower6_field(seg))
        pct = synthetic.seed_inside(seg == 3, bench.SEED_PCT_DTI)
        p = dict(Dw=1.0, rho=0.2, init_scale=0.1 if args.factor <= 1 else 1.0, diffusionEllipsoidScaling=1, diffusionTensors=tensors,
                 NxT1_pct=pct[0], NyT1_pct=pct[1], NzT1_pct=pct[2], resolution_factor=args.factor, stopping_volume=15000.0,
                 stopping_time=1000, precision=args.precision, device=local)
        t0 = time.perf_counter()
        a = FK_DTI_Solver(dict(p, devices="slab")).solve()
        wall = time.perf_counter() - t0
    out = {"workload": f"FK_DTI_Solver.solve(), notebook parameters, resolution_factor {args.factor}, devices='slab'", "n_gpus": world,
           "success": bool(a["success"]), "error": a.get("error"), "final_time": a.get("final_time"),
           "final_volume": float(a.get("final_volume", float("nan"))), "stopping_criteria": a.get("stopping_criteria"), "wall_s": wall}
    if rank == 0 and a["success"]:
        with contextlib.redirect_stdout(io.StringIO()):
            b = FK_DTI_Solver(dict(p)).solve()
        out["one_gpu_final_time"] = b["final_time"]
        out["one_gpu_final_volume"] = float(b["final_volume"])
        out["same_keys"] = sorted(a.keys()) == sorted(b.keys())
        out["final_state_max_abs_diff"] = float(np.max(np.abs(a["final_state"] - b["final_state"])))
        out["same_stop"] = bool(a["final_time"] == b["final_time"] and a["stopping_criteria"] == b["stopping_criteria"])
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
