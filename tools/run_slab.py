"""Slab-decomposed FK_DTI / FK / FK_2c time loop on the 2x up-sampled atlas (BASELINE.json
configs[4]).  Launch with torchrun (one rank per GPU) or plain python (one GPU, ring on itself).

    python -m torch.distributed.run --nproc-per-node 8 tools/run_slab.py --model dti --steps 400 --halo 4
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dti", choices=["dti", "fk", "fk2c"])
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--halo", type=int, default=4)
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--factor", type=float, default=2.0)
    ap.add_argument("--driver", default="p2p", choices=["native", "torch", "p2p"])
    ap.add_argument("--check", action="store_true", help="compare with the undecomposed loop on rank 0 (small --steps)")
    args = ap.parse_args()
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
    from tumorgrowthtoolkit_b200 import _host, _loop, slab, synthetic
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "tests", "golden", "sri_tissue_labels.npz")
    wm, gm = synthetic.tissue_from_labels(np.load(path)["seg"]) if os.path.exists(path) else synthetic.brain_phantom()
    f = args.factor
    wml, gml = _host.resample_linear(wm, f, device=local), _host.resample_linear(gm, f, device=local)
    grid = _host.LowResGrid(wml.shape, wm.shape, f, 1.0, 1.0, 1.0, (0.3, 0.7, 0.5))
    box = _host.bounding_box(gml + wml >= 0.5, 2, wml.shape)
    cw, cg = (np.ascontiguousarray(_host.crop(a, box)) for a in (wml, gml))
    u0 = np.ascontiguousarray(_host.crop(_host.seed_gaussian(grid, 1.0), box))
    h = grid.h
    if args.model == "dti":
        rgb = np.stack([cw + 0.5 * cg, 0.7 * cw + 0.5 * cg, 0.4 * cw + 0.5 * cg], axis=-1)
        rgb /= rgb[rgb.max(axis=-1) > 0].mean()
        nt = max(100 * 1.0 * rgb.max() / min(h) ** 2 * 8 + 100, 100 * 0.2 * 1.1)
        call = lambda mod, **kw: mod.dti_loop(u0, rgb, 1.0, 0.2, 100 / nt, *h, args.steps, precision=args.precision, device=local, **kw)
    elif args.model == "fk":
        nt = max(100 * 1.0 / min(h) ** 2 * 8 + 100, 100 * 0.1 * 1.1)
        call = lambda mod, **kw: mod.fk_loop(u0, cw, cg, 0.1, 1.0, 10.0, 0.1, 100 / nt, *h, args.steps, precision=args.precision, device=local, **kw)
    else:
        nt = 100 * 1.3 / min(h) ** 2 * 8 + 300
        S0 = np.where(cw + cg >= 0.1, 1.0, 0.0)
        call = lambda mod, **kw: mod.fk2c_loop(u0, np.zeros_like(u0), S0, cw, cg, 0.1, 0.9, 0.9, 100.0, 1.3, 0.14, 0.35, 0.5, 0.05,
                                                100 / nt, *h, args.steps, precision=args.precision, device=local, **kw)

    def timed():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        res = call(slab, halo=args.halo, driver=args.driver)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        return res, time.perf_counter() - t0

    timed()                       # warm-up (module load, NCCL channels, memory pools)
    res, wall = timed()
    loop_ms = res["loop_ms"]
    if dist is not None:
        t = torch.tensor([wall, loop_ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, loop_ms = float(t[0]), float(t[1])
    out = {"workload": f"{args.model} slab-decomposed on the {f}x resampled atlas", "box": list(u0.shape),
           "voxels": int(u0.size), "steps": args.steps, "halo": args.halo, "driver": args.driver, "n_gpus": world, "precision": args.precision,
           "wall_s_incl_setup_and_gather": wall, "loop_ms_max_over_ranks": loop_ms,
           "glups_loop": u0.size * args.steps / (loop_ms * 1e-3) / 1e9 if loop_ms > 0 else None, "final_volume": res["volume"]}
    if args.check and rank == 0:
        ref = call(_loop)
        a, b = (res["state"], ref["state"]) if args.model != "fk2c" else (res["state"]["P"], ref["state"]["P"])
        out["max_abs_diff_vs_undecomposed"] = float(np.max(np.abs(a - b)))
        out["volume_rel_diff"] = abs(res["volume"] - ref["volume"]) / abs(ref["volume"])
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
