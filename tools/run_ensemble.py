"""Synthetic-patient ensemble (BASELINE.json configs[3]) on this rank's GPU; under torchrun every
rank takes its shard.  Prints one JSON line from rank 0 with patients/hour of the whole job.

    python tools/run_ensemble.py --patients 64 [--concurrency 4] [--precision fp64] [--lowres]
    python -m torch.distributed.run --nproc-per-node 8 tools/run_ensemble.py --patients 1024
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patients", type=int, default=32)
    ap.add_argument("--concurrency", type=int, default=4)
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--lowres", action="store_true", help="skip the resample of the results to full resolution")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from tumorgrowthtoolkit_b200 import ensemble, synthetic
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sri_tissue_labels.npz")
    wm, gm = synthetic.tissue_from_labels(np.load(path)["seg"]) if os.path.exists(path) else synthetic.brain_phantom()
    params = ensemble.draw_ensemble(args.patients)
    mine = ensemble.shard(params, rank, world)
    runner = ensemble.EnsembleRunner(wm, gm, device=local, precision=args.precision, concurrency=args.concurrency,
                                     full_resolution=not args.lowres)
    runner.run([params[i] for i in mine[:2]], keep_results=False)            # warm-up (module load, pools)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    _, st = runner.run([params[i] for i in mine], keep_results=False)
    wall = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([wall, float(st["voxel_updates"])], dtype=torch.float64, device=f"cuda:{local}")
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        wall, updates = float(tm[0]), float(ts[1])
    else:
        updates = float(st["voxel_updates"])
    if rank == 0:
        print(json.dumps({"workload": "FK_2c synthetic-patient ensemble, run_gen_FK_2c.py parameter law, res 0.65",
                          "patients": args.patients, "n_gpus": world, "concurrency": args.concurrency,
                          "precision": args.precision, "full_resolution_outputs": not args.lowres,
                          "voxels_per_patient": runner.voxels, "wall_s": wall,
                          "patients_per_hour": 3600.0 * args.patients / wall, "glups": updates / wall / 1e9}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
