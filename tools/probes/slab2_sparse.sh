T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577"
pick() { python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in sys.argv[1:]})" "$@"; }
timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slab_fused.py -x -q 2>&1 | tail -3
for m in dti fk2c fk; do for sp in 1 0; do echo "== slab $m fused check SPARSE=$sp"; TGTK_SPARSE=$sp timeout 300 $T tools/run_slab.py --model $m --steps 70 --driver fused --check 2>&1 | pick n_gpus max_abs_diff_vs_undecomposed volume_rel_diff loop_ms_max_over_ranks glups_loop; done; done
for m in dti fk2c; do for sp in 1 0; do echo "== slab $m fused 400 steps SPARSE=$sp"; TGTK_SPARSE=$sp timeout 300 $T tools/run_slab.py --model $m --steps 400 --driver fused 2>&1 | pick n_gpus loop_ms_max_over_ranks glups_loop; done; done
echo "== stop criterion"; timeout 300 $T tools/run_slab.py --model dti --steps 200 --driver fused --stop-fraction 0.5 --check 2>&1 | tail -2 | cut -c1-600
