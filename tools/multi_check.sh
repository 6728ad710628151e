#!/bin/bash
# 2-GPU sanity + slab numbers: multi-rank tests, FK_DTI slab with and without two-step launches, fp32 padded pitch
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
pick() { python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in sys.argv[1:]})" "$@"; }
timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -k "not 4" 2>&1 | tail -3
for tb in 0 1; do echo "== slab dti fp64 p2p TB=$tb"; TGTK_TB=$tb timeout 300 $T tools/run_slab.py --model dti --steps 400 --halo 4 --driver p2p --check 2>&1 | pick n_gpus loop_ms_max_over_ranks glups_loop max_abs_diff_vs_undecomposed; done
echo "== slab dti fp64 native auto"; timeout 300 $T tools/run_slab.py --model dti --steps 400 --halo 4 --driver native 2>&1 | pick n_gpus loop_ms_max_over_ranks glups_loop
echo "== slab dti fp32 (odd Z, padded pitch) check"; timeout 300 $T tools/run_slab.py --model dti --precision fp32 --steps 13 --halo 3 --factor 0.5 --check 2>&1 | pick n_gpus max_abs_diff_vs_undecomposed volume_rel_diff
echo "== slab fk2c fp32 check"; timeout 300 $T tools/run_slab.py --model fk2c --precision fp32 --steps 13 --halo 3 --factor 0.5 --check 2>&1 | pick n_gpus max_abs_diff_vs_undecomposed volume_rel_diff
echo "== slab fk2c fp64 p2p"; timeout 300 $T tools/run_slab.py --model fk2c --steps 400 --halo 4 --driver p2p 2>&1 | pick n_gpus loop_ms_max_over_ranks glups_loop
echo "== bench --gpus $N"; timeout 300 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | pick value n_gpus ms_per_step
