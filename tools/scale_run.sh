#!/bin/bash
# usage: tools/scale_run.sh N   -- bench, ensemble and slab runs on N GPUs of this box
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "1" ]; then T="python"; fi
pick() { python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in sys.argv[1:]})" "$@"; }
echo "== bench.py --gpus $N"; timeout 600 $T bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | pick value n_gpus ms_per_step
echo "== ensemble 32 patients per GPU, low-res outputs"; timeout 600 $T tools/run_ensemble.py --patients $((32*N)) --lowres 2>&1 | pick patients n_gpus wall_s patients_per_hour glups
echo "== ensemble 16 patients per GPU, full-resolution outputs"; timeout 600 $T tools/run_ensemble.py --patients $((16*N)) 2>&1 | pick patients n_gpus wall_s patients_per_hour glups
for d in p2p native; do for m in dti fk2c; do echo "== slab $m halo 4 driver $d"; timeout 600 $T tools/run_slab.py --model $m --steps 400 --halo 4 --driver $d 2>&1 | pick n_gpus halo loop_ms_max_over_ranks glups_loop; done; done
