#!/bin/bash
# round-2 A/B experiments (one gpurun call): 5 blocks/SM with late loads, shuffle-based fp32 z-pair kernels, planes per segment
python tools/sweep_env.py --models 2c --dtypes f64,f32 TGTK_MINB=0,5,14
python tools/sweep_env.py --models 2c --dtypes f64 --shape 288,354,277 --steps 200 TGTK_MINB=0,5,14
python tools/sweep_env.py --models 2c,fk --dtypes f32 TGTK_VARIANT=4,5,6
python tools/sweep_env.py --models 2c,fk --dtypes f32 --shape 288,354,277 --steps 200 TGTK_VARIANT=4,5,6
for lx in 4 5 6 9 12 18; do echo "LX=$lx $(TGTK_LX=$lx python tools/slab_probe.py --planes 36 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['plain_us_per_step'],2), round(d['slab_self_ring_us_per_step'],2), d['slab_tuning']['grid'])")"; done
for lx in 10 15 29 37 49; do echo "atlas LX=$lx $(TGTK_LX=$lx python tools/sweep_env.py --models 2c,dti,fk --dtypes f64 TGTK_VARIANT=4 | tail -1)"; done
timeout 300 python -m pytest tests/test_gpu_v5.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -q -x -m gpu 2>&1 | tail -3
