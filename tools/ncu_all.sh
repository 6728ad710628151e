#!/bin/bash
# One `ncu --set full` capture per model / precision of the default step kernel on the atlas box (run under gpurun,
# one GPU): writes gpurun_out/ncu/<model>_<dtype>.ncu-rep and a text summary next to it (tools/ncu_summary.py).
# usage: tools/ncu_all.sh [shape]      e.g. tools/ncu_all.sh 288,354,277
SHAPE=${1:-145,178,139}
O=gpurun_out/ncu
mkdir -p $O
for m in 2c fk dti dtifull; do for d in f64 f32; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:step --launch-skip 150 --launch-count 1 \
      -o $O/${m}_${d} -f python tools/sweep_env.py --models $m --dtypes $d --steps 100 --shape $SHAPE > $O/${m}_${d}.log 2>&1
  python tools/ncu_summary.py $O/${m}_${d}.ncu-rep > $O/${m}_${d}.summary.txt 2>/dev/null
  head -3 $O/${m}_${d}.summary.txt
done; done
