#!/bin/bash
# Round-end measurement pass on one B200: tests, all six bench lines, the reference arm, launch list and one
# full ncu capture of the headline kernel, step times on the 2x up-sampled box, whole-solve wall times.
set -u
O=gpurun_out/final
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_fk2c_fp64.json 2> $O/bench_fk2c_fp64.err; tail -c 600 $O/bench_fk2c_fp64.err
for m in fk2c fk dti; do for p in fp64 fp32; do
  [ "$m$p" = "fk2cfp64" ] && continue
  timeout 200 python bench.py --model $m --precision $p --no-cpu-baseline > $O/bench_${m}_${p}.json 2> $O/bench_${m}_${p}.err
done; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline", {})
    print(f.split("/")[-1], round(d.get("value", 0), 2), d.get("unit"), "frac", round(r.get("frac", 0), 3), "us", round(r.get("kernel_us", 0), 2),
          "e2e", round(d.get("e2e", {}).get("value", 0), 2), r.get("launch"), d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_fk2c_fp64.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_list.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:fk2c_step --launch-skip 300 --launch-count 1 -o $O/fk2c_fp64_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
python tools/ncu_summary.py $O/fk2c_fp64_final.ncu-rep > $O/fk2c_fp64_final.ncu_summary.txt 2>/dev/null; head -12 $O/fk2c_fp64_final.ncu_summary.txt
python tools/ncu_by_line.py $O/fk2c_fp64_final.ncu-rep _ZN4tgtk12fk2c_step_spIdLi4ELi0EEEvNS_8StepArgsIT_EENS_10SparseArgsE 25 > $O/fk2c_fp64_final.by_line.txt 2>&1
# sparse against dense launches on the benchmark box, all six model / precision pairs
for m in fk2c fk dti; do for p in fp64 fp32; do
  timeout 200 python tools/sweep_atlas.py --model $m --precision $p TGTK_SPARSE=0,1 2>&1 | tail -3
done; done > $O/sparse_vs_dense.txt; cat $O/sparse_vs_dense.txt
timeout 100 python tools/sparse_trace.py > $O/sparse_trace.txt 2>&1; head -7 $O/sparse_trace.txt
timeout 200 python tools/sweep_env.py --shape 288,354,277 --steps 100 > $O/steps_big.txt 2>&1; cat $O/steps_big.txt
timeout 200 python tools/sweep_env.py > $O/steps_atlas.txt 2>&1; cat $O/steps_atlas.txt
timeout 200 python tools/sweep_env.py --shape 95,117,91 > $O/steps_ensemble.txt 2>&1; cat $O/steps_ensemble.txt
timeout 300 python tools/solve_wall.py > $O/solve_wall.txt 2>&1; tail -8 $O/solve_wall.txt
