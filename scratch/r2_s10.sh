#!/bin/bash
# GPU session 10 (1 GPU): whole suite, estimator-kernel A/B, default bench, launch list, ncu --set full captures, compute-sanitizer
set -u
OUT=gpurun_out/r2s10
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-30s %.4g moves/s, %.4f ms/DMC step, frac %.4f (launch %.4f ms), e2e %.4g, launches %d" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"], d["gpu_launches"]))
    for k in d["roofline"].get("kernels", []): print("      ", k)
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log
stamp suite
for v in "A=1" "QMC4_LEAN_HIST=0"; do
  env $v timeout 300 python bench.py --config 2 --steps 4 --equil 100 --no-cpu-baseline --no-secondary > "$OUT/cfg2_$v.json" 2> "$OUT/cfg2_$v.err"
  line "$OUT/cfg2_$v.json" "cfg2 $v"
done
timeout 300 python bench.py --config 3 --steps 4 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg3.json 2> $OUT/cfg3.err
line $OUT/cfg3.json "cfg3"
stamp ab
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
line $OUT/bench_n1.json "default bench"
stamp bench
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
tail -c 600 $OUT/bench_ref.json
stamp ref
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --equil 10 --no-cpu-baseline --no-secondary > $OUT/launches_bench.log 2>&1
wc -l $OUT/launches_bench.csv
stamp launches
B="--steps 1 --warmup 1 --equil 3 --dmc-steps 2 --no-cpu-baseline --no-secondary"
bash scratch/ncu_capture.sh r02f_pair_cfg5 pair_move_kernel 6 -- python bench.py --config 5 $B | tail -1
bash scratch/ncu_capture.sh r02f_hist_cfg2 pair_hist_kernel 6 -- python bench.py --config 2 $B | tail -1
bash scratch/ncu_capture.sh r02f_pair_cfg2 pair_move_kernel 6 -- python bench.py --config 2 $B | tail -1
bash scratch/ncu_capture.sh r02f_est_cfg3 estimator_kernel 6 -- python bench.py --config 3 $B | tail -1
bash scratch/ncu_capture.sh r02f_small_cfg1 small_move_kernel 6 -- python bench.py --config 1 $B | tail -1
bash scratch/ncu_capture.sh r02f_scan_cfg2 scan_multi_kernel 6 -- python bench.py --config 2 $B | tail -1
stamp ncu
for tool in memcheck racecheck synccheck; do
  fams=""; [ $tool = racecheck ] && fams="small pair ragged hist triplet sq mixture"
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python scratch/sanitize_cases.py $fams > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c ' ok:' $OUT/sanitizer_$tool.log) runs; $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/sanitizer_$tool.log | tail -1)"
done
stamp sanitizer
