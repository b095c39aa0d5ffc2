#!/bin/bash
# Round 2, GPU session 1 (1 GPU): /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scratch/r2_s1.sh'
# whole GPU suite without -x, baselines of every BASELINE shape, pipeline-2 A/B, ncu captures, compute-sanitizer.
set -u
OUT=gpurun_out/r2s1
mkdir -p $OUT
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a $OUT/session.log; }
line() { python - "$1" "$2" <<'EOF' | tee -a $OUT/session.log
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%s: %.4g walker-moves/s, %.4f ms/step, roofline %.3f (launch %.4f ms), e2e %.4g, kernel %s" % (
        sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"], d["kernel"]))
except Exception as e:
    print("%s: no result (%s)" % (sys.argv[2], e))
EOF
}

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $OUT/session.log

step "1. whole GPU suite, no -x"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log | tee -a $OUT/session.log
QMC4_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental_pipeline2.py -m gpu -q -p no:cacheprovider > $OUT/pytest_pipeline2.log 2>&1
tail -3 $OUT/pytest_pipeline2.log | tee -a $OUT/session.log

step "2. smoke + config 2 baseline (pipeline 1) and pipeline 2 at 2/3/4 estimator CTAs per SM"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log | tee -a $OUT/session.log
timeout 600 python bench.py --steps 100 --no-cpu-baseline > $OUT/bench_cfg2_p1.json 2> $OUT/bench_cfg2_p1.err; line $OUT/bench_cfg2_p1.json "cfg2 p1"
for occ in 2 3 4; do
  QMC4_PIPELINE=2 QMC4_EST_OCC=$occ timeout 600 python bench.py --steps 100 --no-cpu-baseline > $OUT/bench_cfg2_p2_occ$occ.json 2> $OUT/bench_cfg2_p2_occ$occ.err
  line $OUT/bench_cfg2_p2_occ$occ.json "cfg2 p2 occ$occ"
done

step "3. the other BASELINE shapes (1 GPU)"
for cfg in 5 3 4 1; do
  timeout 900 python bench.py --config $cfg --steps 20 --no-cpu-baseline > $OUT/bench_cfg${cfg}.json 2> $OUT/bench_cfg${cfg}.err
  line $OUT/bench_cfg${cfg}.json "cfg$cfg"
done

step "4. ncu: launch list, then full captures of the dominant kernel (configs 2, 5, 4)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --equil 5 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 8 -c 1 -o $OUT/propagate_cfg2 -f \
  python bench.py --steps 2 --warmup 3 --equil 5 --no-cpu-baseline > $OUT/ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 4 -c 1 -o $OUT/propagate_cfg5 -f \
  python bench.py --config 5 --steps 2 --warmup 3 --equil 2 --no-cpu-baseline > $OUT/ncu_full5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 4 -c 1 -o $OUT/propagate_cfg4 -f \
  python bench.py --config 4 --steps 2 --warmup 3 --equil 2 --no-cpu-baseline > $OUT/ncu_full4.log 2>&1

step "5. compute-sanitizer on the smoke run"
timeout 900 compute-sanitizer --tool memcheck --log-file $OUT/memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck_smoke.out 2>&1
tail -3 $OUT/memcheck_smoke.log | tee -a $OUT/session.log
timeout 900 compute-sanitizer --tool racecheck --log-file $OUT/racecheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck_smoke.out 2>&1
tail -3 $OUT/racecheck_smoke.log | tee -a $OUT/session.log
ls -la $OUT | tee -a $OUT/session.log
