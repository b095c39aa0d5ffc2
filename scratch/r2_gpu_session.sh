#!/bin/bash
# First GPU call of round 2, in ONE gpurun (1 GPU):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scratch/r2_gpu_session.sh'
# Everything that was written while no B200 time was left in round 1 gets its hardware run here, in the order
# that matters if the call is cut short.  Every step has its own timeout; logs land in gpurun_out/r2/.
set -u
OUT=gpurun_out/r2
mkdir -p $OUT
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a $OUT/session.log; }

step "1. GPU test suite (validated files first, never-on-hardware paths last)"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/session.log
# the new-path file on its own, without -x, so that one hardware-only failure does not hide the others
timeout 900 python -m pytest tests/test_zz_gpu_new_paths.py -m gpu -q -p no:cacheprovider > $OUT/pytest_new_paths.log 2>&1
tail -15 $OUT/pytest_new_paths.log | tee -a $OUT/session.log

step "2. smoke + default bench (config 2, pipeline 1)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log | tee -a $OUT/session.log
timeout 600 python bench.py > $OUT/bench_cfg2_p1.json 2> $OUT/bench_cfg2_p1.err; cut -c1-400 $OUT/bench_cfg2_p1.json | tee -a $OUT/session.log

step "3. experimental pipeline 2: bitwise check, then timing with the estimator kernel at 2/3/4 CTAs per SM"
QMC4_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental_pipeline2.py -m gpu -q -p no:cacheprovider > $OUT/pytest_pipeline2.log 2>&1
tail -3 $OUT/pytest_pipeline2.log | tee -a $OUT/session.log
for occ in 2 3 4; do
  QMC4_PIPELINE=2 QMC4_EST_OCC=$occ timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_cfg2_p2_occ$occ.json 2> $OUT/bench_cfg2_p2_occ$occ.err
  python - <<EOF | tee -a $OUT/session.log
import json
try:
    d = json.loads(open("$OUT/bench_cfg2_p2_occ$occ.json").read().strip().splitlines()[-1])
    print("pipeline 2 occ $occ: %.4g walker-moves/s, %.4f ms/step, e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("pipeline 2 occ $occ: no result (%s)" % e)
EOF
done

step "3b. launch-geometry sweep of the default pipeline at config 2 (warps per walker, resident CTAs per SM)"
for knob in "QMC4_WPW=1" "QMC4_WPW=2" "QMC4_WPW=4" "QMC4_CTAS_PER_SM=1" "QMC4_CTAS_PER_SM=2"; do
  env $knob timeout 300 python bench.py --steps 100 --no-cpu-baseline > $OUT/bench_cfg2_${knob/=/_}.json 2> $OUT/bench_cfg2_${knob/=/_}.err
  python - <<EOF2 | tee -a $OUT/session.log
import json
try:
    d = json.loads(open("$OUT/bench_cfg2_${knob/=/_}.json").read().strip().splitlines()[-1])
    print("$knob: %.4g walker-moves/s, %.4f ms/step, roofline %.3f, kernel %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["kernel"]))
except Exception as e:
    print("$knob: no result (%s)" % e)
EOF2
done

step "3c. new statistical parity test of the pure estimators alone (kept out of -x's way above)"
timeout 600 python -m pytest tests/test_zz_gpu_new_paths.py -m gpu -q -p no:cacheprovider -k "pure_estimators or force_energy or superfluid or dipolar" > $OUT/pytest_r1b.log 2>&1
tail -5 $OUT/pytest_r1b.log | tee -a $OUT/session.log

step "4. the other BASELINE shapes (1 GPU)"
for cfg in 5 3 4 1; do
  timeout 900 python bench.py --config $cfg --steps 20 --no-cpu-baseline > $OUT/bench_cfg${cfg}.json 2> $OUT/bench_cfg${cfg}.err
  cut -c1-300 $OUT/bench_cfg${cfg}.json | tee -a $OUT/session.log
done

step "5. ncu: launch list of the bench command, then one full capture of the dominant kernel"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --equil 5 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 8 -c 1 -o $OUT/propagate_cfg2 -f \
  python bench.py --steps 2 --warmup 3 --equil 5 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 4 -c 1 -o $OUT/propagate_cfg5 -f \
  python bench.py --config 5 --steps 2 --warmup 3 --equil 2 --no-cpu-baseline > $OUT/ncu_full5.log 2>&1
ls -la $OUT | tee -a $OUT/session.log
