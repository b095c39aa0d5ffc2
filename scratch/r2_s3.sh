#!/bin/bash
# Round 2, GPU session 3 (1 GPU): the tournament-schedule pair kernel — parity tests, then A/B at 2 and 3 CTAs/SM
set -u
OUT=gpurun_out/r2s3
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_pipeline2.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s %.4g moves/s, %.4f ms/DMC step, frac %.3f (launch %.4f ms) whole %.3f, kernel %s" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["whole_step_frac"], d["kernel"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
for cfg in 5 3 2; do
  for v in "QMC4_LEAN=0" "QMC4_LEAN_OCC=2" "QMC4_LEAN_OCC=3"; do
    env $v timeout 300 python bench.py --config $cfg --steps 5 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg${cfg}_$v.json 2> $OUT/cfg${cfg}_$v.err
    line $OUT/cfg${cfg}_$v.json "cfg$cfg $v"
    tail -2 $OUT/cfg${cfg}_$v.err
  done
done
QMC4_LEAN_OCC=2 bash scratch/ncu_capture.sh r02b_pair_cfg5_occ2 pair_move_kernel 4 -- python bench.py --config 5 --steps 1 --warmup 3 --equil 2 --dmc-steps 2 --no-cpu-baseline --no-secondary
QMC4_LEAN_OCC=3 bash scratch/ncu_capture.sh r02b_pair_cfg5_occ3 pair_move_kernel 4 -- python bench.py --config 5 --steps 1 --warmup 3 --equil 2 --dmc-steps 2 --no-cpu-baseline --no-secondary
