#!/bin/bash
# GPU session 8: multi-CTA scan — whole suite, configs 2/1/5 timing, step trace of config 2
set -u
OUT=gpurun_out/r2s8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-24s %.4g moves/s, %.4f ms/DMC step, frac %.3f (launch %.4f ms) whole %.3f" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["whole_step_frac"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6
for cfg in 2 1 5 4; do
  timeout 300 python bench.py --config $cfg --steps 4 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg$cfg.json 2> $OUT/cfg$cfg.err
  line $OUT/cfg$cfg.json "cfg$cfg"
done
QMC4_TRACE=12 timeout 300 python bench.py --config 2 --steps 1 --equil 20 --dmc-steps 12 --no-cpu-baseline --no-secondary 2>&1 | grep "qmc4 trace" | tail -13
QMC4_TRACE=12 timeout 300 python bench.py --config 1 --steps 1 --equil 20 --dmc-steps 12 --no-cpu-baseline --no-secondary 2>&1 | grep "qmc4 trace" | tail -13
