#!/bin/bash
# GPU session 9: small-system kernel + graphs: whole suite, config 1 A/B, default bench
set -u
OUT=gpurun_out/r2s9
mkdir -p $OUT
export PYTHONUNBUFFERED=1
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-30s %.4g moves/s, %.4f ms/DMC step, frac %.4f (launch %.4f ms), launches %d" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["gpu_launches"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6
for v in "A=1" "QMC4_GRAPHS=0" "QMC4_SMALL=0" "QMC4_SMALL=0 QMC4_GRAPHS=0"; do
  env $v timeout 300 python bench.py --config 1 --steps 4 --equil 100 --no-cpu-baseline --no-secondary > "$OUT/cfg1_$v.json" 2> "$OUT/cfg1_$v.err"
  line "$OUT/cfg1_$v.json" "cfg1 $v"
done
env timeout 300 python bench.py --config 1 --walkers 200000 --steps 4 --equil 50 --dmc-steps 20 --no-cpu-baseline --no-secondary > $OUT/cfg1_200k.json 2> $OUT/cfg1_200k.err
line $OUT/cfg1_200k.json "cfg1 200k walkers packed"
env QMC4_SMALL=0 timeout 300 python bench.py --config 1 --walkers 200000 --steps 4 --equil 50 --dmc-steps 20 --no-cpu-baseline --no-secondary > $OUT/cfg1_200k_gen.json 2> $OUT/cfg1_200k_gen.err
line $OUT/cfg1_200k_gen.json "cfg1 200k walkers 1 warp/walker"
timeout 300 python bench.py --config 2 --steps 4 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg2.json 2> $OUT/cfg2.err
line $OUT/cfg2.json "cfg2"
