#!/bin/bash
# GPU session 5: fixed-point pair kernel — parity suite, then variants on configs 5, 3, 2 + one ncu capture
set -u
OUT=gpurun_out/r2s5
mkdir -p $OUT
export PYTHONUNBUFFERED=1
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-24s %.4g moves/s, %.4f ms/DMC step, frac %.3f (launch %.4f ms) whole %.3f, grid %d regs %d" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["whole_step_frac"], d["kernel"]["grid"], d["kernel"]["regs"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
timeout 900 python -m pytest tests/test_pipeline2.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
for cfg in 5 3 2; do
  for v in 0 1 2; do
    name="cfg${cfg}_v${v}"
    env QMC4_LEAN_VARIANT=$v timeout 300 python bench.py --config $cfg --steps 4 --equil 100 --no-cpu-baseline --no-secondary > $OUT/$name.json 2> $OUT/$name.err
    line $OUT/$name.json "$name"
  done
done
bash scratch/ncu_capture.sh r02c_pair_cfg5 pair_move_kernel 4 -- python bench.py --config 5 --steps 1 --warmup 3 --equil 2 --dmc-steps 2 --no-cpu-baseline --no-secondary > /dev/null
bash scratch/ncu_capture.sh r02c_pair_cfg2 pair_move_kernel 8 -- python bench.py --config 2 --steps 1 --warmup 3 --equil 5 --dmc-steps 4 --no-cpu-baseline --no-secondary > /dev/null
