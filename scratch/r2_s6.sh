#!/bin/bash
# GPU session 6: cyclic triplet pass + phasor S(q): whole GPU suite, configs 4 and 3, ncu of config 4
set -u
OUT=gpurun_out/r2s6
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
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
for cfg in 4 3; do
  timeout 300 python bench.py --config $cfg --steps 4 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg$cfg.json 2> $OUT/cfg$cfg.err
  line $OUT/cfg$cfg.json "cfg$cfg"
done
QMC4_TRIPLET_CYCLIC=0 timeout 300 python bench.py --config 4 --steps 4 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg4_old.json 2> $OUT/cfg4_old.err
line $OUT/cfg4_old.json "cfg4 (3x visits)"
bash scratch/ncu_capture.sh r02d_move_cfg4 move_kernel 4 -- python bench.py --config 4 --steps 1 --warmup 3 --equil 2 --dmc-steps 2 --no-cpu-baseline --no-secondary > /dev/null
bash scratch/ncu_capture.sh r02d_est_cfg4 estimator_kernel 4 -- python bench.py --config 4 --steps 1 --warmup 3 --equil 2 --dmc-steps 2 --no-cpu-baseline --no-secondary > /dev/null
bash scratch/ncu_capture.sh r02d_est_cfg3 estimator_kernel 4 -- python bench.py --config 3 --steps 1 --warmup 3 --equil 2 --dmc-steps 2 --no-cpu-baseline --no-secondary > /dev/null
