#!/bin/bash
set -u
OUT=gpurun_out/strong_dbg
mkdir -p $OUT
export PYTHONUNBUFFERED=1 QMC4_BENCH_VERBOSE=1
timeout 200 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29741 bench.py --gpus 2 --scaling strong --steps 2 --warmup 1 --equil 60 --no-cpu-baseline --no-secondary --watchdog 150 > $OUT/n2_strong.json 2> $OUT/n2_strong.err
grep "bench rank\|Error\|error" $OUT/n2_strong.err | head -40
timeout 200 python bench.py --walkers 50000 --steps 2 --warmup 1 --equil 60 --no-cpu-baseline --no-secondary > $OUT/n1_50k.json 2> $OUT/n1_50k.err
grep "bench rank\|Error\|error" $OUT/n1_50k.err | head -20
tail -c 400 $OUT/n1_50k.json
unset QMC4_BENCH_VERBOSE
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-30s %.4g moves/s, %.4f ms/DMC step, frac %.4f (launch %.4f ms) regs %s smem %s grid %s" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["kernel"].get("regs"), d["kernel"].get("smem_bytes"), d["kernel"].get("grid")))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
for v in 0 3 0 3; do
  QMC4_LEAN_VARIANT=$v timeout 200 python bench.py --config 5 --steps 5 --equil 50 --no-cpu-baseline --no-secondary > $OUT/cfg5_v$v.json 2> $OUT/cfg5_v$v.err
  line $OUT/cfg5_v$v.json "cfg5 variant $v"
done
for v in 0 3; do
  QMC4_LEAN_VARIANT=$v timeout 200 python bench.py --config 2 --steps 5 --equil 50 --no-cpu-baseline --no-secondary > $OUT/cfg2_v$v.json 2> $OUT/cfg2_v$v.err
  line $OUT/cfg2_v$v.json "cfg2 variant $v"
  QMC4_LEAN_VARIANT=$v timeout 200 python bench.py --config 3 --steps 5 --equil 50 --no-cpu-baseline --no-secondary > $OUT/cfg3_v$v.json 2> $OUT/cfg3_v$v.err
  line $OUT/cfg3_v$v.json "cfg3 variant $v"
done
