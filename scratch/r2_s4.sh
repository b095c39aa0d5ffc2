#!/bin/bash
# GPU session 4: pair-kernel launch-geometry variants x unroll factors on configs 5, 3, 2
set -u
OUT=gpurun_out/r2s4
mkdir -p $OUT
export PYTHONUNBUFFERED=1
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-34s %.4g moves/s, %.4f ms/DMC step, frac %.3f (launch %.4f ms), grid %d regs %d" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["kernel"]["grid"], d["kernel"]["regs"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
timeout 600 python -m pytest tests/test_pipeline2.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for cfg in 5 3 2; do
  for lib in "" scratch/variants/libqmc4b200_u1.so scratch/variants/libqmc4b200_u4.so; do
    for v in 0 1 2; do
      name="cfg${cfg}_v${v}_$(basename ${lib:-u2} .so)"
      env QMC4_LEAN_VARIANT=$v ${lib:+QMC4_B200_LIB=$PWD/$lib} timeout 300 python bench.py --config $cfg --steps 4 --equil 100 --no-cpu-baseline --no-secondary > $OUT/$name.json 2> $OUT/$name.err
      line $OUT/$name.json "$name"
    done
  done
done
