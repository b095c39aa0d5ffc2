#!/bin/bash
# GPU session 11 (1 GPU) after the scan fix: whole suite, default bench, strong-scaling N=1 point (100 000 walkers), config 1 at 200 k walkers A/B
set -u
OUT=gpurun_out/r2s11
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-30s %.4g moves/s, %.4f ms/DMC step, frac %.4f (launch %.4f ms), e2e %.4g, pop %d, launches %d" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"], d["population"], d["gpu_launches"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log
stamp suite
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
line $OUT/bench_n1.json "default bench"
stamp bench
timeout 300 python bench.py --scaling strong --steps 10 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg5_strong_n1.json 2> $OUT/cfg5_strong_n1.err
line $OUT/cfg5_strong_n1.json "cfg5 strong N=1 (100k walkers)"
stamp strong
timeout 300 python bench.py --config 1 --walkers 200000 --steps 4 --equil 50 --dmc-steps 20 --no-cpu-baseline --no-secondary > $OUT/cfg1_200k.json 2> $OUT/cfg1_200k.err
line $OUT/cfg1_200k.json "cfg1 200k walkers packed"
QMC4_SMALL=0 timeout 300 python bench.py --config 1 --walkers 200000 --steps 4 --equil 50 --dmc-steps 20 --no-cpu-baseline --no-secondary > $OUT/cfg1_200k_gen.json 2> $OUT/cfg1_200k_gen.err
line $OUT/cfg1_200k_gen.json "cfg1 200k walkers 1 warp/walker"
timeout 300 python bench.py --config 2 --walkers 131072 --steps 4 --equil 50 --dmc-steps 20 --no-cpu-baseline --no-secondary > $OUT/cfg2_128k.json 2> $OUT/cfg2_128k.err
line $OUT/cfg2_128k.json "cfg2 131072 walkers"
stamp big
