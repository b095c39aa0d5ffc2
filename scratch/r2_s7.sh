#!/bin/bash
# GPU session 7: tuned triplet loop, then the default bench line (config 5 + secondary shapes + CPU baselines)
set -u
OUT=gpurun_out/r2s7
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_new_paths.py -m gpu -q -p no:cacheprovider -k "three_body or trimer or triplet or mixture3 or impurity" 2>&1 | tail -3
timeout 300 python bench.py --config 4 --steps 4 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg4.json 2> $OUT/cfg4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s7/cfg4.json").read().strip().splitlines()[-1])
print("cfg4: %.4g moves/s, %.4f ms/DMC step, frac %.3f (launch %.4f ms)" % (d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"]))
PY
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -3 $OUT/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s7/bench_default.json").read().strip().splitlines()[-1])
print("cfg5: %.4g moves/s, %.3f ms/DMC step, frac %.3f (launch %.3f ms) whole-step %.3f, e2e %.4g, launches %d, clocks %s" % (
    d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["whole_step_frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
print("cpu:", d.get("cpu_baseline"))
for k, v in d.get("secondary", {}).items():
    if "error" in v: print(k, v); continue
    print("%s: %.4g moves/s, %.4f ms/DMC step, frac %.3f (launch %.4f ms), cpu %s" % (k, v["value"], v["ms_per_dmc_step"], v["roofline"]["frac"] or -1, v["roofline"]["avg_launch_ms"], v.get("cpu_baseline", {}).get("value")))
PY
