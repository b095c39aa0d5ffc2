#!/bin/bash
# Round 2, GPU session 2 (1 GPU): new production pipeline — GPU suite, default bench line (config 5 + secondary shapes)
set -u
OUT=gpurun_out/r2s2
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"
tail -5 $OUT/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s2/bench_default.json").read().strip().splitlines()[-1])
print("cfg5: %.4g moves/s, %.3f ms/DMC step, frac %.3f (launch %.3f ms) whole-step %.3f, e2e %.4g, launches %d, clocks %s" % (
    d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["whole_step_frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
print("cpu:", d.get("cpu_baseline"))
for k, v in d.get("secondary", {}).items():
    if "error" in v: print(k, v); continue
    print("%s: %.4g moves/s, %.4f ms/DMC step, frac %.3f (launch %.4f ms), cpu %s" % (k, v["value"], v["ms_per_dmc_step"], v["roofline"]["frac"] or -1, v["roofline"]["avg_launch_ms"], v.get("cpu_baseline", {}).get("value")))
PY
