#!/bin/bash
set -u
OUT=gpurun_out/strong2
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 200 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29742 bench.py --gpus 2 --scaling strong --steps 10 --equil 100 --no-cpu-baseline --no-secondary --watchdog 180 > $OUT/cfg5_strong_n2.json 2> $OUT/cfg5_strong_n2.err
python - $OUT/cfg5_strong_n2.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("migration") or {}
    ins = m.get("in_step_migration", {})
    print("cfg5 strong N=%d %.4g moves/s, %.4f ms/DMC step (move %.4f ms, frac %.3f, whole %.3f), e2e %.4g, pop %d, events %s out %s ok %s" % (
        d["n_gpus"], d["value"], d["ms_per_dmc_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"],
        d["e2e"]["value"], d["population"], ins.get("events_sum_over_ranks"), ins.get("walkers_out"), m.get("ok")))
except Exception as e:
    print("no result", e)
PY
grep -h -v "^$\|OMP_NUM\|\*\*\*\*\|destroy_process_group" $OUT/*.err | tail -5
