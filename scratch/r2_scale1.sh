#!/bin/bash
# one scaling point: bash scratch/r2_scale1.sh N [extra bench args]  (charged N x)
set -u
N=${1:-8}; shift
OUT=gpurun_out/scale1_n$N
mkdir -p $OUT
export PYTHONUNBUFFERED=1
if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29720"; fi
timeout 240 $RUN bench.py --gpus $N --steps 10 --no-cpu-baseline --no-secondary --watchdog 200 "$@" > $OUT/run.json 2> $OUT/run.err
python - $OUT/run.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("migration") or {}
    ins = m.get("in_step_migration", {})
    print("N=%d %s %.4g moves/s, %.4f ms/DMC step (move kernel %.4f ms, frac %.3f, whole %.3f), e2e %.4g, pop %d, events %s out %s ok %s" % (
        d["n_gpus"], d["config"]["scaling"], d["value"], d["ms_per_dmc_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"],
        d["e2e"]["value"], d["population"], ins.get("events_sum_over_ranks"), ins.get("walkers_out"), m.get("ok")))
    print("   pops", m.get("rank_populations"))
except Exception as e:
    print("no result", e)
PY
grep -h -v "^$\|OMP_NUM\|\*\*\*\*\|destroy_process_group" $OUT/run.err | tail -5
