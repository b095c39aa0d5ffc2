#!/bin/bash
# scaling session: bash scratch/r2_scale.sh N   -> config 5 weak + strong (+ config 2 weak) at N GPUs (charged N x)
set -u
N=${1:-8}
OUT=gpurun_out/scale_n$N
mkdir -p $OUT
export PYTHONUNBUFFERED=1
if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29720"; fi
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("migration") or {}
    ins = m.get("in_step_migration", {})
    print("%-22s N=%d %.4g moves/s, %.4f ms/DMC step (move kernel %.4f ms, frac %.3f, whole %.3f), e2e %.4g, pop %d, events %s out %s ok %s" % (
        sys.argv[2], d["n_gpus"], d["value"], d["ms_per_dmc_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"],
        d["e2e"]["value"], d["population"], ins.get("events_sum_over_ranks"), ins.get("walkers_out"), m.get("ok")))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
timeout 240 $RUN bench.py --gpus $N --steps 10 --no-cpu-baseline --no-secondary --watchdog 200 > $OUT/cfg5_weak.json 2> $OUT/cfg5_weak.err; show $OUT/cfg5_weak.json "cfg5 weak"
timeout 240 $RUN bench.py --gpus $N --scaling strong --steps 10 --no-cpu-baseline --no-secondary --watchdog 200 > $OUT/cfg5_strong.json 2> $OUT/cfg5_strong.err; show $OUT/cfg5_strong.json "cfg5 strong"
timeout 240 $RUN bench.py --gpus $N --config 2 --steps 10 --no-cpu-baseline --no-secondary --watchdog 200 > $OUT/cfg2_weak.json 2> $OUT/cfg2_weak.err; show $OUT/cfg2_weak.json "cfg2 weak"
grep -h -v "^$\|OMP_NUM\|\*\*\*\*\|destroy_process_group" $OUT/*.err | tail -5
