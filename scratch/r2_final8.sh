#!/bin/bash
# final 8-GPU check (charged 8 x): config 2 weak with the current estimator kernels, then the driver's own command line for config 5
set -u
OUT=gpurun_out/final8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29751"
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("migration") or {}
    ins = m.get("in_step_migration", {})
    print("%-14s N=%d %.4g moves/s, %.4f ms/DMC step (move %.4f ms, frac %.3f, whole %.3f), e2e %.4g, pop %d, events %s out %s ok %s" % (
        sys.argv[2], d["n_gpus"], d["value"], d["ms_per_dmc_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"],
        d["e2e"]["value"], d["population"], ins.get("events_sum_over_ranks"), ins.get("walkers_out"), m.get("ok")))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
timeout 200 $RUN bench.py --gpus 8 --config 2 --steps 10 --equil 100 --no-cpu-baseline --no-secondary --watchdog 180 > $OUT/cfg2_weak_n8.json 2> $OUT/cfg2_weak_n8.err
show $OUT/cfg2_weak_n8.json "cfg2 weak"
timeout 200 python bench.py --config 2 --steps 10 --equil 100 --no-cpu-baseline --no-secondary > $OUT/cfg2_weak_n1.json 2> $OUT/cfg2_weak_n1.err
show $OUT/cfg2_weak_n1.json "cfg2 weak"
timeout 300 $RUN bench.py --gpus 8 --steps 20 --warmup 3 > $OUT/cfg5_default_n8.json 2> $OUT/cfg5_default_n8.err
show $OUT/cfg5_default_n8.json "cfg5 default"
timeout 200 $RUN bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $OUT/ref_n8.json 2> $OUT/ref_n8.err
tail -c 300 $OUT/ref_n8.json; echo
grep -h -v "^$\|OMP_NUM\|\*\*\*\*\|destroy_process_group" $OUT/*.err | tail -5
