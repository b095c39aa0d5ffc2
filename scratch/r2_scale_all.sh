#!/bin/bash
# final scaling table on ONE 8-GPU box (charged 8 x): config 5 weak N=8,4,2,1 and strong N=4,2 (strong N=8 == weak N=8; strong N=1 on the 1-GPU session)
set -u
OUT=gpurun_out/scale_all
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("migration") or {}
    ins = m.get("in_step_migration", {})
    print("%-12s N=%d %.4g moves/s, %.4f ms/DMC step (move %.4f ms, frac %.3f, whole %.3f), e2e %.4g, pop %d, events %s out %s ok %s clocks %s" % (
        sys.argv[2], d["n_gpus"], d["value"], d["ms_per_dmc_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"],
        d["e2e"]["value"], d["population"], ins.get("events_sum_over_ranks"), ins.get("walkers_out"), m.get("ok"), d["clocks"].get("sm_mhz")))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
one() { # N scaling tag
  local N=$1 SC=$2
  if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 2972$N"; fi
  timeout 200 $RUN bench.py --gpus $N --scaling $SC --steps 10 --equil 100 --no-cpu-baseline --no-secondary --watchdog 180 > $OUT/cfg5_${SC}_n$N.json 2> $OUT/cfg5_${SC}_n$N.err
  show $OUT/cfg5_${SC}_n$N.json "cfg5 $SC"
}
one 8 weak
one 4 weak
one 2 weak
one 1 weak
one 4 strong
one 2 strong
grep -h -v "^$\|OMP_NUM\|\*\*\*\*\|destroy_process_group" $OUT/*.err | tail -5
