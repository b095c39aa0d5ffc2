#!/bin/bash
set -u
N=2
OUT=gpurun_out/multi_dbg
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
run() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout 200 $TR --master-port 29711 bench.py --gpus $N --config 2 --steps 5 --equil 100 --no-cpu-baseline --watchdog 150 $EXTRA > $OUT/$name.json 2> $OUT/$name.err
  python - $OUT/$name.json $name <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s %.4g moves/s, %.4f ms/DMC step (launch %.4f ms), e2e %.4g, events %s" % (sys.argv[2], d["value"], d["ms_per_dmc_step"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"], d["migration"]["in_step_migration"]["events_sum_over_ranks"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
}
EXTRA="" run base A=1
EXTRA="--migrate-band 1e9" run band_inf A=1
EXTRA="--migrate-every 1000000" run every_inf A=1
EXTRA="" run nosampler QMC4_BENCH_NO_SAMPLER=1
EXTRA="--no-flush" run noflush A=1
EXTRA="--dmc-steps 10 --steps 50" run short_steps A=1
