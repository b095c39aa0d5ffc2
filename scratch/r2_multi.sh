#!/bin/bash
# multi-GPU session: bash scratch/r2_multi.sh N [configs...]   (charged N x)
set -u
N=${1:-2}; shift
CFGS=${@:-"5 2"}
OUT=gpurun_out/multi_n$N
mkdir -p $OUT
export PYTHONUNBUFFERED=1 NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
if [ "$N" = "2" ]; then
  timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "multi_gpu or two_gpus or two_gpu" > $OUT/pytest_multi.log 2>&1
  tail -4 $OUT/pytest_multi.log
fi
for cfg in $CFGS; do
  timeout 200 $TR --master-port 2970$cfg bench.py --gpus $N --config $cfg --steps 10 --no-cpu-baseline --watchdog 150 \
    > $OUT/cfg${cfg}.json 2> $OUT/cfg${cfg}.err
  echo "cfg$cfg rc=$?"
  python - $OUT/cfg${cfg}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N=%d %s: %.4g moves/s, %.4f ms/DMC step, frac %.3f (launch %.4f ms), e2e %.4g, pop %s" % (d["n_gpus"], d["config"]["workload"][:7], d["value"], d["ms_per_dmc_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"], d["population"]))
    print("   migration:", json.dumps(d.get("migration")))
except Exception as e:
    print("no result", e)
PY
  grep -v "^$\|OMP_NUM\|\*\*\*\*\|destroy_process_group" $OUT/cfg${cfg}.err | tail -15
done
