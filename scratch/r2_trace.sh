#!/bin/bash
set -u
N=${1:-2}
CFG=${2:-2}
STEPS=${3:-10}
OUT=gpurun_out/trace
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
QMC4_TRACE=100 timeout 300 $TR --master-port 29712 bench.py --gpus $N --config $CFG --steps $STEPS --equil 100 --no-cpu-baseline --watchdog 250 > $OUT/t.json 2> $OUT/t.err
python scratch/trace_calls.py $OUT/t.err 0
python - <<'PY'
import json
d=json.loads(open('gpurun_out/trace/t.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_dmc_step'], d['roofline']['avg_launch_ms'], d['migration']['in_step_migration'])
PY
grep -v "qmc4 trace" $OUT/t.err | grep -v "^$\|OMP_NUM\|\*\*\*\*\|destroy_process_group" | tail
gzip -f $OUT/t.err
