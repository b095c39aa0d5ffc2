#!/bin/bash
# 8-GPU call of round 2 (after scratch/r2_gpu_session.sh has passed on 1 GPU):
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash scratch/r2_gpu8_session.sh'
# Closes the two multi-GPU items round 1 left open: config 5 (N=512, 100k walkers) at 8 GPUs — the north-star's
# 85 % target — with stderr kept and a watchdog, and the 2-GPU hardware run of the record migration.  Charged 8x:
# every step has a short timeout, cheapest first.
set -u
OUT=gpurun_out/r2_gpu8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
step() { echo "=== $1" | tee -a $OUT/session.log; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"

step "1. multi-GPU tests: host driver --gpus 2, record migration on 2 GPUs"
timeout 300 python -m pytest tests/test_host_driver.py tests/test_zz_gpu_new_paths.py -m gpu -q -p no:cacheprovider \
  -k "multi_gpu or two_gpus" > $OUT/pytest_multi.log 2>&1
tail -3 $OUT/pytest_multi.log | tee -a $OUT/session.log

step "2. config 5 weak scaling: 1 GPU then 8 GPUs (12 500 walkers per GPU), stderr kept, watchdog 240 s"
timeout 300 python bench.py --config 5 --steps 20 --equil 50 --no-cpu-baseline --watchdog 240 \
  > $OUT/cfg5_n1.json 2> $OUT/cfg5_n1.err
cut -c1-300 $OUT/cfg5_n1.json | tee -a $OUT/session.log
timeout 400 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --config 5 --steps 20 --equil 50 --no-cpu-baseline --watchdog 300 \
  > $OUT/cfg5_n8.json 2> $OUT/cfg5_n8.err
cut -c1-300 $OUT/cfg5_n8.json | tee -a $OUT/session.log
tail -5 $OUT/cfg5_n8.err | tee -a $OUT/session.log

step "3. config 2 at 8 GPUs: pipeline 1 (default) against pipeline 2 (all-reduce off the critical path)"
for p in 1 2; do
  QMC4_PIPELINE=$p timeout 300 $TR --nproc-per-node 8 --master-port 2955$p bench.py --gpus 8 --steps 200 --no-cpu-baseline --watchdog 240 \
    > $OUT/cfg2_n8_p$p.json 2> $OUT/cfg2_n8_p$p.err
  cut -c1-300 $OUT/cfg2_n8_p$p.json | tee -a $OUT/session.log
done

step "4. config 3 (N=256, S(q), 32k walkers over 8 GPUs = 4096 per GPU)"
timeout 300 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --config 3 --steps 50 --equil 50 --no-cpu-baseline --watchdog 240 \
  > $OUT/cfg3_n8.json 2> $OUT/cfg3_n8.err
cut -c1-300 $OUT/cfg3_n8.json | tee -a $OUT/session.log
ls -la $OUT | tee -a $OUT/session.log
