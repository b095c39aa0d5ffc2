"""profiling driver: python scratch/prof_cfg.py <config> <walkers> <steps> [sample_every]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qmc4_b200.system import System
cfgno, W, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
se = int(sys.argv[4]) if len(sys.argv) > 4 else 1
cfg = bench.workload(cfgno, W, 1)
s = System.from_input(cfg)
s.upload(bench.initial_walkers(cfg, W, 1))
s.step(steps, se)
b = s.block_results()
pairs = bench.work_per_move(cfg)[0]
print("cfg %d W %d: %.4f ms/step  %.3e moves/s %.3e pairs/s acc %.4f pop %d" % (cfgno, W, b["device_ms"] / b["steps_done"],
      b["n_moves"] / b["device_ms"] * 1e3, b["n_moves"] * pairs / b["device_ms"] * 1e3, b["n_accepted"] / b["n_moves"], b["population"]))
s.close()
