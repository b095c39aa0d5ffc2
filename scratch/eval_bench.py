import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qmc4_b200.system import System
for cfgno, W in ((5, 2368*2), (3, 4736*2), (2, 16384)):
    cfg = bench.workload(cfgno, W, 1); cfg["measurements"] = []
    s = System.from_input(cfg)
    R = torch.from_numpy(bench.initial_walkers(cfg, W, 1)).cuda()
    E = torch.empty(W, dtype=torch.float64, device="cuda"); F = torch.empty_like(R)
    def run(n):
        for _ in range(n):
            rc = s.lib.qmc_eval_device(s.ctx, R.data_ptr(), W, None, F.data_ptr(), None, None, E.data_ptr(), None)
            assert rc == 0
        s.synchronize()
    run(3)
    t0 = time.perf_counter(); run(20); dt = (time.perf_counter() - t0) / 20
    pairs = bench.work_per_move(cfg)[0]
    print("eval cfg %d W %d: %.4f ms  %.3e pair-evals/s  (grid %d, ctas/sm %d, regs %d)" % (cfgno, W, dt*1e3, W*pairs/dt, s.info("grid"), s.info("ctas_per_sm"), s.info("regs_eval")))
    s.close()
