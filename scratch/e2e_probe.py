"""Where does the e2e bench step spend its time?  python scratch/e2e_probe.py [config]"""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from qmc4_b200 import capi
from qmc4_b200.system import System
cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 5
wpg = bench.DEFAULT_WALKERS[cfgno]
cfg = bench.workload(cfgno, wpg, 1)
s = System.from_input(cfg)
s.upload(bench.initial_walkers(cfg, wpg, 1000))
s.step(50, 1); s.block_results()
Rh, _ = s.download(with_energy=True)
nW, dofs = len(Rh), s.dofs
pin_in = torch.from_numpy(Rh.copy()).pin_memory()
capw = int(s.info("capacity"))
pin_out = torch.empty((capw, dofs), dtype=torch.float64).pin_memory()
pin_E = torch.empty(capw, dtype=torch.float64).pin_memory()
lib = s.lib
dmc_steps = bench.DMC_STEPS_PER_STEP[cfgno]
for it in range(4):
    t = [time.perf_counter()]
    lib.qmc_upload_walkers(s.ctx, C.cast(pin_in.data_ptr(), capi.c_double_p), nW); t.append(time.perf_counter())
    s.step(dmc_steps, 1); t.append(time.perf_counter())
    br = s.block_results(); t.append(time.perf_counter())
    n = C.c_int64(0)
    lib.qmc_download_walkers(s.ctx, C.cast(pin_out.data_ptr(), capi.c_double_p), pin_out.shape[0], C.byref(n), C.cast(pin_E.data_ptr(), capi.c_double_p)); t.append(time.perf_counter())
    print("iter %d: upload %.1f ms, step %.1f ms, block_results %.1f ms, download %.1f ms (nW %d, %.0f MB each way)" % (
        it, 1e3 * (t[1] - t[0]), 1e3 * (t[2] - t[1]), 1e3 * (t[3] - t[2]), 1e3 * (t[4] - t[3]), nW, nW * dofs * 8 / 1e6))
