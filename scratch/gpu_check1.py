"""First on-GPU validation (scratch, not part of the test suite)."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qmc4_b200.system import System
from oracle.oracle import Ref

ref = Ref()
rng = np.random.default_rng(1)

def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

def cfg_pbc(N, L, meas=None, jast="gaussiangaussian"):
    J = {"kind": "gaussiangaussian", "r0": 1.0, "u0": -0.5} if jast == "gaussiangaussian" else jast
    c = {"N": [N], "lBox": [L, L, L], "method": "dmc", "timeStep": 1e-3, "walkers": 64, "seed": 7,
         "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": J}],
         "potentials": [{"kind": "potentialTwoBody", "sets": [0, 0], "functor": {"kind": "potentialBarrier", "R0": 1.0, "V0": 2.0}}],
         "measurements": meas or []}
    return c

def check_eval(name, cfg, R):
    s = System.from_input(cfg)
    t0 = time.time(); o = s.eval(R); t1 = time.time()
    r = ref.eval(cfg, R)
    print("%-28s wpw=%d ksel=%d  logpsi %.2e grad %.2e lap %.2e V %.2e E %.2e valid %s  (%.3fs)" % (
        name, s.info("warps_per_walker"), s.info("kernel_variant"), rel(o["logpsi"], r["logpsi"]), rel(o["grad"], r["grad"]),
        rel(o["lap"], r["lap"]), rel(o["V"], r["V"]) if np.abs(r["V"]).max() > 0 else np.abs(o["V"]).max(), rel(o["E"], r["E"]),
        bool((o["valid"] == r["valid"]).all()), t1 - t0))
    s.close()

for N, L in ((10, 3.0), (32, 3.2), (33, 3.2), (64, 4.0), (100, 4.7), (128, 5.0), (256, 6.3496), (512, 8.0)):
    nW = 8
    R = rng.uniform(-L, L, size=(nW, 3, N))
    check_eval("pbc GG+barrier N=%d" % N, cfg_pbc(N, L), R)

# trap config 1
c1 = {"N": [10], "lBox": [1e4] * 3, "method": "dmc", "timeStep": 1e-3, "walkers": 100, "seed": 3,
      "wavefunctions": [{"kind": "jastrow1b", "set": 0, "jastrow": {"kind": "gaussian", "alpha": 0.5}},
                        {"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "gaussian", "alpha": 0.05}}],
      "potentials": [{"kind": "harmonic1b", "omega": 1.0, "set": 0}], "measurements": []}
check_eval("trap N=10", c1, rng.normal(0, 0.7, size=(16, 3, 10)))

# generic functors
sw = {"kind": "squareWell", "V0": 2.4674011002723395, "R0": 1.0, "Rm": 1.6, "alpha": 0.5, "aInverse": 0.0, "cut_off": 2.5}
check_eval("squareWell N=20", cfg_pbc(20, 5.0, jast=sw), rng.uniform(-2.5, 2.5, size=(8, 3, 20)))
hs = {"kind": "hardSphereGauss", "a": 0.1, "Rm": 0.8, "D": 2.0}
check_eval("hardSphereGauss N=40", cfg_pbc(40, 6.0, jast=hs), rng.uniform(-3, 3, size=(8, 3, 40)))

# mixture: Dis term
cm = {"N": [12, 20], "lBox": [5.0, 5.0, 5.0], "method": "dmc", "timeStep": 1e-3, "walkers": 10, "seed": 3,
      "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "gaussiangaussian", "r0": 1.0, "u0": -0.5}},
                        {"kind": "jastrow2bUnDis", "sets": [1, 1], "jastrow": {"kind": "gaussiangaussian", "r0": 0.8, "u0": -0.3}},
                        {"kind": "jastrow2bDis", "sets": [0, 1], "jastrow": {"kind": "gaussiangaussian", "r0": 1.2, "u0": 0.4}}],
      "potentials": [{"kind": "potentialTwoBody", "sets": [0, 1], "functor": {"kind": "squareWell", "R0": 1.0, "V0": 2.0}}],
      "measurements": []}
check_eval("mixture 12+20", cm, rng.uniform(-2.5, 2.5, size=(8, 32 * 3)))

# three body
c4 = {"N": [32], "lBox": [5.0] * 3, "method": "dmc", "timeStep": 1e-3, "walkers": 10, "seed": 3,
      "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "gaussiangaussian", "r0": 1.0, "u0": -0.5}},
                        {"kind": "jastrow3bUnDis", "sets": [0], "jastrow": {"kind": "gaussiangaussian", "r0": 1.5, "u0": -0.2}}],
      "potentials": [{"kind": "potentialThreeBodyUnDis", "sets": [0, 0, 0], "functor": {"kind": "squareWell", "R0": 1.5, "V0": 1.0}}],
      "measurements": []}
check_eval("three-body N=32", c4, rng.uniform(-2.5, 2.5, size=(8, 3, 32)))

# ---- traced step vs reference pieces
cfg = cfg_pbc(64, 4.0)
cfg["walkers"] = 256
s = System.from_input(cfg)
R0 = rng.uniform(-2, 2, size=(256, 3, 64))
s.upload(R0)
e0 = ref.eval(cfg, R0)
tr = s.step_traced()
par = tr["parent"]
Rp = ref.drift_diffuse(cfg, 1e-3, R0[par].reshape(256, -1), e0["grad"][par], tr["xi"])
ep = ref.eval(cfg, Rp)
q = ref.green_ratio(cfg, 1e-3, Rp, ep["grad"], R0[par].reshape(256, -1), e0["grad"][par])
print("trace: Rprop %.2e Fprop %.2e logpsi %.2e E %.2e q %.2e acc %.3f n_after %d xi mean %.3f std %.3f" % (
    rel(tr["Rprop"], Rp), rel(tr["Fprop"], ep["grad"]), rel(tr["logpsi"], ep["logpsi"]), rel(tr["Eprop"], ep["E"]),
    rel(tr["q"], q), tr["accepted"].mean(), tr["n_after"], tr["xi"].mean(), tr["xi"].std()))
A = 2 * (ep["logpsi"] - e0["logpsi"][par]) + q
acc = (A > 0) | (A > np.log(tr["u_metropolis"]))
print("accept decisions equal:", bool((acc == (tr["accepted"] != 0)).all()))
s.close()

# ---- short DMC, trap: energy should approach 15 + interaction shift
c1["walkers"] = 1000
s = System.from_input(c1)
s.upload(np.tile(rng.normal(0, 0.7, size=(1, 30)), (1000, 1)))
for b in range(5):
    s.step(200, 1)
    r = s.block_results()
    print("block %d E=%.5f pop=%d acc=%.5f eT=%.4f ms/step=%.4f overflow=%d" % (
        b, r["energy_sum"] / r["energy_count"], r["population"], r["n_accepted"] / r["n_moves"], r["energy_shift"],
        r["device_ms"] / r["steps_done"], r["overflow"]))
print("fp64 peak TFLOP/s:", s.info("fp64_peak_tflops"), "regs", s.info("regs_propagate"))
s.close()

# ---- throughput probe config 2 and N=512
for N, L, W in ((64, 4.0, 16384), (256, 6.3496, 4096), (512, 8.0, 2048)):
    cfg = cfg_pbc(N, L); cfg["walkers"] = W
    s = System.from_input(cfg)
    s.upload(rng.uniform(-L / 2, L / 2, size=(W, 3 * N)))
    s.step(5, 1); s.block_results()
    s.step(20, 1); r = s.block_results()
    ms = r["device_ms"] / r["steps_done"]
    pairs = r["n_moves"] / r["steps_done"] * N * (N - 1) / 2
    print("N=%d W=%d: %.3f ms/step  %.3e pair-evals/s  acc %.4f pop %d E %.4f grid %d wpw %d smem %d" % (
        N, W, ms, pairs / (ms * 1e-3), r["n_accepted"] / r["n_moves"], r["population"], r["energy_sum"] / r["energy_count"],
        s.info("grid"), s.info("warps_per_walker"), s.info("smem_bytes")))
    s.close()
