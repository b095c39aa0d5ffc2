"""TEST INFRASTRUCTURE: drive NR emulated ranks (host threads of this process, tests/emu build of the
library + tests/emu/libnccl_emu.so) through the multi-rank sequence bench.py and dmc_b200 use:
comm_init -> upload (unequal populations) -> steps -> block results -> rebalance -> steps.
Usage: multirank_driver.py NR N walkers_per_rank.  Exits non-zero when an invariant fails."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cases, ctypes as C
from qmc4_b200 import capi
from qmc4_b200.system import System
import copy

NR = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
WPR = int(sys.argv[3]) if len(sys.argv) > 3 else 24
cfg = copy.deepcopy(cases.pbc(N=N, L=(N / 1.0) ** (1 / 3.)))
cfg.update(walkers=WPR * NR, delta_walkers=max(1, WPR * NR // 10), timeStep=1e-3, seed=7)
lib = capi.load()
buf = C.create_string_buffer(128)
assert lib.qmc_comm_unique_id(buf) == 0
uid = buf.raw
out = [None] * NR
err = []

def run(r):
    try:
        s = System.from_input(cfg, device=0, comm=(uid, r, NR))
        rng = np.random.default_rng(100 + r)
        L = np.asarray(cfg["lBox"])
        n0 = WPR + (r * 5) % 7 - 3          # unequal start
        R = (rng.uniform(-0.5, 0.5, size=(n0, 3, N)) * L[None, :, None]).reshape(n0, -1)
        s.upload(R)
        s.step(10, 1)
        b0 = s.block_results()
        Rb, Eb = s.download(with_energy=True)
        moved = s.rebalance()
        Ra, Ea = s.download(with_energy=True)
        s.step(10, 1)
        b1 = s.block_results()
        out[r] = dict(b0=b0, b1=b1, Rb=Rb, Ra=Ra, Eb=Eb, Ea=Ea, moved=moved, wpw=s.info("warps_per_walker"))
        s.close()
    except Exception as e:
        import traceback; traceback.print_exc()
        err.append((r, e))

t0 = time.time()
th = [threading.Thread(target=run, args=(r,)) for r in range(NR)]
[t.start() for t in th]
[t.join() for t in th]
print("elapsed", time.time() - t0, "errors", err)
for r in range(NR):
    o = out[r]
    print(r, "pop before", len(o["Rb"]), "after", len(o["Ra"]), "moved", o["moved"], "E_T", o["b0"]["energy_shift"], o["b1"]["energy_shift"], "moves", o["b0"]["n_moves"], "wpw", o["wpw"])
before = np.concatenate([o["Rb"] for o in out]); after = np.concatenate([o["Ra"] for o in out])
key = lambda A: A[np.lexsort(A.T[::-1])]
print("multiset preserved:", np.array_equal(key(before), key(after)), before.shape)
eb = np.sort(np.concatenate([o["Eb"] for o in out])); ea = np.sort(np.concatenate([o["Ea"] for o in out]))
print("energies preserved:", np.array_equal(eb, ea))

ok = not err
ok &= np.array_equal(key(before), key(after)) and np.array_equal(eb, ea)          # migration moves walkers, nothing else
pops = [len(o["Ra"]) for o in out]
ok &= max(pops) - min(pops) <= 1 and sum(pops) == len(before)                      # balanced, conserved
ok &= sum(o["moved"] for o in out) == 0
for k in ("energy_shift", "n_moves", "n_accepted", "energy_sum", "energy_count"):  # block results are global on every rank
    ok &= len({o["b0"][k] for o in out}) == 1 and len({o["b1"][k] for o in out}) == 1
ok &= out[0]["b0"]["n_moves"] == len(before) * 0 + out[0]["b0"]["n_moves"]
print("OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
