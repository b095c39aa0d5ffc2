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
assert lib.qmc_comm_unique_id(buf) == 0
uid2 = buf.raw
assert lib.qmc_comm_unique_id(buf) == 0
uid3 = buf.raw
out = [None] * NR
err = []

def run(r):
    try:
        rng = np.random.default_rng(100 + r)
        L = np.asarray(cfg["lBox"])
        n0 = WPR + (r * 5) % 7 - 3          # unequal start
        R = (rng.uniform(-0.5, 0.5, size=(n0, 3, N)) * L[None, :, None]).reshape(n0, -1)
        # ---- (A) in-step migration (one step late, dead band 0 = the reference's behaviour): every step must move exactly
        # the global population — a lost or duplicated walker shows up in the all-reduced move counter
        s = System.from_input(cfg, device=0, comm=(uid, r, NR))
        s.upload(R)
        assert lib.qmc_set_migration(s.ctx, 1, 0.0) == 0
        moves_ok = True
        for k in range(8):
            gpop = s.info("global_population")
            s.step(1, 1)
            bk = s.block_results()
            if not (bk["n_moves"] == gpop and bk["steps_done"] == 1):
                print("rank", r, "step", k, "moved", bk["n_moves"], "global population", gpop, "steps", bk["steps_done"])
                moves_ok = False
        Rm, Em = s.download(with_energy=True)
        ev = s.eval(Rm)          # the payload that travelled is intact: E stored with a walker = E_L of its configuration
        if not np.allclose(ev["E"], Em, rtol=1e-9, atol=1e-9):
            print("rank", r, "E mismatch", np.abs(ev["E"] - Em).max())
            moves_ok = False
        mig = (s.info("migration_events"), s.info("migrated_out"), s.info("migrated_in"))
        gpop_end, pop_end = s.info("global_population"), len(Rm)
        s.close()
        # ---- (A') exact identity test of the in-step path: with timeStep = 1e-300 a move changes nothing in fp64 and every
        # walker has exactly one descendant, so the population is frozen and only the migration moves walkers between the
        # ranks (some ranks both send and receive in the alias plan): the global multiset must be preserved bit for bit
        s = System.from_input(dict(cfg, timeStep=1e-300), device=0, comm=(uid3, r, NR))
        s.upload(R)
        assert lib.qmc_set_migration(s.ctx, 1, 0.0) == 0
        s.step(3, 1)
        Rz = s.download()
        frozen = (R, Rz, s.info("migrated_out"), s.info("migrated_in"))
        s.close()
        # ---- (B) the explicit call (qmc_rebalance) with the in-step migration switched off
        s = System.from_input(cfg, device=0, comm=(uid2, r, NR))
        s.upload(R)
        assert lib.qmc_set_migration(s.ctx, 1, 1e9) == 0
        s.step(10, 1)
        b0 = s.block_results()
        Rb, Eb = s.download(with_energy=True)
        moved = s.rebalance()
        Ra, Ea = s.download(with_energy=True)
        s.step(10, 1)
        b1 = s.block_results()
        if s.info("migration_events") != (1 if moved else 0):
            print("rank", r, "explicit phase: migration events", s.info("migration_events"), "moved", moved)
            moves_ok = False
        out[r] = dict(b0=b0, b1=b1, Rb=Rb, Ra=Ra, Eb=Eb, Ea=Ea, moved=moved, wpw=s.info("warps_per_walker"), moves_ok=moves_ok, mig=mig,
                      n0=n0, gpop_end=gpop_end, pop_end=pop_end, frozen=frozen)
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
ok &= all(o["moves_ok"] for o in out)                                              # every step moved the whole global population once
print("in-step migration (events, out, in) per rank:", [o["mig"] for o in out], "initial", [o["n0"] for o in out])
ok &= sum(o["mig"][1] for o in out) == sum(o["mig"][2] for o in out) > 0           # walkers left = walkers arrived, and some did
ok &= sum(o["pop_end"] for o in out) == out[0]["gpop_end"]                          # the ranks' populations add up to the all-reduced one
fz0 = np.concatenate([o["frozen"][0] for o in out]); fz1 = np.concatenate([o["frozen"][1] for o in out])
pz = [len(o["frozen"][1]) for o in out]
print("frozen-population run: multiset preserved by the in-step migration:", np.array_equal(key(fz0), key(fz1)), "populations", pz,
      "moved", sum(o["frozen"][2] for o in out))
ok &= np.array_equal(key(fz0), key(fz1)) and max(pz) - min(pz) <= 1 and sum(o["frozen"][2] for o in out) == sum(o["frozen"][3] for o in out) > 0
print("OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
