"""TEST INFRASTRUCTURE: forward walking + superfluidFraction over NR emulated ranks with walker migration.
Every rank runs traced DMC steps (sampling every SE-th step) and rebalances after some events, so walkers change rank
carrying their genealogical records.  The global genealogy is rebuilt from the traces (rows are matched by value: a moved
walker is found wherever it went) and the oracle's per-walker-ring restatement of the reference's storers/estimators must
reproduce the rank-reduced device estimators.
Usage: multirank_fw_driver.py NR SE [MID].  Exits non-zero when a check fails."""
import os, sys, threading, copy
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cases, ctypes as C
from qmc4_b200 import capi
from qmc4_b200.system import System
from oracle.oracle import Port

NR = int(sys.argv[1]) if len(sys.argv) > 1 else 2
SE = int(sys.argv[2]) if len(sys.argv) > 2 else 1
MID = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # 1: also rebalance between sampling events (needs SE > 1)
K = 9
REBALANCE_AFTER = (1, 2, 4, 7)          # events after which the ranks rebalance (records of 2..4 levels travel)
cfg = copy.deepcopy(cases.forward_walking_cases()["pbc_pair"])
cfg["measurements"].append({"kind": "superfluidFraction", "label": "rho", "sets": [0, 0], "bins": 4})
WPR = 10
cfg.update(walkers=WPR * NR, delta_walkers=max(1, WPR * NR // 8))
N = cfg["N"][0]
lib = capi.load()
buf = C.create_string_buffer(128)
assert lib.qmc_comm_unique_id(buf) == 0
uid = buf.raw
bar = threading.Barrier(NR)
pub = [[None] * NR for _ in range(K)]        # per event, per rank: (population after the event, ancestor ROWS)
res = [None] * NR
xchg = [None] * NR
err = []


def run(r):
    try:
        # QMC4_FW_DRIVER_REAL_GPUS=1 (tests/test_zz_gpu_new_paths.py on a multi-GPU box): rank r on cuda:r with the real NCCL
        s = System.from_input(cfg, device=r if os.environ.get("QMC4_FW_DRIVER_REAL_GPUS") == "1" else 0, comm=(uid, r, NR), capacity=8 * WPR)
        rng = np.random.default_rng(50 + r)
        L = np.asarray(cfg["lBox"])
        n0 = WPR + (6 if r == 0 else -3)                      # unequal start: the first rebalance moves walkers
        R = (rng.uniform(-0.5, 0.5, size=(n0, 3, N)) * L[None, :, None]).reshape(n0, -1)
        s.upload(R)
        anc_rows, moved = R, []                               # for every current walker: the configuration of its ancestor at the last event

        def rebalance():
            nonlocal anc_rows
            Rb, _ = s.download(with_energy=True)
            xchg[r] = (Rb, anc_rows)
            bar.wait()
            moved.append(s.rebalance())
            Ra, _ = s.download(with_energy=True)
            lookup = {}
            for q in range(NR):
                for row, a in zip(*xchg[q]):
                    lookup.setdefault(row.tobytes(), a)       # clones share their lineage
            anc_rows = np.stack([lookup[row.tobytes()] for row in Ra])
            bar.wait()

        for t in range(K * SE):
            sample = (t + 1) % SE == 0
            tr = s.step_traced(accumulate=sample)
            step_anc = np.repeat(np.arange(tr["n_before"]), tr["n_desc"][:tr["n_before"]])
            anc_rows = anc_rows[step_anc]
            e = (t + 1) // SE - 1
            if sample:
                Re, _ = s.download(with_energy=True)
                pub[e][r] = (Re, anc_rows)
                anc_rows = Re
                if e in REBALANCE_AFTER:
                    rebalance()
            elif MID and (t // SE) in REBALANCE_AFTER and t % SE == 0:
                rebalance()                                   # between two sampling events: the composed ancestor map is what travels
        out = {}
        for e_ in s.estimators:
            out[e_["label"]] = s.estimator(e_["index"])
        res[r] = dict(est=out, moved=moved, depth=s.info("fw_link_depth"))
        s.close()
    except Exception as ex:
        import traceback; traceback.print_exc()
        err.append((r, ex))
        bar.abort()


th = [threading.Thread(target=run, args=(r,)) for r in range(NR)]
[t.start() for t in th]
[t.join() for t in th]
ok = not err
if ok:
    pops, ancs = [], []
    for e in range(K):
        cur = np.concatenate([pub[e][r][0] for r in range(NR)])
        if e == 0:
            anc = np.arange(len(cur), dtype=np.int32)
        else:
            index = {}
            for i, row in enumerate(pops[-1]):
                index.setdefault(row.tobytes(), i)               # clones share their lineage: any copy will do
            anc = np.asarray([index[row.tobytes()] for r in range(NR) for row in pub[e][r][1]], dtype=np.int32)
        pops.append(cur)
        ancs.append(anc)
    port = Port()
    iM, meas = cases.fw_measurement(cfg)
    print("moved per rank:", [o["moved"] for o in res], "populations:", [len(p) for p in pops])
    ok &= any(m != 0 for o in res for m in o["moved"])                                   # records did travel
    for steps in meas["forwardWalkingSteps"]:
        exp, we = port.forward_walking(cfg, pops, ancs, iM, meas["recordSteps"], steps, cases.meas_bins(meas))
        for r in range(NR):                                                              # reduced over ranks: same everywhere
            got, w = res[r]["est"]["%s_fw%d" % (meas["label"], steps)]
            good = w == we and w > 0 and np.abs(got - exp).max() <= 1e-10 * np.abs(exp).max()
            print("fw", steps, "rank", r, "weight", w, we, "ok" if good else "MISMATCH")
            ok &= bool(good)
    exp, we = port.superfluid_fraction(cfg, pops, ancs, len(cfg["measurements"]) - 1)
    for r in range(NR):
        got, w = res[r]["est"]["rho"]
        good = np.array_equal(w, we) and np.abs(got - exp).max() <= 1e-10 * np.abs(exp).max()
        print("superfluid rank", r, w, "ok" if good else "MISMATCH")
        ok &= bool(good)
print("OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
