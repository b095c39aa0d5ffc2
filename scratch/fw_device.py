"""forward walking on the (emulated or real) device vs the oracle port fed with the traced genealogy"""
import sys, copy
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, cases
from oracle.oracle import Port
from qmc4_b200.system import System

N, L, W = 8, 2.2, 24
RS, FWS, K, bins = 4, [0, 1, 3], 10, 16
sample_every = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = copy.deepcopy(cases.pbc(N, L, walkers=W))
cfg.update(timeStep=2e-2, delta_walkers=3, seed=11)
cfg["measurements"] = [{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": bins, "minx": 0.0, "maxx": L / 2,
                        "recordSteps": RS, "forwardWalkingSteps": FWS}]
s = System.from_input(cfg)
print([e["label"] for e in s.estimators], s.info("fw_depth"))
rng = np.random.default_rng(3)
R0 = (rng.uniform(-0.5, 0.5, size=(W, 3, N)) * L).reshape(W, -1)
s.upload(R0)
pops, ancs = [], []
comp = None
for t in range(K * sample_every):
    sample = (t + 1) % sample_every == 0
    tr = s.step_traced(accumulate=sample)
    anc_step = np.repeat(np.arange(tr["n_before"]), tr["n_desc"][:tr["n_before"]])      # post-branch walker -> walker index of this step
    assert len(anc_step) == tr["n_after"], (len(anc_step), tr["n_after"])
    comp = anc_step if comp is None else comp[anc_step]
    if sample:
        R, _ = s.download(with_energy=True)
        pops.append(R)
        ancs.append(comp.astype(np.int32) if len(pops) > 1 else np.arange(len(R), dtype=np.int32))
        comp = None
print("populations", [len(p) for p in pops], "events", s.info("fw_events"))
port = Port()
for k, fw in enumerate(FWS):
    e = s.estimators[k]
    got, w = s.estimator(e["index"])
    exp, we = port.forward_walking(cfg, pops, ancs, 0, RS, fw, bins)
    print(e["label"], "weight", w, we, "maxabs", np.abs(got - exp).max(), "scale", np.abs(exp).max())
plain, wp = s.estimator(s.hidden_estimators[0]["index"])
print("plain weight", wp)
