import sys, copy
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, cases
from oracle.oracle import Ref, Port

def genealogy(cfg, K, n0, rng):
    N = cfg["N"][0]; L = np.asarray(cfg["lBox"])
    pops, ancs = [], []
    n = n0
    for e in range(K):
        if e == 0:
            anc = np.arange(n)
        else:
            nprev = len(pops[-1])
            n = max(2, nprev + rng.integers(-2, 3))
            anc = np.sort(rng.integers(0, nprev, size=n))
        R = (rng.uniform(-0.5, 0.5, size=(n, 3, N)) * L[None, :, None]).reshape(n, -1)
        pops.append(R); ancs.append(anc.astype(np.int32))
    return pops, ancs

cfg = copy.deepcopy(cases.pbc(N=6, L=3.0))
cfg["measurements"] = [{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 12, "minx": 0.0, "maxx": 1.5,
                        "recordSteps": 4, "forwardWalkingSteps": [0, 1, 3]}]
rng = np.random.default_rng(5)
pops, ancs = genealogy(cfg, 9, 5, rng)
ref, port = Ref(), Port()
for k, fw in enumerate([0, 1, 3]):
    a, wa = ref.forward_walking(cfg, pops, ancs, k, 12)
    b, wb = port.forward_walking(cfg, pops, ancs, 0, 4, fw, 12)
    print(fw, wa, wb, np.abs(a - b).max(), a.sum())
