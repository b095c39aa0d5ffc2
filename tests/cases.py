"""Shared parity cases: the reference's input.json documents + seeded walker sets.

Sizes are small enough for the CPU oracle to finish in seconds.  The same dictionaries are fed,
unchanged, to the reference (oracle/_ref), the C restatement (oracle/qmc_oracle.c) and the CUDA
product (through the C ABI)."""
import copy

import numpy as np

GG = {"kind": "gaussiangaussian", "r0": 1.0, "u0": -0.5}


def pbc(N, L, jastrow=GG, pot="potentialBarrier", meas=None, walkers=64, **kw):
    c = {"N": [N], "lBox": [L, L, L], "method": "dmc", "timeStep": 1e-3, "walkers": walkers, "seed": 7,
         "stepsPerBlock": 10, "nBlocks": 1, "correlationSteps": 0,
         "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": copy.deepcopy(jastrow)}],
         "potentials": [] if pot is None else [{"kind": "potentialTwoBody", "sets": [0, 0], "functor": {"kind": pot, "R0": 1.0, "V0": 2.0}}],
         "measurements": meas or []}
    c.update(kw)
    return c


def trap(N=10, walkers=100):
    return {"N": [N], "lBox": [1e4, 1e4, 1e4], "method": "dmc", "timeStep": 1e-3, "walkers": walkers, "seed": 1235,
            "stepsPerBlock": 10, "nBlocks": 1, "correlationSteps": 0,
            "wavefunctions": [{"kind": "jastrow1b", "set": 0, "jastrow": {"kind": "gaussian", "alpha": 0.5}},
                              {"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "gaussian", "alpha": 0.05}}],
            "potentials": [{"kind": "harmonic1b", "omega": 1.0, "set": 0}], "measurements": []}


def three_body(N=16, L=5.0, hs=None, meas=True):
    w3 = {"kind": "jastrow3bUnDis", "sets": [0], "jastrow": {"kind": "gaussiangaussian", "r0": 1.5, "u0": -0.2}}
    if hs is not None:
        w3["hardSphereRadius"] = hs
    return {"N": [N], "lBox": [L] * 3, "method": "dmc", "timeStep": 1e-3, "walkers": 32, "seed": 5,
            "stepsPerBlock": 10, "nBlocks": 1, "correlationSteps": 0,
            "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": dict(GG)}, w3],
            "potentials": [{"kind": "potentialThreeBodyUnDis", "sets": [0, 0, 0], "functor": {"kind": "squareWell", "R0": 1.5, "V0": 1.0}}],
            "measurements": [{"kind": "trimerCorrelationUnDis", "label": "trimer", "sets": [0], "bins": 50, "minx": 0, "maxx": L / 2}] if meas else []}


def mixture():
    return {"N": [12, 20], "lBox": [5.0, 5.0, 5.0], "method": "dmc", "timeStep": 1e-3, "walkers": 16, "seed": 3,
            "stepsPerBlock": 10, "nBlocks": 1, "correlationSteps": 0,
            "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": dict(GG)},
                              {"kind": "jastrow2bUnDis", "sets": [1, 1], "jastrow": {"kind": "gaussiangaussian", "r0": 0.8, "u0": -0.3}},
                              {"kind": "jastrow2bDis", "sets": [0, 1], "jastrow": {"kind": "gaussiangaussian", "r0": 1.2, "u0": 0.4}},
                              {"kind": "jastrow1b", "set": 1, "jastrow": {"kind": "gaussian", "alpha": 0.01}}],
            "potentials": [{"kind": "potentialTwoBody", "sets": [0, 1], "functor": {"kind": "squareWell", "R0": 1.0, "V0": 2.0}},
                           {"kind": "squareWell2b", "sets": [0, 0], "V0": 1.5, "R0": 0.9},
                           {"kind": "dipolar2b", "sets": [1, 1], "D": 0.3}],
            "measurements": [{"kind": "pairCorrelation", "label": "g01", "sets": [0, 1], "bins": 40, "minx": 0, "maxx": 2.5},
                             {"kind": "pairCorrelation", "label": "g00", "sets": [0, 0], "bins": 30, "minx": 0.1, "maxx": 2.5}]}


def mixture3(hs=None, N=(5, 4, 3), L=4.0):
    """three DIFFERENT species: jastrow3bDis + potentialThreeBodyDis + trimerCorrelationDis"""
    w3 = {"kind": "jastrow3bDis", "sets": [0, 1, 2], "jastrow": {"kind": "gaussiangaussian", "r0": 1.5, "u0": -0.2}}
    if hs is not None:
        w3["hardSphereRadius"] = hs
    return {"N": list(N), "lBox": [L] * 3, "method": "dmc", "timeStep": 1e-3, "walkers": 24, "seed": 9,
            "stepsPerBlock": 10, "nBlocks": 1, "correlationSteps": 0,
            "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": dict(GG)},
                              {"kind": "jastrow2bDis", "sets": [1, 2], "jastrow": {"kind": "gaussiangaussian", "r0": 0.9, "u0": 0.3}},
                              w3],
            "potentials": [{"kind": "potentialThreeBodyDis", "sets": [0, 1, 2], "functor": {"kind": "squareWell", "R0": 1.8, "V0": 1.0}},
                           {"kind": "potentialThreeBodyDis", "sets": [0, 1, 2], "functor": {"kind": "potentialBarrier", "R0": 1.2, "V0": 0.7}}],
            "measurements": [{"kind": "trimerCorrelationDis", "label": "trimerDis", "sets": [0, 1, 2], "bins": 20, "minx": 0.0, "maxx": 2.0},
                             {"kind": "centerOfMassSquared", "label": "cm2", "set": 1}]}


def impurity(N=(6, 5, 1), L=4.0):
    """two species around ONE impurity (set C holds one particle): trimerCorrelation + trimerhyperRadius"""
    return {"N": list(N), "lBox": [L] * 3, "method": "dmc", "timeStep": 1e-3, "walkers": 24, "seed": 10,
            "stepsPerBlock": 10, "nBlocks": 1, "correlationSteps": 0,
            "wavefunctions": [{"kind": "jastrow2bDis", "sets": [0, 2], "jastrow": dict(GG)},
                              {"kind": "jastrow2bDis", "sets": [1, 2], "jastrow": {"kind": "gaussiangaussian", "r0": 0.8, "u0": -0.4}},
                              {"kind": "jastrow3bDis", "sets": [0, 1, 2], "jastrow": {"kind": "poschTeller", "R0": 1.0, "C": 1.0, "alpha": 0.5, "Rm": 1.5, "cut_off": 4.0}}],
            "potentials": [{"kind": "potentialTwoBody", "sets": [0, 2], "functor": {"kind": "squareWell", "R0": 1.0, "V0": 2.0}}],
            "measurements": [{"kind": "trimerCorrelation", "label": "trimer", "sets": [0, 1, 2], "bins": 16, "minx": 0.0, "maxx": 2.0},
                             {"kind": "trimerhyperRadius", "label": "hyper", "sets": [0, 1, 2]},
                             {"kind": "centerOfMassSquared", "label": "cm2", "set": 0}]}


DIPOLAR_REP = {"kind": "dipolar_rep", "D": 1.3, "cut_off": 2.5, "matching_point": 1.4, "alpha": 0.6, "C": 0.8}
NEW_PATH_EVAL_CASES = ("mixture3", "mixture3_hs", "impurity", "jastrow_dipolar_rep")     # GPU-tested from tests/test_zz_gpu_new_paths.py


def three_set_hist_cases():
    """observables over three sets and scalar observables, evaluated on given configurations"""
    return {"mixture3": mixture3(), "impurity": impurity()}


def meas_bins(m):
    """bins of a measurement: scalar observables are 1-bin estimators (sum, count)"""
    return int(m.get("bins", 1))


SPLINE_COEF = [-6.80202700e-05, 3.40101350e-05, -6.80202700e-05, -3.74111485e-04] + [-(0.0101010101 * (i - 1)) ** 2 for i in range(4, 102)]

JASTROWS = {
    "squareWell": {"kind": "squareWell", "V0": 2.4674011002723395, "R0": 1.0, "Rm": 1.6, "alpha": 0.5, "aInverse": 0.0, "cut_off": 2.5},
    "hardSphereGauss": {"kind": "hardSphereGauss", "a": 0.05, "Rm": 0.8, "D": 2.0},
    "poschTeller": {"kind": "poschTeller", "R0": 1.0, "C": 1.0, "alpha": 0.5, "Rm": 1.5, "cut_off": 2.5},
    "hardSphere": {"kind": "hardSphere", "a": 0.01, "k": 0.6, "cut_off": 2.0},
    "squareWellBoundState": {"kind": "squareWellBoundState", "V0": 6.0, "R0": 1.0, "Rm": 1.5, "cut_off": 2.5, "aInverse": 0.5, "eb": 0.25, "alpha": 0.7},
    "bSpline": {"kind": "bSpline", "stepSize": 0.025, "valueRight": -1.0, "derivativeRight": -0.5,
                "coefficients": [-(0.025 * (i - 1)) ** 2 * 0.16 for i in range(0, 103)]},
    "delta_phonons": {"kind": "delta_phonons", "k": 0.9, "delta": 0.35, "beta": 1.3, "z": 0.8, "cut_off": 2.2, "g": 2.0},
    "delta_phonons_ginf": {"kind": "delta_phonons", "k": 0.7, "delta": 0.5, "beta": 0.9, "z": 1.1, "cut_off": 2.4, "g": "inf"},
    "delta_in_trap": {"kind": "delta_in_trap", "a": 0.3},
    "delta_bound_state_phonons": {"kind": "delta_bound_state_phonons", "xI": 0.9, "g": -1.0, "cut_off": 2.3, "k": 0.6, "A": 1.7, "beta": 1.2}
}


def structure_factor(N=24, L=4.0):
    return pbc(N, L, meas=[{"kind": "structureFactor", "label": "sq", "sets": [0, 0], "bins": 16, "maxq": 16 * 2 * np.pi / L,
                            "lBox": L, "channel": 1},
                           {"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 100, "minx": 0, "maxx": L / 2}])


def eval_cases():
    """name -> (cfg, R) for per-configuration parity (1)."""
    rng = np.random.default_rng(20261017)
    out = {}
    for N, L in ((2, 3.0), (3, 3.0), (10, 3.0), (31, 3.2), (32, 3.2), (33, 3.2), (64, 4.0), (65, 4.0), (100, 4.7)):
        out["pbc_gg_barrier_N%d" % N] = (pbc(N, L), rng.uniform(-L, L, size=(6, 3 * N)))
    out["pbc_gg_nopot_N40"] = (pbc(40, 3.5, pot=None), rng.uniform(-4, 4, size=(6, 120)))
    out["pbc_gg_squarewell_N20"] = (pbc(20, 3.0, pot="squareWell"), rng.uniform(-4, 4, size=(6, 60)))
    out["trap_N10"] = (trap(10), rng.normal(0, 0.7, size=(8, 30)))
    out["trap_N7"] = (trap(7), rng.normal(0, 0.7, size=(8, 21)))
    out["mixture"] = (mixture(), rng.uniform(-2.5, 2.5, size=(6, 96)))
    out["three_body_N16"] = (three_body(16), rng.uniform(-2.5, 2.5, size=(6, 48)))
    out["three_body_hs_N12"] = (three_body(12, hs=1.6), rng.uniform(-2.5, 2.5, size=(12, 36)))
    for name, J in JASTROWS.items():
        N, L = 18, 5.0
        R = rng.uniform(-2.5, 2.5, size=(6, 3 * N))
        if name in ("hardSphere",):
            L = 4.0
        out["jastrow_" + name] = (pbc(N, L, jastrow=J), R)
    # lattice configuration: exact minimum-image ties (t/L = +-0.5) in a power-of-two box
    N, L = 8, 4.0
    g = np.array([[x, y, z] for x in (0., 2.) for y in (0., 2.) for z in (0., 2.)]).T.reshape(-1)
    out["lattice_ties_N8"] = (pbc(N, L), np.stack([g, g + 4.0, -g]))
    # distinguishable three-body terms (appended last: the seeded stream of the cases above is unchanged)
    out["mixture3"] = (mixture3(), rng.uniform(-2.0, 2.0, size=(6, 36)))
    out["mixture3_hs"] = (mixture3(hs=1.3), rng.uniform(-2.0, 2.0, size=(12, 36)))
    out["impurity"] = (impurity(), rng.uniform(-2.0, 2.0, size=(6, 36)))
    # dipolar_rep (jastrowDipolar.h: Bessel-K branch below matching_point, sin^alpha tail up to cut_off, flat beyond)
    out["jastrow_dipolar_rep"] = (pbc(18, 5.2, jastrow=DIPOLAR_REP), rng.uniform(-2.6, 2.6, size=(6, 54)))
    # round 2 (appended last again): poschTellerPotential2b (potential.cpp:100-114) and the three-body terms at BASELINE
    # config 4's size (N=32: one full 32-particle block)
    pt = pbc(20, 3.0, pot=None)
    pt["potentials"] = [{"kind": "poschTeller2b", "sets": [0, 0], "R0": 0.8}]
    out["posch_teller2b_N20"] = (pt, rng.uniform(-3, 3, size=(6, 60)))
    out["three_body_N32"] = (three_body(32), rng.uniform(-2.5, 2.5, size=(4, 96)))
    out["three_body_hs_N32"] = (three_body(32, hs=0.9), rng.uniform(-2.5, 2.5, size=(8, 96)))
    return out


def dmc_cases():
    """Small full-DMC runs for statistical parity (2): energies / g(r) within 2 sigma of the reference's CPU run."""
    t = trap(10, walkers=400)
    t.update(timeStep=5e-3, stepsPerBlock=200, nBlocks=160, seed=1235, initialConditionGenerator={"lBox": [2.0, 2.0, 2.0]})
    p = pbc(16, 2.6, walkers=400, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 26, "minx": 0, "maxx": 1.3}])
    p.update(timeStep=5e-3, stepsPerBlock=200, nBlocks=160, seed=77)
    return {"trap_N10": t, "pbc_N16": p}


def dmc_pure_estimators_case():
    """Full DMC run for statistical parity (2) of the pure-estimator family against the reference's CPU run: forward-walking g(r)
    (ring of 12 records, 0 and 8 steps forward) and the superfluid-fraction winding estimator (12 lags)."""
    p = pbc(16, 2.6, walkers=400, meas=[
        {"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 13, "minx": 0, "maxx": 1.3, "recordSteps": 12, "forwardWalkingSteps": [0, 8]},
        {"kind": "superfluidFraction", "label": "rho", "sets": [0, 0], "bins": 12}])
    p.update(timeStep=5e-3, stepsPerBlock=100, nBlocks=140, seed=79, correlationSteps=3)
    return p


def vmc_case():
    """method "vmc" (vmcDriver.cpp): the trap of config 1 sampled from |psi|^2; the reference runs one chain"""
    t = trap(10)
    t.update(method="vmc", timeStep=0.1, stepsPerBlock=2000, nBlocks=120, correlationSteps=4, seed=31,
             initialConditionGenerator={"lBox": [2.0, 2.0, 2.0]})
    t.pop("walkers")
    t["measurements"] = [{"kind": "centerOfMassSquared", "label": "cm2", "set": 0}]
    return t


def forward_walking_cases():
    """Small systems with a stored histogram estimator ("recordSteps") and forward-walking estimators
    ("forwardWalkingSteps": [..], estimatorsFactory.cpp:47-74).  name -> cfg."""
    fw = {"recordSteps": 4, "forwardWalkingSteps": [0, 1, 3]}
    p = pbc(8, 2.2, walkers=24, meas=[dict({"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 16, "minx": 0.0, "maxx": 1.1}, **fw)])
    p.update(timeStep=2e-2, delta_walkers=3, seed=11)
    t = three_body(N=7, L=3.0)
    t.update(timeStep=2e-2, walkers=20, delta_walkers=3)
    t["measurements"] = [dict({"kind": "trimerCorrelationUnDis", "label": "trimer", "sets": [0], "bins": 12, "minx": 0.0, "maxx": 1.5},
                              recordSteps=3, forwardWalkingSteps=[2, 0])]
    m = mixture()
    m.update(N=[4, 5], lBox=[3.0, 3.0, 3.0], timeStep=1e-2, walkers=18, delta_walkers=3)
    m["measurements"] = [{"kind": "pairCorrelation", "label": "g00", "sets": [0, 0], "bins": 9, "minx": 0.1, "maxx": 1.5},
                         dict({"kind": "pairCorrelation", "label": "g01", "sets": [0, 1], "bins": 10, "minx": 0.0, "maxx": 1.5},
                              recordSteps=5, forwardWalkingSteps=[4, 2])]
    # scalar observable with a record ring (realScalarStorer / realScalarForwardWalkingEstimator,
    # correlationEstimator.cpp:5-80) next to a plain three-set histogram; appended last (seeded fixtures)
    h = impurity(N=(3, 3, 1), L=3.0)
    h.update(timeStep=2e-2, walkers=20, delta_walkers=3)
    h["measurements"] = [{"kind": "trimerCorrelation", "label": "trimer", "sets": [0, 1, 2], "bins": 8, "minx": 0.0, "maxx": 1.5},
                         {"kind": "trimerhyperRadius", "label": "hyper", "sets": [0, 1, 2], "recordSteps": 3, "forwardWalkingSteps": [0, 2]}]
    return {"pbc_pair": p, "three_body_trimer": t, "mixture_g01": m, "impurity_hyper": h}


def superfluid_cases():
    """"superfluidFraction" measurements (estimatorsFactory.cpp:80-116; centerOfMassStorer + superfluidFractionEstimator,
    correlationEstimator.cpp:187-357): the winding estimator of one species and the cross term of a mixture.  Every
    measurement uses the same bins (= the storer's recordSteps, as examples/dmc/dipolarMixture.json does)."""
    p = pbc(8, 2.2, walkers=24, meas=[{"kind": "superfluidFraction", "label": "rho00", "sets": [0, 0], "bins": 4}])
    p.update(timeStep=2e-2, delta_walkers=3, seed=13)
    m = mixture()
    m.update(N=[4, 5], lBox=[3.0, 3.0, 3.0], timeStep=1e-2, walkers=18, delta_walkers=3)
    m["measurements"] = [{"kind": "superfluidFraction", "label": "rho11", "sets": [1, 1], "bins": 5},
                         {"kind": "pairCorrelation", "label": "g01", "sets": [0, 1], "bins": 10, "minx": 0.0, "maxx": 1.5},
                         {"kind": "superfluidFraction", "label": "rho01", "sets": [0, 1], "bins": 5},
                         {"kind": "superfluidFraction", "label": "rho00", "sets": [0, 0], "bins": 5}]
    return {"pbc_rho": p, "mixture_rho": m}


def fw_measurement(cfg):
    """(index among the measurements, measurement dict) of the forward-walking measurement of a case"""
    for i, m in enumerate(cfg["measurements"]):
        if "forwardWalkingSteps" in m:
            return i, m
    raise KeyError("no forward-walking measurement")


def synthetic_genealogy(cfg, K, n0, seed):
    """K sampling events of random configurations with random cloning/killing between them"""
    import numpy as np
    rng = np.random.default_rng(seed)
    L = np.asarray(cfg["lBox"], dtype=float)
    pops, ancs = [], []
    n = n0
    for e in range(K):
        if e == 0:
            anc = np.arange(n)
        else:
            nprev = len(pops[-1])
            n = max(2, nprev + int(rng.integers(-2, 3)))
            anc = np.sort(rng.integers(0, nprev, size=n))
        cols = []
        for N in cfg["N"]:
            cols.append((rng.uniform(-0.5, 0.5, size=(n, 3, N)) * L[None, :, None]).reshape(n, -1))
        pops.append(np.concatenate(cols, axis=1))
        ancs.append(anc.astype(np.int32))
    return pops, ancs
