"""Small runs of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck).
   python scratch/sanitize_cases.py [family ...]   families: small pair ragged hist triplet sq mixture fw vmc pipeline1 rebalance"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from qmc4_b200.system import System


def run(name, cfg, R, steps=3, env=None, expect=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        s = System.from_input(cfg)
        s.upload(R)
        for _ in range(steps):
            s.step(1, 1)
        pop = s.download()
        tag = {k: s.info(k) for k in (expect or [])}
        n_est = len(cfg.get("measurements", []))
        tot = [float(np.sum(s.estimator(i)[0])) for i in range(n_est)]
        s.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert np.all(np.isfinite(pop))
    print("%-10s ok: population %d, estimator sums %s %s" % (name, pop.shape[0], ["%.6g" % t for t in tot], tag), flush=True)


def fam_small():
    cfg = cases.trap(2, walkers=96)
    run("small", cfg, np.random.default_rng(1).normal(size=(96, 6)), expect=["small_kernel"])
    cfg = cases.trap(5, walkers=40)
    run("small5", cfg, np.random.default_rng(2).normal(size=(40, 15)), expect=["small_kernel"])


def fam_pair():
    cfg = cases.pbc(64, 4.0, walkers=48)
    run("pair64", cfg, np.random.default_rng(3).uniform(-4, 4, size=(48, 192)), expect=["pair_kernel"])
    cfg = cases.pbc(160, 5.4, walkers=12)
    run("pair160", cfg, np.random.default_rng(4).uniform(-4, 4, size=(12, 480)), expect=["pair_kernel"])


def fam_ragged():
    cfg = cases.pbc(33, 3.2, walkers=24, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 64, "minx": 0.1, "maxx": 1.6}])
    run("ragged33", cfg, np.random.default_rng(5).uniform(-3, 3, size=(24, 99)), expect=["pair_kernel", "pair_hist_kernel"])
    cfg = cases.pbc(100, 4.7, walkers=10, lBox=[4.7, 5.0, 4.4])
    run("noncubic", cfg, np.random.default_rng(6).uniform(-3, 3, size=(10, 300)), expect=["pair_kernel"])


def fam_hist():
    cfg = cases.pbc(64, 4.0, walkers=48, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 1000, "minx": 0, "maxx": 2.0}])
    R = np.random.default_rng(7).uniform(-4, 4, size=(48, 192))
    run("hist_fx", cfg, R, expect=["pair_hist_kernel"])
    run("hist_gen", cfg, R, env={"QMC4_LEAN_HIST": "0"}, expect=["pair_hist_kernel"])


def fam_triplet():
    cfg = cases.three_body(32, 5.0, hs=0.2)
    cfg["walkers"] = 24
    run("triplet", cfg, np.random.default_rng(8).uniform(-2.5, 2.5, size=(24, 96)), expect=["triplet_cyclic"])
    run("triplet_g", cfg, np.random.default_rng(8).uniform(-2.5, 2.5, size=(24, 96)), env={"QMC4_TRIPLET_CYCLIC": "0"}, expect=["triplet_cyclic"])


def fam_sq():
    cfg = cases.structure_factor(24, 4.0)
    cfg["walkers"] = 24
    run("sq", cfg, np.random.default_rng(9).uniform(-2, 2, size=(24, 72)))


def fam_mixture():
    cfg = cases.mixture()
    run("mixture", cfg, np.random.default_rng(10).uniform(-2.5, 2.5, size=(16, 96)))
    cfg = cases.mixture3(hs=0.1)
    cfg["walkers"] = 12
    run("mixture3", cfg, np.random.default_rng(11).uniform(-2, 2, size=(12, 36)))


def fam_fw():
    for name, cfg in cases.forward_walking_cases().items():
        W = cfg["walkers"]
        dofs = 3 * sum(cfg["N"])
        L = cfg["lBox"][0]
        run("fw_" + name[:6], cfg, np.random.default_rng(12).uniform(-L / 2, L / 2, size=(W, dofs)), steps=6)


def fam_vmc():
    cfg = cases.vmc_case()
    cfg["walkers"] = 32
    run("vmc", cfg, np.random.default_rng(13).normal(size=(32, 30)), steps=4)


def fam_pipeline1():
    cfg = cases.pbc(64, 4.0, walkers=32)
    run("pipeline1", cfg, np.random.default_rng(14).uniform(-4, 4, size=(32, 192)), env={"QMC4_PIPELINE": "1"})
    run("general", cfg, np.random.default_rng(14).uniform(-4, 4, size=(32, 192)), env={"QMC4_LEAN": "0"})


FAMILIES = {k[4:]: v for k, v in list(globals().items()) if k.startswith("fam_")}

if __name__ == "__main__":
    for f in (sys.argv[1:] or list(FAMILIES)):
        FAMILIES[f]()
