#!/usr/bin/env python
"""Regenerates tests/golden/*.npz and *.json from the REFERENCE ITSELF (oracle/_ref, built from the
unmodified /root/reference sources by `make -C oracle ref`).  Run here (the container that has
/root/reference); the fixtures travel to the GPU box, the reference does not.

    python tests/golden/make_golden.py [--dmc]      (--dmc also re-runs the minutes-long DMC fixtures)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle.oracle import REF_BIN, Ref  # noqa: E402


def superfluid(ref):
    """superfluidFraction: the reference's centerOfMassStorer + superfluidFractionEstimator over synthetic genealogies
    (11 events over rings of 4/5 records) -> superfluid.npz;  `make_golden.py --superfluid` writes only this file"""
    out = {}
    for name, cfg in cases.superfluid_cases().items():
        pops, ancs = cases.synthetic_genealogy(cfg, 11, 6, 2000 + len(name))
        out[name + "/n"] = np.asarray([len(p_) for p_ in pops])
        out[name + "/R"] = np.concatenate(pops)
        out[name + "/anc"] = np.concatenate(ancs)
        for iM, m_ in enumerate(cfg["measurements"]):          # no forward-walking entries here: estimator index = measurement index
            if m_["kind"] == "superfluidFraction":
                sums, ws = ref.superfluid_fraction(cfg, pops, ancs, iM, m_["bins"])
                out["%s/%s/sums" % (name, m_["label"])] = sums
                out["%s/%s/weights" % (name, m_["label"])] = ws
    np.savez_compressed(os.path.join(HERE, "superfluid.npz"), **out)


def force_energy(ref):
    """forceEnergy::operator() (energy.cpp:37-61) of the reference on the evaluation cases -> force_energy.npz"""
    out = {}
    for name in ("pbc_gg_barrier_N33", "mixture", "trap_N10", "three_body"):
        if name not in cases.eval_cases():
            continue
        cfg, R = cases.eval_cases()[name]
        R = R[ref.eval(cfg, R)["valid"] == 1]
        out[name + "/R"] = R
        out[name + "/forceEnergy"] = ref.force_energy(cfg, R)
    np.savez_compressed(os.path.join(HERE, "force_energy.npz"), **out)


def parse_blocks(path):
    """per-block value columns of a histogram/vector estimator file (header line, blocks separated by blank lines)"""
    txt = open(path).read()
    txt = txt[txt.index("\n") + 1:]
    out = []
    for blk in txt.split("\n\n"):
        rows = [r.split() for r in blk.strip().split("\n") if r.strip()]
        if rows:
            out.append([float(r[1]) for r in rows])
    return out


def dmc_pure(ref):
    """the reference's own binary on cases.dmc_pure_estimators_case() -> dmc_reference_pure_run.json (about a minute)"""
    cfg = cases.dmc_pure_estimators_case()
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "input.json"), "w") as f:
            json.dump(cfg, f)
        subprocess.check_call([REF_BIN, "input.json"], cwd=tmp, stdout=subprocess.DEVNULL)
        entry = {"input": cfg, "energy_per_block": np.loadtxt(os.path.join(tmp, "energy.dat"), skiprows=1).tolist()}
        for name in ("pair_fw0", "pair_fw8", "rho"):
            entry[name + "_per_block"] = parse_blocks(os.path.join(tmp, name + ".dat"))
    with open(os.path.join(HERE, "dmc_reference_pure_run.json"), "w") as f:
        json.dump(entry, f)


def main():
    ref = Ref()
    if "--dmc-pure" in sys.argv:
        dmc_pure(ref)
        return
    if "--superfluid" in sys.argv or "--force-energy" in sys.argv:
        if "--superfluid" in sys.argv:
            superfluid(ref)
        if "--force-energy" in sys.argv:
            force_energy(ref)
        return
    if "--append-eval" in sys.argv:
        # add the cases named after the flag to eval_cases.npz without touching the fixtures that are already pinned
        names = sys.argv[sys.argv.index("--append-eval") + 1:]
        old = dict(np.load(os.path.join(HERE, "eval_cases.npz")))
        for name in names:
            cfg, R = cases.eval_cases()[name]
            r = ref.eval(cfg, R)
            old[name + "/R"] = R
            for k, v in r.items():
                old[name + "/" + k] = v
            print(name, {k: np.asarray(v).shape for k, v in r.items()})
        np.savez_compressed(os.path.join(HERE, "eval_cases.npz"), **old)
        return
    rng = np.random.default_rng(424242)
    out = {}
    for name, (cfg, R) in cases.eval_cases().items():
        r = ref.eval(cfg, R)
        out[name + "/R"] = R
        for k, v in r.items():
            out[name + "/" + k] = v
    np.savez_compressed(os.path.join(HERE, "eval_cases.npz"), **out)

    pieces = {}
    # drift-diffusion + Green ratio + histograms on fixed inputs
    for name in ("pbc_gg_barrier_N33", "mixture", "trap_N10"):
        cfg, R = cases.eval_cases()[name]
        e0 = ref.eval(cfg, R)
        xi = rng.normal(size=R.shape)
        Rn = ref.drift_diffuse(cfg, cfg["timeStep"], R, e0["grad"], xi)
        e1 = ref.eval(cfg, Rn)
        q = ref.green_ratio(cfg, cfg["timeStep"], Rn, e1["grad"], R, e0["grad"])
        pieces[name + "/xi"] = xi
        pieces[name + "/Rnew"] = Rn
        pieces[name + "/q"] = q
    hist_cases = {"mixture": (cases.mixture(), 2), "three_body_N16": (cases.three_body(16), 1),
                  "pair_N33": (cases.pbc(33, 3.2, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 64, "minx": 0, "maxx": 1.6}]), 1)}
    for name, (cfg, nest) in hist_cases.items():
        R = rng.uniform(-cfg["lBox"][0] / 2, cfg["lBox"][0] / 2, size=(10, 3 * sum(cfg["N"])))
        pieces["hist_" + name + "/R"] = R
        for i in range(nest):
            bins = cfg["measurements"][i]["bins"]
            s, w = ref.histogram(cfg, R, i, bins)
            pieces["hist_%s/sums%d" % (name, i)] = s
            pieces["hist_%s/weight%d" % (name, i)] = np.asarray([w])
    # branching with the reference's own RNG (std::ranlux48) and its in-place algorithm
    for tag, (mean, delta) in {"lower": (40, 10), "upper": (800, 10), "window": (100, 10)}.items():
        Eo = rng.normal(3.0, 40.0, size=100)
        En = Eo + rng.normal(0, 5.0, size=100)
        b = ref.branch(1e-3, mean, delta, 10, En, Eo)
        pieces["branch_%s/Enew" % tag] = En
        pieces["branch_%s/Eold" % tag] = Eo
        pieces["branch_%s/cfg" % tag] = np.asarray([1e-3, mean, delta])
        for k in ("nDesc", "u", "Eafter"):
            pieces["branch_%s/%s" % (tag, k)] = b[k]
        pieces["branch_%s/shifts" % tag] = np.asarray([b["energyShift"], b["energyShiftAfter"]])
    # planner
    pops = np.stack([rng.integers(0, 60, size=8) for _ in range(40)])
    srcs, amts = zip(*[ref.plan(p) for p in pops])
    pieces["plan/pops"], pieces["plan/sources"], pieces["plan/amounts"] = pops, np.stack(srcs), np.stack(amts)
    # three-set histogram observables and scalar observables (own stream: the fixtures above do not move)
    rng2 = np.random.default_rng(515151)
    for name, cfg in cases.three_set_hist_cases().items():
        R = rng2.uniform(-cfg["lBox"][0] / 2, cfg["lBox"][0] / 2, size=(10, 3 * sum(cfg["N"])))
        pieces["hist_" + name + "/R"] = R
        for i, m_ in enumerate(cfg["measurements"]):
            s, w = ref.histogram(cfg, R, i, cases.meas_bins(m_))
            pieces["hist_%s/sums%d" % (name, i)] = s
            pieces["hist_%s/weight%d" % (name, i)] = np.asarray([w])
    np.savez_compressed(os.path.join(HERE, "pieces.npz"), **pieces)

    # forward walking: the reference's storer + forward-walking estimator classes over synthetic genealogies
    fwd = {}
    for name, cfg in cases.forward_walking_cases().items():
        iM, meas = cases.fw_measurement(cfg)
        pops, ancs = cases.synthetic_genealogy(cfg, 9, 6, 1000 + len(name))
        fwd[name + "/n"] = np.asarray([len(p_) for p_ in pops])
        fwd[name + "/R"] = np.concatenate(pops)
        fwd[name + "/anc"] = np.concatenate(ancs)
        # index in the vector estimatorFactory::create returns: plain estimators before it count one each
        first = sum(len(m_.get("forwardWalkingSteps", [0])) if "forwardWalkingSteps" in m_ else 1 for m_ in cfg["measurements"][:iM])
        for k, steps in enumerate(meas["forwardWalkingSteps"]):
            sums, w = ref.forward_walking(cfg, pops, ancs, first + k, cases.meas_bins(meas))
            fwd["%s/fw%d/sums" % (name, steps)] = sums
            fwd["%s/fw%d/weight" % (name, steps)] = np.asarray(w)
    np.savez_compressed(os.path.join(HERE, "forward_walking.npz"), **fwd)

    superfluid(ref)
    force_energy(ref)

    if "--dmc" in sys.argv:
        res = {}
        for name, cfg in cases.dmc_cases().items():
            with tempfile.TemporaryDirectory() as tmp:
                with open(os.path.join(tmp, "input.json"), "w") as f:
                    json.dump(cfg, f)
                subprocess.check_call([REF_BIN, "input.json"], cwd=tmp, stdout=subprocess.DEVNULL)
                e = np.loadtxt(os.path.join(tmp, "energy.dat"), skiprows=1)
                entry = {"input": cfg, "energy_per_block": e.tolist()}
                for m in cfg["measurements"]:
                    blocks = open(os.path.join(tmp, m["label"] + ".dat")).read().strip().split("\n\n")
                    g = []
                    for i, b in enumerate(blocks):
                        rows = [l.split() for l in b.strip().split("\n")]
                        rows = [r for r in rows if len(r) == 2 and r[0] != "x"]
                        g.append([float(r[1]) for r in rows])
                    entry[m["label"] + "_per_block"] = g
                res[name] = entry
        with open(os.path.join(HERE, "dmc_reference_runs.json"), "w") as f:
            json.dump(res, f)
    if "--dmc" in sys.argv:
        dmc_pure(ref)
    if "--vmc" in sys.argv or "--dmc" in sys.argv:
        cfg = cases.vmc_case()
        with tempfile.TemporaryDirectory() as tmp:
            with open(os.path.join(tmp, "input.json"), "w") as f:
                json.dump(cfg, f)
            subprocess.check_call([REF_BIN, "input.json"], cwd=tmp, stdout=subprocess.DEVNULL)
            entry = {"input": cfg, "energy_per_block": np.loadtxt(os.path.join(tmp, "energy.dat"), skiprows=1).tolist(),
                     "cm2_per_block": np.loadtxt(os.path.join(tmp, "cm2.dat"), skiprows=1).tolist()}
        with open(os.path.join(HERE, "vmc_reference_run.json"), "w") as f:
            json.dump(entry, f)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
