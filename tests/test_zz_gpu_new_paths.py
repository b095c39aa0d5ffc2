"""GPU parity tests (-m gpu) of the paths that were written while no B200 time was left in round 1:
forward walking and the §8(f)2-4 widening.  Their logic is parity-green on the thread
emulation of the CUDA sources (tests/test_emulated_kernels.py runs this file against tests/emu), but they
have not executed on hardware yet — the file sorts last so that, under `pytest -x`, a hardware-only failure
here cannot hide the suite that was validated on the B200 (test_gpu_parity.py, test_host_driver.py).
Tolerance for floating point: 1e-10 relative (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import cases
import test_host_driver as hd
from qmc4_b200.system import System

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# ---- forward walking (kernels_fw.cu) ---------------------------------------------------
def _traced_genealogy(s, cfg, events, sample_every, rng):
    """run events*sample_every traced DMC steps; returns the post-branching populations at the sampling
    events and, for each, the index of every walker's ancestor in the previous event's population"""
    W = cfg["walkers"]
    L = np.asarray(cfg["lBox"], dtype=float)
    cols = [(rng.uniform(-0.5, 0.5, size=(W, 3, N)) * L[None, :, None]).reshape(W, -1) for N in cfg["N"]]
    s.upload(np.concatenate(cols, axis=1))
    pops, ancs, comp = [], [], None
    for t in range(events * sample_every):
        sample = (t + 1) % sample_every == 0
        tr = s.step_traced(accumulate=sample)
        step_anc = np.repeat(np.arange(tr["n_before"]), tr["n_desc"][:tr["n_before"]])     # branching.cpp:34-61 as a map
        assert len(step_anc) == tr["n_after"]
        comp = step_anc if comp is None else comp[step_anc]
        if sample:
            R, _ = s.download(with_energy=True)
            pops.append(R)
            ancs.append(comp.astype(np.int32) if len(pops) > 1 else np.arange(len(R), dtype=np.int32))
            comp = None
    return pops, ancs


@pytest.mark.parametrize("sample_every", [1, 3])
@pytest.mark.parametrize("name", sorted(cases.forward_walking_cases()))
def test_forward_walking_vs_oracle_on_the_traced_genealogy(port, name, sample_every):
    """the device never copies a record: it must still reproduce the reference's per-walker rings that are
    deep-copied at every branching, for the genealogy the run actually produced"""
    cfg = cases.forward_walking_cases()[name]
    iM, meas = cases.fw_measurement(cfg)
    s = System.from_input(cfg)
    assert s.info("fw_depth") == max(meas["forwardWalkingSteps"]) + 1
    pops, ancs = _traced_genealogy(s, cfg, 9, sample_every, np.random.default_rng(3))
    assert len({len(p) for p in pops}) > 1, "no branching happened: the test would not exercise the genealogy"
    labels = [e["label"] for e in s.estimators]
    for steps in meas["forwardWalkingSteps"]:
        e = s.estimators[labels.index("%s_fw%d" % (meas["label"], steps))]     # estimatorsFactory.cpp:58
        got, w = s.estimator(e["index"])
        exp, we = port.forward_walking(cfg, pops, ancs, iM, meas["recordSteps"], steps, cases.meas_bins(meas))
        assert w == we and w > 0
        assert rel(got, exp) < RTOL, steps
    assert meas["label"] not in labels                 # the plain estimator is replaced, not added (estimatorsFactory.cpp:47)
    s.clear_estimators()
    got, w = s.estimator(s.estimators[0]["index"])
    assert w == 0 and not got.any()
    s.close()


def test_forward_walking_ring_restarts_with_a_new_population(port):
    cfg = cases.forward_walking_cases()["pbc_pair"]
    iM, meas = cases.fw_measurement(cfg)
    s = System.from_input(cfg)
    rng = np.random.default_rng(8)
    _traced_genealogy(s, cfg, 5, 1, rng)
    s.clear_estimators()
    pops, ancs = _traced_genealogy(s, cfg, 6, 1, rng)          # upload: storer::reserve, rings filling again
    assert s.info("fw_events") == 6
    e = s.estimators[0]
    got, w = s.estimator(e["index"])
    exp, we = port.forward_walking(cfg, pops, ancs, iM, meas["recordSteps"], e["forwardWalkingSteps"], cases.meas_bins(meas))
    assert w == we and rel(got, exp) < RTOL
    s.close()


# ---- superfluidFraction (centerOfMassStorer + superfluidFractionEstimator on the ancestor links) ----------------
@pytest.mark.parametrize("sample_every", [1, 2])
@pytest.mark.parametrize("name", sorted(cases.superfluid_cases()))
def test_superfluid_fraction_vs_oracle_on_the_traced_genealogy(port, name, sample_every):
    """per-slot centre-of-mass records + link hops against the per-walker rings the reference copies at every branching;
    11 events over rings of 4/5 records: the ring turns over twice, so the restart of the lags at record 0 is covered"""
    cfg = cases.superfluid_cases()[name]
    s = System.from_input(cfg)
    bins = [m["bins"] for m in cfg["measurements"] if m["kind"] == "superfluidFraction"][0]
    assert s.info("fw_link_depth") == bins
    pops, ancs = _traced_genealogy(s, cfg, 11, sample_every, np.random.default_rng(5))
    assert len({len(p) for p in pops}) > 1, "no branching happened: the test would not exercise the genealogy"
    by_label = {e["label"]: e for e in s.estimators}
    for iM, m in enumerate(cfg["measurements"]):
        if m["kind"] != "superfluidFraction":
            continue
        got, w = s.estimator(by_label[m["label"]]["index"])
        exp, we = port.superfluid_fraction(cfg, pops, ancs, iM)
        np.testing.assert_array_equal(w, we)
        assert w[0] > w[-1] > 0
        assert rel(got, exp) < RTOL, m["label"]
    s.clear_estimators()
    got, w = s.estimator(by_label[[m for m in cfg["measurements"] if m["kind"] == "superfluidFraction"][0]["label"]]["index"])
    assert not w.any() and not got.any()
    s.close()


def test_superfluid_fraction_without_branching_and_errors(port):
    """svmc (branching off, main.cpp:393-395): identity genealogy; the record clock restarts with a new population;
    vmc + superfluidFraction is invalidInput in the reference (main.cpp:279-282); bins must agree (the storer's recordSteps)"""
    cfg = dict(cases.superfluid_cases()["pbc_rho"], method="svmc")
    s = System.from_input(cfg)
    rng = np.random.default_rng(2)
    _traced_genealogy(s, cfg, 3, 1, rng)
    s.clear_estimators()
    pops, ancs = _traced_genealogy(s, cfg, 6, 1, rng)
    assert len({len(p) for p in pops}) == 1
    got, w = s.estimator(s.estimators[0]["index"])
    exp, we = port.superfluid_fraction(cfg, pops, ancs, 0)
    np.testing.assert_array_equal(w, we)
    assert rel(got, exp) < RTOL
    s.close()
    with pytest.raises(Exception, match="vmc and superfluid"):
        System.from_input(dict(cfg, method="vmc"))
    bad = cases.superfluid_cases()["mixture_rho"]
    bad["measurements"][2]["bins"] = 7
    with pytest.raises(Exception, match="same .bins"):
        System.from_input(bad)
    bad = cases.superfluid_cases()["pbc_rho"]
    bad["measurements"][0]["sets"] = [0]
    with pytest.raises(Exception, match="two sets"):
        System.from_input(bad)


def test_pure_estimators_are_bitwise_reproducible():
    """fixed-order reductions everywhere (column sums over slot chunks, CTA tree + chunk order for the winding products):
    two runs of the same input give identical bits, the property DESIGN.md §3 claims for the physics path"""
    import copy
    cfg = copy.deepcopy(cases.forward_walking_cases()["pbc_pair"])
    cfg["measurements"].append({"kind": "superfluidFraction", "label": "rho", "sets": [0, 0], "bins": 4})
    L = np.asarray(cfg["lBox"])
    R = (np.random.default_rng(4).uniform(-0.5, 0.5, size=(cfg["walkers"], 3, cfg["N"][0])) * L[None, :, None]).reshape(cfg["walkers"], -1)
    runs = []
    for _ in range(2):
        s = System.from_input(cfg)
        s.upload(R)
        s.step(14, 2)
        runs.append({e["label"]: s.estimator(e["index"]) for e in s.estimators})
        s.close()
    for label in runs[0]:
        (a, wa), (b, wb) = runs[0][label], runs[1][label]
        assert np.array_equal(a, b) and np.array_equal(wa, wb), label
        assert np.asarray(wa).sum() > 0


def test_superfluid_fraction_measurement_file(tmp_path):
    """<label>.dat of the host driver (superfluidFractionEstimator::write, correlationEstimator.cpp:305-320): header
    "x <label>", per block the rows "lag average" of the lags that received data.  Physics check: the pair Jastrow's
    drift sums to zero over the particles, so the centre of mass diffuses freely, W(lag) = <(sum_i dx_i)^2> = 3 N tau lag
    x acceptance (superfluid fraction 1 for this translation-invariant gas)."""
    cfg = cases.superfluid_cases()["pbc_rho"]
    cfg = dict(cfg, stepsPerBlock=8, nBlocks=3, correlationSteps=0, walkers=40, delta_walkers=6)
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    m = cfg["measurements"][0]
    txt = open(tmp_path / (m["label"] + ".dat")).read()
    lines = txt.split("\n")
    assert lines[0] == "x " + m["label"]
    blocks = [[r for r in b.split("\n") if r.strip()] for b in txt[len(lines[0]) + 1:].split("\n\n")]
    blocks = [b for b in blocks if b]
    assert len(blocks) == 3 and all(len(b) == m["bins"] for b in blocks)
    vals = np.array([[[float(v) for v in r.split()] for r in b] for b in blocks])
    np.testing.assert_array_equal(vals[:, :, 0], np.tile(np.arange(m["bins"]), (3, 1)))
    assert (vals[:, 0, 1] == 0).all()                                         # lag 0: a record against itself
    W = vals[:, :, 1].mean(axis=0)
    free = 3 * cfg["N"][0] * cfg["timeStep"] * np.arange(m["bins"])
    assert np.all(np.abs(W[1:] / free[1:] - 1) < 0.35), (W, free)


def test_forward_walking_records_migrate_between_two_gpus():
    """the multi-rank record migration of qmc_rebalance on real GPUs + NCCL (tests/emu/multirank_fw_driver.py, one host
    thread per GPU): needs 2 GPUs, skipped on the 1-GPU box; its logic runs on emulated ranks in the CPU suite"""
    import subprocess
    import sys
    import torch
    if hd.BIN.endswith("_emu") or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "emu", "multirank_fw_driver.py"), "2", "1"],
                       env=dict(os.environ, QMC4_FW_DRIVER_REAL_GPUS="1"), cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(hd.BIN.endswith("_emu"), reason="5.6e6 walker-moves: hardware only")
def test_pure_estimators_statistical_parity_with_reference_run(tmp_path):
    """parity (2) for the pure-estimator family: forward-walking g(r) (0 and 8 sampling steps forward) and the superfluid-
    fraction winding estimator of a full DMC run against the reference's own CPU run of the same input
    (tests/golden/dmc_reference_pure_run.json, make_golden.py --dmc-pure), bin by bin within 2 sigma of the blocked errors"""
    import json
    fx = json.load(open(os.path.join(hd.GOLD, "dmc_reference_pure_run.json")))
    cfg = fx["input"]
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    skip = 20                                             # equilibration blocks

    def blocks(name):
        txt = open(tmp_path / (name + ".dat")).read()
        txt = txt[txt.index("\n") + 1:]
        out = [[float(r.split()[1]) for r in b.strip().split("\n") if r.strip()] for b in txt.split("\n\n")]
        return np.array([o for o in out if o])

    e = np.loadtxt(tmp_path / "energy.dat", skiprows=1)[skip:]
    e_ref = np.asarray(fx["energy_per_block"])[skip:]
    assert abs(e.mean() - e_ref.mean()) < 2.5 * np.hypot(hd._blocked_error(e), hd._blocked_error(e_ref))
    for name in ("pair_fw0", "pair_fw8", "rho"):
        got, ref = blocks(name), np.asarray(fx[name + "_per_block"])
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        got, ref = got[skip:], ref[skip:]
        cols = [i for i in range(ref.shape[1]) if ref[:, i].std() > 0]          # lag 0 of rho is identically 0
        sg = np.hypot([hd._blocked_error(got[:, i]) for i in cols], [hd._blocked_error(ref[:, i]) for i in cols])
        z = np.abs(got[:, cols].mean(0) - ref[:, cols].mean(0)) / np.maximum(sg, 1e-12)
        assert (z < 2).mean() > 0.7 and z.max() < 4.5, (name, z)
    rho = blocks("rho")[skip:].mean(0)
    free = 3 * cfg["N"][0] * cfg["timeStep"] * (cfg["correlationSteps"] + 1) * np.arange(len(rho))
    assert np.all(np.abs(rho[1:] / free[1:] - 1) < 0.1)                         # superfluid fraction ~ acceptance ~ 0.96


def test_forward_walking_errors():
    cfg = cases.structure_factor()
    cfg["measurements"][0].update(recordSteps=4, forwardWalkingSteps=[1])
    with pytest.raises(Exception, match="Forward walking estimator for vector"):      # estimatorsFactory.cpp:66-69
        System.from_input(cfg)
    cfg = cases.forward_walking_cases()["pbc_pair"]
    cfg["measurements"][0]["forwardWalkingSteps"] = [4]                                # assert, correlationEstimator.cpp:137
    with pytest.raises(Exception, match="recordSteps"):
        System.from_input(cfg)


def test_forward_walking_measurement_files(tmp_path):
    """"recordSteps" + "forwardWalkingSteps": [..] on a measurement: one <label>_fw<steps>.dat per entry instead of
    <label>.dat (estimatorsFactory.cpp:47-74), rows at the bin CENTRES (accumulators.h:146-156), nothing but the
    blank separator while the walkers' rings are filling (correlationEstimator.cpp:171-178)"""
    cfg = cases.forward_walking_cases()["pbc_pair"]
    cfg = dict(cfg, stepsPerBlock=3, nBlocks=3, correlationSteps=1)
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    m = cfg["measurements"][0]
    assert not os.path.exists(tmp_path / (m["label"] + ".dat"))
    dx = (m["maxx"] - m["minx"]) / m["bins"]
    for steps in m["forwardWalkingSteps"]:
        txt = open(tmp_path / ("%s_fw%d.dat" % (m["label"], steps))).read()
        lines = txt.split("\n")
        assert lines[0] == "x %s_fw%d" % (m["label"], steps)
        blocks = [b for b in txt[len(lines[0]) + 1:].split("\n\n")]
        # block 0: 3 sampling events < recordSteps = 4 -> no rows; blocks 1 and 2: bins rows each
        rows = [[r for r in b.split("\n") if r.strip()] for b in blocks]
        rows = [r for r in rows if r]
        assert len(rows) == 2 and all(len(r) == m["bins"] for r in rows), [len(r) for r in rows]
        first = np.array([[float(v) for v in r.split()] for r in rows[0]])
        np.testing.assert_allclose(first[:, 0], m["minx"] + (np.arange(m["bins"]) + 0.5) * dx, rtol=1e-5)
        assert np.isfinite(first[:, 1]).all() and first[:, 1].sum() > 0
    # the structure factor has no forward-walking estimator in the reference either
    bad = cases.structure_factor()
    bad["measurements"][0].update(recordSteps=4, forwardWalkingSteps=[1])
    p = hd.run(bad, str(tmp_path))
    assert p.returncode != 0 and "Forward walking estimator for vector" in p.stderr


# ---- distinguishable three-body terms, three-set observables, scalar observables ---------
import test_gpu_parity as gp  # noqa: E402

PIECES = gp.PIECES


@pytest.mark.parametrize("name", cases.NEW_PATH_EVAL_CASES)
def test_distinguishable_three_body_eval_vs_golden_and_oracle(port, name):
    """jastrow3bDis + potentialThreeBodyDis + hardSphereConstraintThreeBodyDis: per-configuration parity (1)
    against the reference's golden vectors and the oracle"""
    gp.test_eval_vs_golden_and_oracle(port, name)


@pytest.mark.parametrize("wpw", ["1", "2", "4", "8"])
def test_distinguishable_three_body_independent_of_work_decomposition(port, wpw, monkeypatch):
    monkeypatch.setenv("QMC4_WPW", wpw)
    cfg, R = cases.eval_cases()["mixture3"]
    s = System.from_input(cfg)
    assert s.info("warps_per_walker") == int(wpw)
    o = s.eval(R)
    s.close()
    r = port.eval(cfg, R)
    for k in ("logpsi", "grad", "lap", "V", "E"):
        assert rel(o[k], r[k]) < RTOL, k


@pytest.mark.parametrize("name", ["mixture3", "impurity"])
def test_distinguishable_three_body_traced_step(port, name):
    gp.test_traced_step_vs_oracle_update(port, name)


def test_distinguishable_constraint_rejects_and_invalid_start_throws(port):
    cfg, R = cases.eval_cases()["mixture3_hs"]
    valid = gp.EVAL["mixture3_hs/valid"].astype(bool)
    cfg = dict(cfg, walkers=int(valid.sum()), timeStep=0.05)
    s = System.from_input(cfg)
    with pytest.raises(Exception, match="Inalid initial"):
        s.upload(R)                              # dmcDriver.cpp:166-170
    s.close()
    s = System.from_input(cfg)
    s.upload(R[valid])
    tr = s.step_traced()
    ep = port.eval(cfg, tr["Rprop"])
    np.testing.assert_array_equal(tr["valid"], ep["valid"])
    assert not (tr["accepted"][tr["valid"] == 0]).any()
    s.close()


@pytest.mark.parametrize("name", sorted(cases.three_set_hist_cases()))
def test_three_set_and_scalar_observables_on_given_configurations(port, name):
    """trimerCorrelation, trimerCorrelationDis (histograms), trimerhyperRadius, centerOfMassSquared (scalars:
    sum of the observable and number of walkers) against the fixtures produced by the reference"""
    cfg = cases.three_set_hist_cases()[name]
    R = PIECES["hist_" + name + "/R"]
    s = System.from_input(cfg)
    s.accumulate_configs(R)
    for i, m in enumerate(cfg["measurements"]):
        sums, w = s.estimator(s.estimators[i]["index"])
        assert w == PIECES["hist_%s/weight%d" % (name, i)][0]
        assert rel(sums, PIECES["hist_%s/sums%d" % (name, i)]) < RTOL, m["kind"]
    s.clear_estimators()
    sums, w = s.estimator(s.estimators[0]["index"])
    assert w == 0 and not sums.any()
    s.close()


@pytest.mark.parametrize("name", sorted(cases.three_set_hist_cases()))
def test_three_set_and_scalar_observables_in_the_step_use_branching_weights(port, name):
    """accumulated inside the step == oracle over the post-branching population (dmcDriver.cpp:257-264)"""
    cfg = dict(cases.three_set_hist_cases()[name], walkers=48, timeStep=0.05)
    R = np.random.default_rng(31).uniform(-2, 2, size=(48, 3 * sum(cfg["N"])))
    s = System.from_input(cfg)
    s.upload(R)
    tr = s.step_traced(accumulate=True)
    assert (tr["n_desc"] != 1).any()
    Rn = s.download()
    for i, m in enumerate(cfg["measurements"]):
        sums, w = s.estimator(s.estimators[i]["index"])
        ref_s, ref_w = port.histogram(cfg, Rn, i, cases.meas_bins(m))
        assert w == ref_w == len(Rn)
        assert rel(sums, ref_s) < 1e-9, m["kind"]
    s.close()


def test_three_set_observable_errors():
    cfg = cases.impurity(N=(4, 3, 2))                 # set C must hold one particle (assert, trimerCorrelation.cpp:11)
    with pytest.raises(Exception, match="one particle"):
        System.from_input(cfg)
    cfg = cases.mixture3()
    cfg["wavefunctions"][2]["sets"] = [0, 1, 1]       # jastrowWavefunctionThreeBody.h:173-176
    with pytest.raises(Exception, match="should be distinguishible"):
        System.from_input(cfg)


def test_host_driver_three_set_and_scalar_measurement_files(tmp_path):
    """dmc_b200 with the reference's input for an impurity system: <label>.dat of a scalar observable is the
    label followed by one average per block in %.12e (realScalarEstimator::write, estimators.cpp:89-98);
    trimerCorrelation is a histogram file; a scalar forward-walking estimator writes <label>_fw<k>.dat"""
    cfg = dict(cases.impurity(), walkers=16, delta_walkers=2, nBlocks=3, stepsPerBlock=4)
    cfg["measurements"].append({"kind": "trimerhyperRadius", "label": "hfw", "sets": [0, 1, 2], "recordSteps": 3, "forwardWalkingSteps": [1]})
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    for label in ("hyper", "cm2"):
        lines = open(tmp_path / (label + ".dat")).read().split("\n")
        vals = [x for x in lines[1:] if x.strip()]
        assert lines[0] == label and len(vals) == 3
        assert all(len(v.split("e")[0].split(".")[1]) == 12 and float(v) > 0 for v in vals)
    blocks = open(tmp_path / "trimer.dat").read().strip().split("\n\n")
    assert blocks[0].split("\n")[0] == "x trimer" and len(blocks) == 3
    fw = open(tmp_path / "hfw_fw1.dat").read().split("\n")
    assert fw[0] == "hfw_fw1" and len([x for x in fw[1:] if x.strip()]) == 3 - 0       # 4 sampling events per block >= recordSteps
    assert not os.path.exists(tmp_path / "hfw.dat")
    # a three-species input with jastrow3bDis + potentialThreeBodyDis runs end to end
    cfg = dict(cases.mixture3(), walkers=12, delta_walkers=2, nBlocks=2, stepsPerBlock=3)
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert os.path.exists(tmp_path / "trimerDis.dat")
    bad = cases.mixture3()
    bad["potentials"][0]["sets"] = [0, 0, 2]
    p = hd.run(bad, str(tmp_path))
    assert p.returncode != 0 and "should be distinguishible" in p.stderr


# ---- many-chain VMC (method "vmc": vmcDriver.cpp) -----------------------------------------
@pytest.mark.parametrize("name,sigma", [("trap_N10", 0.3), ("mixture", 0.02), ("pbc_gg_barrier_N33", 0.05), ("mixture3_hs", 0.05)])
def test_vmc_traced_step_vs_oracle(port, name, sigma):
    """one vmcDriver::step per chain: Gaussian move without drift, acceptance on 2*(log psi' - log psi) only,
    rejected / invalid proposals keep the old configuration, no branching — against the oracle's restatement
    given the random numbers the kernel drew"""
    cfg, R = cases.eval_cases()[name]
    valid = port.eval(cfg, R)["valid"].astype(bool)
    R = R[valid]
    cfg = dict(cfg, method="vmc", walkers=len(R), timeStep=sigma)
    s = System.from_input(cfg)
    s.upload(R)
    tr = s.step_traced(accumulate=True)
    assert rel(tr["Rprop"], R + tr["xi"] * np.sqrt(sigma)) < 1e-14            # vmcMoves.cpp:31
    up = port.vmc_step(cfg, sigma, R, tr["xi"], tr["u_metropolis"])
    np.testing.assert_array_equal(up["accepted"], tr["accepted"])
    assert (tr["n_desc"] == 1).all() and tr["n_after"] == len(R)
    Rn, En = s.download(with_energy=True)
    assert rel(Rn, up["R"]) < 1e-13 and rel(En, up["E"]) < RTOL
    b = s.block_results()
    assert b["n_moves"] == len(R) and b["n_accepted"] == up["accepted"].sum() and b["energy_count"] == len(R)
    assert abs(b["energy_sum"] - up["E"].sum()) < 1e-9 * np.abs(up["E"]).sum()
    s.close()


def test_vmc_rejections_happen_and_errors():
    cfg, R = cases.eval_cases()["pbc_gg_barrier_N33"]
    cfg = dict(cfg, method="vmc", walkers=len(R), timeStep=0.5)
    s = System.from_input(cfg)
    s.upload(R)
    s.step(20, 1)
    b = s.block_results()
    assert 0 < b["n_accepted"] < b["n_moves"] == 20 * len(R) and b["population"] == len(R)
    s.close()
    fw = cases.forward_walking_cases()["pbc_pair"]
    with pytest.raises(Exception, match="forward walking needs branching"):
        System.from_input(dict(fw, method="vmc"))


def test_host_driver_vmc_block_report_and_files(tmp_path):
    """method "vmc": vmcDriver::out's block report (Block / Acc. Ratio / Energy, no Step or Curr. Walkers lines,
    vmcDriver.cpp:96-121), energy.dat, scalar measurement file, checkpoint of the chains"""
    import re
    cfg = dict(cases.vmc_case(), walkers=8, stepsPerBlock=10, nBlocks=3, correlationSteps=1)
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    out = re.sub(r"\x1b\[[0-9;]*m", "", p.stdout)
    assert out.count("Block: ") == 3 and out.count("Acc. Ratio: ") == 3 and out.count("Energy: ") == 3
    assert "Curr. Walkers" not in out and "Step: " not in out
    e = np.loadtxt(tmp_path / "energy.dat", skiprows=1)
    c = np.loadtxt(tmp_path / "cm2.dat", skiprows=1)
    assert len(e) == 3 and len(c) == 3 and np.isfinite(e).all() and (c > 0).all()
    import json
    ck = json.load(open(tmp_path / "configurations" / "walkers-Rank0.json"))
    assert len(ck["configurations"]) == 8
    # without "walkers": one chain per GPU, as the reference runs it
    cfg.pop("walkers")
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert len(json.load(open(tmp_path / "configurations" / "walkers-Rank0.json"))["configurations"]) == 1


# ---- forceEnergy (energy.cpp:37-61; attached for vmc and svmc, main.cpp:318-335,395-408) ----------------------
@pytest.mark.parametrize("method", ["svmc", "vmc"])
@pytest.mark.parametrize("name", ["pbc_gg_barrier_N33", "mixture", "trap_N10"])
def test_force_energy_vs_oracle_on_traced_steps(port, name, method):
    """0.5 sum |grad log psi|^2 + V of every walker after every sampling step (accepted or carried-through), against the
    oracle's forceEnergy::operator() on the downloaded populations; sampling every 2nd step"""
    cfg, R = cases.eval_cases()[name]
    R = R[port.eval(cfg, R)["valid"] == 1][:12]
    cfg = dict(cfg, method=method, walkers=len(R), timeStep=0.05 if method == "vmc" else 5e-3,
               measurements=list(cfg.get("measurements", [])) + [{"kind": "forceEnergy", "label": "forceEnergy"}])
    s = System.from_input(cfg)
    s.upload(R)
    exp_sum, exp_n, rejected = 0.0, 0, 0
    for t in range(6):
        sample = t % 2 == 1
        tr = s.step_traced(accumulate=sample)
        rejected += int((tr["accepted"][:len(R)] == 0).sum())
        if sample:
            Rnow, _ = s.download(with_energy=True)
            assert len(Rnow) == len(R)
            exp_sum += port.force_energy(cfg, Rnow).sum()
            exp_n += len(Rnow)
    got, n = s.force_energy()
    assert n == exp_n and abs(got - exp_sum) < RTOL * abs(exp_sum)
    s.clear_estimators()
    assert s.force_energy() == (0.0, 0.0)
    s.close()
    # method dmc: the reference never attaches the estimator (main.cpp:395-408) -> the measurement is ignored; asking the
    # library for it with branching on is an error
    s = System.from_input(dict(cfg, method="dmc"))
    assert not s.force_energy_enabled
    with pytest.raises(Exception, match="not enabled"):
        s.force_energy()
    s.close()
    s = System(cfg["lBox"], cfg["N"])
    for jw in cfg["wavefunctions"]:
        s.add_wavefunction(jw)
    s._ck(s.lib.qmc_enable_force_energy(s.ctx))
    with pytest.raises(Exception, match="without branching"):
        s.set_dmc(1e-3, 10)
    s.close()


def test_host_driver_force_energy_file(tmp_path):
    """forceEnergy.dat: header always (main.cpp:273 constructs the estimator unconditionally), one %.12e average per block
    for svmc with a {"kind": "forceEnergy"} measurement, header only for dmc.  Under |psi|^2 sampling <forceEnergy> =
    <E_L> (integration by parts), so for the harmonic trap both sit near the variational energy."""
    emu = hd.BIN.endswith("_emu")                 # the thread emulation is ~1e4 x slower: fewer chains, looser statistics
    cfg = dict(cases.vmc_case(), walkers=10 if emu else 512, stepsPerBlock=8 if emu else 200, nBlocks=4, correlationSteps=1)
    cfg["measurements"] = cfg["measurements"] + [{"kind": "forceEnergy", "label": "forceEnergy"}]
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    lines = open(tmp_path / "forceEnergy.dat").read().split("\n")
    assert lines[0] == "forceEnergy"
    fe = np.array([float(x) for x in lines[1:] if x.strip()])
    e = np.loadtxt(tmp_path / "energy.dat", skiprows=1)
    assert len(fe) == 4 and len(e) == 4 and np.isfinite(fe).all()
    assert abs(fe[1:].mean() - e[1:].mean()) < (0.4 if emu else 0.05) * abs(e[1:].mean())
    os.remove(tmp_path / "forceEnergy.dat")
    p = hd.run(dict(cfg, method="dmc", walkers=16, timeStep=1e-3, stepsPerBlock=3, nBlocks=2), str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert open(tmp_path / "forceEnergy.dat").read().strip() == "forceEnergy"


@pytest.mark.skipif(hd.BIN.endswith("_emu"), reason="statistics need ~1e7 chain steps: hardware only")
def test_vmc_statistical_parity_with_reference_run(tmp_path):
    """parity (2) for the VMC leg of BASELINE config 1: variational energy and <cm2> of 512 GPU chains against the
    reference's own single-chain run (tests/golden/vmc_reference_run.json, generated by make_golden.py --vmc)"""
    import json
    ref = json.load(open(os.path.join(hd.GOLD, "vmc_reference_run.json")))
    cfg = dict(ref["input"], walkers=512, stepsPerBlock=100, nBlocks=60, seed=5)
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    for key, fname in (("energy_per_block", "energy.dat"), ("cm2_per_block", "cm2.dat")):
        a = np.loadtxt(tmp_path / fname, skiprows=1)[10:]
        b = np.asarray(ref[key])[20:]
        sa, sb = hd._blocked_error(a), hd._blocked_error(b)
        z = abs(a.mean() - b.mean()) / np.hypot(sa, sb)
        assert z < 3.0, (key, a.mean(), sa, b.mean(), sb)


# ---- wave<i>.dat: the radial functors tabulated by the device (main.cpp:220-227, jastrow.h:42-63) -------------
@pytest.mark.parametrize("name", ["gaussiangaussian", "gaussian", "dipolar_rep"] + sorted(cases.JASTROWS))
def test_wavefunction_tabulation_vs_oracle(port, name):
    J = {"gaussiangaussian": cases.GG, "gaussian": {"kind": "gaussian", "alpha": 0.3}, "dipolar_rep": cases.DIPOLAR_REP}.get(name) or cases.JASTROWS[name]
    L = 4.0 if name == "hardSphere" else 5.0
    cfg = cases.pbc(6, L, jastrow=J)
    s = System.from_input(cfg)
    # dipolar_rep: u'' ~ y^2 (K-ratio differences) with y = 2 sqrt(D/x): below x = 0.05 the cancellation amplifies the last-ulp
    # differences between libm and the CUDA math library past the parity tolerance, so the table starts there
    lo = 0.05 if name == "dipolar_rep" else 0.0
    got = s.tabulate_wavefunction(0, lo, L / 2, 500)
    s.close()
    exp = port.tabulate(J, lo, L / 2, 500)
    np.testing.assert_array_equal(got[:, 0], exp[:, 0])
    ok = np.isfinite(exp).all(axis=1)
    assert ok.sum() > 100
    np.testing.assert_array_equal(np.isfinite(got).all(axis=1), ok)
    for col in (1, 2, 3):
        assert rel(got[ok, col], exp[ok, col]) < RTOL, col


def test_host_driver_writes_wave_tables(tmp_path):
    """wave<i>.dat per wavefunction term: 10000 rows for two-body terms, 100000 for three-body terms, empty for
    one-body terms (jastrowWavefunctionTwoBody.h:81-83, jastrowWavefunctionThreeBody.h:140-144, wavefunction.h:64)"""
    cfg = dict(cases.three_body(N=5, L=3.0), walkers=4, delta_walkers=1, nBlocks=1, stepsPerBlock=1)
    cfg["wavefunctions"].append({"kind": "jastrow1b", "set": 0, "jastrow": {"kind": "gaussian", "alpha": 0.01}})
    p = hd.run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    rows = [len(open(tmp_path / ("wave%d.dat" % i)).read().splitlines()) for i in range(3)]
    assert rows == [10000, 100000, 0]
    first = [float(v) for v in open(tmp_path / "wave0.dat").readline().split()]
    alpha, u0, x = 0.5, -0.5, 0.5 * 1.5 / 10000
    assert abs(first[0] - x) < 1e-9 and abs(first[1] - u0 * np.exp(-alpha * x * x)) < 1e-5     # jastrowGaussGauss.h:19
