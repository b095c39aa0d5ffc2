"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
  * the golden vectors produced by the reference itself (tests/golden/),
  * the CPU oracle on the same seeded inputs,
  * size-independent properties at BASELINE sizes (no oracle needed),
  * the reference's own CPU DMC runs (statistical parity (2), 2 sigma).
Tolerance for floating point: 1e-10 relative (BASELINE.json north_star), written below."""
import json
import os

import numpy as np
import pytest

import cases
from qmc4_b200.system import System, invalidInput

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EVAL = np.load(os.path.join(GOLD, "eval_cases.npz"))
PIECES = np.load(os.path.join(GOLD, "pieces.npz"))
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("name", sorted(n for n in cases.eval_cases() if n not in cases.NEW_PATH_EVAL_CASES))
def test_eval_vs_golden_and_oracle(port, name):
    cfg, R = cases.eval_cases()[name]
    s = System.from_input(cfg)
    o = s.eval(R)
    s.close()
    r = port.eval(cfg, R)
    np.testing.assert_array_equal(o["valid"], EVAL[name + "/valid"])
    for k in ("logpsi", "grad", "lap", "V", "E"):
        assert rel(o[k], EVAL[name + "/" + k]) < RTOL, ("golden", k)
        assert rel(o[k], r[k]) < RTOL, ("oracle", k)


@pytest.mark.parametrize("wpw", ["1", "2", "4", "8"])
def test_eval_independent_of_work_decomposition(port, wpw, monkeypatch):
    """same numbers (to rounding) whichever number of warps shares a walker"""
    monkeypatch.setenv("QMC4_WPW", wpw)
    cfg, R = cases.eval_cases()["pbc_gg_barrier_N100"]
    s = System.from_input(cfg)
    assert s.info("warps_per_walker") == int(wpw)
    o = s.eval(R)
    s.close()
    for k in ("logpsi", "grad", "lap", "V", "E"):
        assert rel(o[k], EVAL["pbc_gg_barrier_N100/" + k]) < RTOL, k


def test_generic_kernel_agrees_with_specialised(monkeypatch):
    cfg, R = cases.eval_cases()["pbc_gg_barrier_N65"]
    s = System.from_input(cfg)
    a = s.eval(R)
    s.close()
    monkeypatch.setenv("QMC4_FORCE_GENERIC", "1")
    s = System.from_input(cfg)
    assert s.info("kernel_variant") == 3
    b = s.eval(R)
    s.close()
    for k in ("logpsi", "grad", "lap", "V", "E"):
        assert rel(a[k], b[k]) < RTOL, k


@pytest.mark.parametrize("N,L", [(64, 4.0), (256, 6.3496), (512, 8.0)])
def test_eval_baseline_sizes_vs_oracle(port, N, L):
    cfg = cases.pbc(N, L)
    R = np.random.default_rng(N).uniform(-L / 2, L / 2, size=(3, 3 * N))
    s = System.from_input(cfg)
    o = s.eval(R)
    s.close()
    r = port.eval(cfg, R)
    for k in ("logpsi", "grad", "lap", "V", "E"):
        assert rel(o[k], r[k]) < RTOL, k


def test_properties_at_full_size():
    """config 5 shape (N=512): size-independent invariants of the pair path."""
    N, L = 512, 8.0
    cfg = cases.pbc(N, L)
    rng = np.random.default_rng(5)
    R = rng.uniform(-L / 2, L / 2, size=(16, 3, N))
    s = System.from_input(cfg)
    a = s.eval(R)
    # (i) Newton's third law: pair-only wavefunction => sum_i grad_i == 0
    g = a["grad"].reshape(16, 3, N)
    assert np.abs(g.sum(axis=2)).max() < 1e-9 * np.abs(g).max() * N
    # (ii) periodicity: shifting particles by lattice vectors changes nothing
    shift = rng.integers(-3, 4, size=R.shape) * L
    b = s.eval(R + shift)
    for k in ("logpsi", "grad", "lap", "V", "E"):
        assert rel(b[k], a[k]) < 1e-9, k
    # (iii) permutation symmetry: relabelling particles permutes the gradient
    perm = rng.permutation(N)
    c = s.eval(R[:, :, perm])
    assert rel(c["logpsi"], a["logpsi"]) < 1e-11
    assert rel(c["E"], a["E"]) < 1e-9
    assert rel(c["grad"].reshape(16, 3, N), g[:, :, perm]) < 1e-10
    # (iv) V counts pairs inside the barrier: integer multiple of V0
    assert np.allclose(a["V"] / 2.0, np.round(a["V"] / 2.0))
    s.close()


@pytest.mark.parametrize("name", ["pbc_gg_barrier_N33", "mixture", "trap_N10", "three_body_N16", "pbc_gg_barrier_N100"])
def test_traced_step_vs_oracle_update(port, name):
    """One propagate step, walker by walker: proposal, evaluation, Green ratio, Metropolis decision,
    committed state and nDescendants against dmcDriver::update / branching restated in the oracle,
    given the random numbers the kernel drew."""
    cfg, R = cases.eval_cases()[name]
    cfg = dict(cfg, walkers=len(R))
    s = System.from_input(cfg)
    s.upload(R)
    e0 = port.eval(cfg, R)
    eT = port.energy_shift(e0["E"], cfg["timeStep"], cfg["walkers"], int(cfg["walkers"] * 0.1))
    assert abs(s.block_results()["energy_shift"] - eT) < 1e-10 * abs(eT)
    tr = s.step_traced()
    par = tr["parent"]
    np.testing.assert_array_equal(par, np.arange(len(R)))
    Rp = port.drift_diffuse(cfg, cfg["timeStep"], R, e0["grad"], tr["xi"])
    assert rel(tr["Rprop"], Rp) < 1e-14
    ep = port.eval(cfg, tr["Rprop"])
    for k, t in (("logpsi", "logpsi"), ("grad", "Fprop"), ("lap", "lap"), ("V", "V"), ("E", "Eprop")):
        assert rel(tr[t], ep[k]) < RTOL, k
    q = port.green_ratio(cfg, cfg["timeStep"], tr["Rprop"], ep["grad"], R, e0["grad"])
    assert rel(tr["q"], q) < 1e-9
    up = port.update(cfg, cfg["timeStep"], R, tr["xi"], tr["u_metropolis"])
    np.testing.assert_array_equal(up["accepted"], tr["accepted"])
    nd = port.n_descendants(up["E"], e0["E"], eT, cfg["timeStep"], tr["u_branch"])
    np.testing.assert_array_equal(nd, tr["n_desc"])
    # branched population: same multiset of walkers as the reference's in-place algorithm
    Rn, En = s.download(with_energy=True)
    after = port.branch(nd, up["E"])
    assert len(En) == len(after) == tr["n_after"]
    assert rel(np.sort(En), np.sort(after)) < RTOL
    # E_T of the new population (dmcDriver.cpp:141-143)
    eT2 = port.energy_shift(after, cfg["timeStep"], cfg["walkers"], int(cfg["walkers"] * 0.1))
    assert abs(s.block_results()["energy_shift"] - eT2) < 1e-9 * abs(eT2)
    s.close()


def test_rejected_walkers_keep_parent_state(port):
    """reject => wNew = wOld (dmcDriver.cpp:73-76): huge time step makes most moves fail"""
    cfg, R = cases.eval_cases()["pbc_gg_barrier_N33"]
    cfg = dict(cfg, walkers=len(R), timeStep=0.5, method="svmc")
    s = System.from_input(cfg)
    s.upload(R)
    e0 = port.eval(cfg, R)
    tr = s.step_traced()
    up = port.update(cfg, 0.5, R, tr["xi"], tr["u_metropolis"])
    np.testing.assert_array_equal(up["accepted"], tr["accepted"])
    assert (tr["accepted"] == 0).any()
    Rn, En = s.download(with_energy=True)
    assert rel(Rn, up["R"]) < 1e-13 and rel(En, up["E"]) < RTOL
    assert (tr["n_desc"] == 1).all()           # svmc: branching disabled (main.cpp:393-395)
    s.close()


def test_constraint_violation_rejects_and_invalid_start_throws(port):
    cfg, R = cases.eval_cases()["three_body_hs_N12"]
    valid = EVAL["three_body_hs_N12/valid"].astype(bool)
    cfg = dict(cfg, walkers=int(valid.sum()))
    s = System.from_input(cfg)
    with pytest.raises((invalidInput, RuntimeError), match="Inalid initial"):
        s.upload(R)                              # dmcDriver.cpp:166-170
    s.close()
    s = System.from_input(cfg)
    s.upload(R[valid])
    tr = s.step_traced()
    ep = port.eval(cfg, tr["Rprop"])
    np.testing.assert_array_equal(tr["valid"], ep["valid"])
    assert not (tr["accepted"][tr["valid"] == 0]).any()
    s.close()


def _hist_cases():
    return {"mixture": (cases.mixture(), 2), "three_body_N16": (cases.three_body(16), 1),
            "pair_N33": (cases.pbc(33, 3.2, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 64, "minx": 0, "maxx": 1.6}]), 1)}


@pytest.mark.parametrize("name", sorted(_hist_cases()))
def test_estimators_on_given_configurations(port, name):
    cfg, nest = _hist_cases()[name]
    R = PIECES["hist_" + name + "/R"]
    s = System.from_input(cfg)
    s.accumulate_configs(R)
    for i in range(nest):
        sums, w = s.estimator(i)
        assert w == PIECES["hist_%s/weight%d" % (name, i)][0]
        assert rel(sums, PIECES["hist_%s/sums%d" % (name, i)]) < RTOL
    s.clear_estimators()
    sums, w = s.estimator(0)
    assert w == 0 and not sums.any()
    s.close()


def test_structure_factor_vs_oracle(port):
    cfg = cases.structure_factor()
    R = np.random.default_rng(8).uniform(-6, 6, size=(12, 72))
    s = System.from_input(cfg)
    s.accumulate_configs(R)
    for i in range(2):
        sums, w = s.estimator(i)
        ref_s, ref_w = port.histogram(cfg, R, i, cfg["measurements"][i]["bins"])
        assert w == ref_w
        assert rel(sums, ref_s) < 1e-9
    s.close()


def test_estimators_in_the_step_use_branching_weights(port):
    """g(r) accumulated inside the propagate kernel == oracle histogram of the post-branching
    population (dmcDriver.cpp:257-264), including rejected moves and killed/copied walkers."""
    cfg = cases.pbc(33, 3.2, walkers=200, timeStep=0.05,
                    meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 64, "minx": 0, "maxx": 1.6}])
    R = np.random.default_rng(21).uniform(-1.6, 1.6, size=(200, 99))
    s = System.from_input(cfg)
    s.upload(R)
    tr = s.step_traced(accumulate=True)
    assert (tr["accepted"] == 0).any() and (tr["n_desc"] != 1).any()
    Rn = s.download()
    sums, w = s.estimator(0)
    ref_s, ref_w = port.histogram(cfg, Rn, 0, 64)
    assert w == ref_w == len(Rn)
    assert rel(sums, ref_s) < 1e-9
    b = s.block_results()
    assert b["energy_count"] == len(Rn)
    s.close()


def test_trimer_and_sq_in_the_step(port):
    cfg = cases.three_body(16)
    cfg = dict(cfg, walkers=64, timeStep=0.02)
    R = np.random.default_rng(22).uniform(-2.5, 2.5, size=(64, 48))
    s = System.from_input(cfg)
    s.upload(R)
    s.step_traced(accumulate=True)
    Rn = s.download()
    sums, w = s.estimator(0)
    ref_s, ref_w = port.histogram(cfg, Rn, 0, 50)
    assert w == ref_w and rel(sums, ref_s) < 1e-9
    s.close()


def test_run_to_run_determinism():
    cfg = cases.pbc(64, 4.0, walkers=300)
    R = np.random.default_rng(1).uniform(-2, 2, size=(300, 192))
    outs = []
    for _ in range(2):
        s = System.from_input(cfg)
        s.upload(R)
        s.step(20, 1)
        outs.append(s.download(with_energy=True))
        s.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


def test_population_control_and_capacity_flag():
    cfg = cases.trap(10, walkers=200)
    s = System.from_input(cfg)
    s.upload(np.tile(np.random.default_rng(2).normal(0, 0.7, size=(1, 30)), (200, 1)))
    s.step(400, 1)
    b = s.block_results()
    assert b["overflow"] == 0 and 120 < b["population"] < 300
    assert 0.99 < b["n_accepted"] / b["n_moves"] <= 1.0
    s.close()


def test_normal_deviates_statistics():
    cfg = cases.pbc(64, 4.0, walkers=2000)
    s = System.from_input(cfg)
    s.upload(np.random.default_rng(3).uniform(-2, 2, size=(2000, 192)))
    xi = s.step_traced()["xi"].ravel()
    s.close()
    n = len(xi)
    assert abs(xi.mean()) < 5 / np.sqrt(n)
    assert abs(xi.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs((xi ** 4).mean() - 3) < 5 * np.sqrt(96 / n)
    assert abs((xi ** 3).mean()) < 5 * np.sqrt(15 / n)
    from scipy import stats
    assert stats.kstest(xi[:200000], "norm").pvalue > 1e-4


def _blocked_error(x):
    """standard error with reblocking (Flyvbjerg-Petersen): max over blocking levels"""
    x = np.asarray(x, float)
    errs = []
    while len(x) >= 8:
        errs.append(x.std(ddof=1) / np.sqrt(len(x)))
        x = 0.5 * (x[: len(x) // 2 * 2: 2] + x[1: len(x) // 2 * 2: 2])
    return max(errs)


@pytest.mark.parametrize("name", ["trap_N10", "pbc_N16"])
def test_dmc_statistical_parity_with_reference_run(name):
    """parity (2): mixed-estimator energy and g(r) within 2 sigma of the reference's CPU run
    (fixtures: tests/golden/dmc_reference_runs.json, produced by oracle/_ref/dmc_ref)."""
    fx = json.load(open(os.path.join(GOLD, "dmc_reference_runs.json")))[name]
    cfg = fx["input"]
    skip = 15
    e_ref = np.asarray(fx["energy_per_block"])[skip:]
    N = sum(cfg["N"])

    def run(seed):
        rng = np.random.default_rng(99 + seed)
        s = System.from_input(dict(cfg, seed=int(cfg["seed"]) + seed))
        box = cfg.get("initialConditionGenerator", {}).get("lBox", cfg["lBox"])
        start = rng.uniform(-0.5, 0.5, size=(1, 3, N)) * np.asarray(box)[None, :, None]
        s.upload(np.tile(start.reshape(1, -1), (cfg["walkers"], 1)))      # main.cpp:380-388: all walkers share one start
        e, g = [], []
        for b in range(cfg["nBlocks"]):
            s.step(cfg["stepsPerBlock"], cfg["correlationSteps"] + 1)
            r = s.block_results()
            e.append(r["energy_sum"] / r["energy_count"])
            if cfg["measurements"]:
                sums, w = s.estimator(0)
                g.append(sums / w)
                s.clear_estimators()
        s.close()
        return np.asarray(e)[skip:], (np.asarray(g)[skip:] if g else None)

    def z_energy(es):
        m = np.mean([e.mean() for e in es])
        err = np.sqrt(sum(_blocked_error(e) ** 2 for e in es)) / len(es)
        sig = np.hypot(err, _blocked_error(e_ref))
        return abs(m - e_ref.mean()) / max(sig, 1e-300), (m, e_ref.mean(), sig)

    # A 2-sigma criterion rejects a correct code one time in twenty.  The trajectory is a deterministic function of the seed
    # and of every rounding in the kernels, so a run that lands outside is followed by ONE independent run (next seed) and the
    # pooled estimate is tested at the same 2 sigma: a bias survives the pooling, a fluctuation does not.
    runs = [run(0)]
    z, info = z_energy([r[0] for r in runs])
    if z >= 2:
        runs.append(run(1))
        z, info = z_energy([r[0] for r in runs])
    assert z < 2, info
    if runs[0][1] is not None:
        g_ref = np.asarray(fx[cfg["measurements"][0]["label"] + "_per_block"])[skip:]

        def z_g(gs):
            m = np.mean([g.mean(0) for g in gs], axis=0)
            err = np.sqrt(sum(np.asarray([_blocked_error(g[:, i]) for i in range(g.shape[1])]) ** 2 for g in gs)) / len(gs)
            sg = np.hypot(err, [_blocked_error(g_ref[:, i]) for i in range(g_ref.shape[1])])
            return np.abs(m - g_ref.mean(0)) / np.maximum(sg, 1e-12)
        z = z_g([r[1] for r in runs])
        if not ((z < 2).mean() > 0.8 and z.max() < 4) and len(runs) == 1:
            runs.append(run(1))
            z = z_g([r[1] for r in runs])
        # bins are correlated; require the bulk within 2 sigma and none beyond 4
        assert (z < 2).mean() > 0.8 and z.max() < 4, z
