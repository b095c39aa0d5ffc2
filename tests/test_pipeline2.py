"""GPU tests of the production step pipeline (pipeline 2, qmc4_b200/csrc/kernels_v2.cu: move -> scan -> estimators, the
collective and the migration on a second stream) against the fused pipeline 1 of kernels.cu (QMC4_PIPELINE=1), which the
traced-step parity tests pin on the oracle: trajectories must be bitwise identical (same Philox draws, same formulas) and
the fixed-point g(r) identical."""
import os

import numpy as np
import pytest

import cases

pytestmark = [pytest.mark.gpu]


def _run(pipeline, cfg, R, steps):
    from qmc4_b200.system import System
    if pipeline == 1:
        os.environ["QMC4_PIPELINE"] = "1"
    else:
        os.environ.pop("QMC4_PIPELINE", None)       # pipeline 2 is the default
    os.environ["QMC4_LEAN"] = "0"                   # the general kernels of both pipelines share every summation order
    os.environ["QMC4_SMALL"] = "0"
    try:
        s = System.from_input(cfg)
        assert s.info("pipeline") == pipeline
        s.upload(R)
        s.step(steps, 1)
        b = s.block_results()
        Rn, En = s.download(with_energy=True)
        g = s.estimator(0) if cfg["measurements"] else None
        s.close()
    finally:
        os.environ.pop("QMC4_PIPELINE", None)
        os.environ.pop("QMC4_LEAN", None)
        os.environ.pop("QMC4_SMALL", None)
    return b, Rn, En, g


@pytest.mark.parametrize("name", ["pbc64", "trap", "three_body"])
def test_pipeline2_is_bitwise_pipeline1(name):
    if name == "pbc64":
        cfg = cases.pbc(64, 4.0, walkers=5000, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 200, "minx": 0, "maxx": 2.0}])
        R = np.random.default_rng(1).uniform(-2, 2, size=(5000, 192))
    elif name == "trap":
        cfg = cases.trap(10, walkers=700)
        R = np.random.default_rng(2).normal(0, 0.7, size=(700, 30))
    else:
        cfg = dict(cases.three_body(16), walkers=300)
        R = np.random.default_rng(3).uniform(-2.5, 2.5, size=(300, 48))
    b1, R1, E1, g1 = _run(1, cfg, R, 25)
    b2, R2, E2, g2 = _run(2, cfg, R, 25)
    np.testing.assert_array_equal(R1, R2)
    np.testing.assert_array_equal(E1, E2)
    for k in ("energy_sum", "energy_count", "n_moves", "n_accepted", "population", "energy_shift"):
        assert b1[k] == b2[k], k
    if g1 is not None:
        assert g1[1] == g2[1]
        np.testing.assert_allclose(g1[0], g2[0], rtol=1e-12)


def _traced(cfg, R, lean):
    from qmc4_b200.system import System
    os.environ["QMC4_LEAN"] = "1" if lean else "0"
    try:
        s = System.from_input(cfg)
        assert s.info("pair_kernel") == (1 if lean else 0)
        s.upload(R)
        tr = s.step_traced(accumulate=bool(cfg["measurements"]))
        Rn, En = s.download(with_energy=True)
        b = s.block_results()
        s.close()
    finally:
        os.environ.pop("QMC4_LEAN", None)
    return tr, Rn, En, b


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize("N,L,pot", [(2, 3.0, "potentialBarrier"), (10, 3.0, "squareWell"), (16, 2.6, "potentialBarrier"), (17, 2.6, None), (31, 3.2, "potentialBarrier"), (32, 3.2, "potentialBarrier"), (33, 3.2, None), (64, 4.0, "potentialBarrier"),
                                     (65, 4.0, "squareWell"), (100, 4.7, "potentialBarrier"), (160, 5.5, None), (256, 6.3496, "potentialBarrier"),
                                     (512, 8.0, "potentialBarrier")])
def test_pair_kernel_agrees_with_the_general_kernels(N, L, pot):
    """the tournament-schedule kernel of the one-species / one-pair-pass shapes (kernels_pair.cu) against the general move
    kernel on the same step: identical random numbers and proposals, identical decisions, sums that only differ by the order
    of the floating-point additions"""
    W = 24 if N <= 256 else 10
    cfg = cases.pbc(N, L, pot=pot, walkers=W)
    R = np.random.default_rng(N).uniform(-L, L, size=(W, 3 * N))
    tg, Rg, Eg, bg = _traced(cfg, R, lean=False)
    tl, Rl, El, bl = _traced(cfg, R, lean=True)
    np.testing.assert_array_equal(tl["xi"], tg["xi"])
    np.testing.assert_array_equal(tl["Rprop"], tg["Rprop"])
    np.testing.assert_array_equal(tl["u_metropolis"], tg["u_metropolis"])
    np.testing.assert_array_equal(tl["u_branch"], tg["u_branch"])
    np.testing.assert_array_equal(tl["V"], tg["V"])                      # a pair count times V0: exact
    assert _rel(tl["logpsi"], tg["logpsi"]) < 1e-13 and _rel(tl["lap"], tg["lap"]) < 1e-12
    assert _rel(tl["Fprop"], tg["Fprop"]) < 1e-12 and _rel(tl["Eprop"], tg["Eprop"]) < 1e-11 and _rel(tl["q"], tg["q"]) < 1e-9
    np.testing.assert_array_equal(tl["accepted"], tg["accepted"])
    np.testing.assert_array_equal(tl["n_desc"], tg["n_desc"])
    assert tl["n_after"] == tg["n_after"] and _rel(np.sort(El), np.sort(Eg)) < 1e-11
    assert bl["n_moves"] == bg["n_moves"] == W and bl["n_accepted"] == bg["n_accepted"]


@pytest.mark.parametrize("N", [40, 130])
def test_pair_kernel_in_a_non_cubic_box(port, N):
    """per-dimension box scales of the fixed-point minimum image (kernels_pair.cu), against the general kernels and the oracle"""
    W = 12
    box = [3.0, 3.7, 4.4] if N == 40 else [4.6, 5.3, 6.1]
    cfg = cases.pbc(N, box[0], walkers=W)
    cfg["lBox"] = box
    R = np.random.default_rng(7 * N).uniform(-1.0, 1.0, size=(W, 3, N)) * np.asarray(box)[None, :, None]
    R = R.reshape(W, 3 * N)
    tg, Rg, Eg, bg = _traced(cfg, R, lean=False)
    tl, Rl, El, bl = _traced(cfg, R, lean=True)
    np.testing.assert_array_equal(tl["Rprop"], tg["Rprop"])
    np.testing.assert_array_equal(tl["V"], tg["V"])
    assert _rel(tl["Fprop"], tg["Fprop"]) < 1e-12 and _rel(tl["Eprop"], tg["Eprop"]) < 1e-11 and _rel(tl["logpsi"], tg["logpsi"]) < 1e-13
    np.testing.assert_array_equal(tl["accepted"], tg["accepted"])
    ep = port.eval(cfg, tl["Rprop"])
    for k, t in (("logpsi", "logpsi"), ("grad", "Fprop"), ("lap", "lap"), ("V", "V"), ("E", "Eprop")):
        assert _rel(tl[t], ep[k]) < 1e-10, k


@pytest.mark.parametrize("N,L", [(64, 4.0), (256, 6.3496), (512, 8.0)])
def test_pair_kernel_traced_step_vs_oracle_at_baseline_sizes(port, N, L):
    """BASELINE configs 2, 3, 5: proposal, evaluation, Green ratio and Metropolis decision of the production kernel against
    dmcDriver::update restated in the oracle, given the random numbers the kernel drew"""
    from qmc4_b200.system import System
    W = 6
    cfg = cases.pbc(N, L, walkers=W)
    R = np.random.default_rng(1000 + N).uniform(-L / 2, L / 2, size=(W, 3 * N))
    s = System.from_input(cfg)
    assert s.info("pair_kernel") == 1
    s.upload(R)
    e0 = port.eval(cfg, R)
    tr = s.step_traced()
    ep = port.eval(cfg, tr["Rprop"])
    for k, t in (("logpsi", "logpsi"), ("grad", "Fprop"), ("lap", "lap"), ("V", "V"), ("E", "Eprop")):
        assert _rel(tr[t], ep[k]) < 1e-10, k
    q = port.green_ratio(cfg, cfg["timeStep"], tr["Rprop"], ep["grad"], R, e0["grad"])
    assert _rel(tr["q"], q) < 1e-9
    up = port.update(cfg, cfg["timeStep"], R, tr["xi"], tr["u_metropolis"])
    np.testing.assert_array_equal(up["accepted"], tr["accepted"])
    s.close()


def _traced_small(cfg, R, small):
    from qmc4_b200.system import System
    os.environ["QMC4_SMALL"] = "1" if small else "0"
    os.environ["QMC4_LEAN"] = "0"
    try:
        s = System.from_input(cfg)
        assert s.info("small_kernel") == (1 if small else 0)
        s.upload(R)
        tr = s.step_traced()
        Rn, En = s.download(with_energy=True)
        b = s.block_results()
        wpw = s.info("small_walkers_per_warp")
        s.close()
    finally:
        os.environ.pop("QMC4_SMALL", None)
        os.environ.pop("QMC4_LEAN", None)
    return tr, Rn, En, b, wpw


def _small_cases():
    out = {}
    for N in (10, 7, 3):
        out["trap_N%d" % N] = (cases.trap(N, walkers=37), np.random.default_rng(N).normal(0, 0.7, size=(37, 3 * N)), {10: 2, 7: 4, 3: 8}[N])
    out["pbc_gg_N16"] = (cases.pbc(16, 2.6, walkers=21), np.random.default_rng(5).uniform(-2.6, 2.6, size=(21, 48)), 2)
    out["pbc_squareWellJastrow_N12"] = (cases.pbc(12, 5.0, jastrow=cases.JASTROWS["squareWell"], pot="squareWell", walkers=19),
                                        np.random.default_rng(6).uniform(-2.5, 2.5, size=(19, 36)), 2)
    pt = cases.pbc(9, 3.0, jastrow=cases.JASTROWS["poschTeller"], pot=None, walkers=26)
    pt["potentials"] = [{"kind": "poschTeller2b", "sets": [0, 0], "R0": 0.8}, {"kind": "dipolar2b", "sets": [0, 0], "D": 0.3}]
    pt["lBox"] = [5.0, 5.0, 5.0]
    out["pbc_poschTeller_dipolar_N9"] = (pt, np.random.default_rng(7).uniform(-2.5, 2.5, size=(26, 27)), 2)
    return out


@pytest.mark.parametrize("name", sorted(_small_cases()))
def test_small_system_kernel_agrees_with_the_general_kernels_and_the_oracle(port, name):
    """several walkers per warp (kernels_small.cu) against one warp per walker on the same step — identical random numbers,
    proposals and decisions, sums that differ by the order of additions — and against dmcDriver::update in the oracle"""
    cfg, R, per_warp = _small_cases()[name]
    tg, Rg, Eg, bg, _ = _traced_small(cfg, R, small=False)
    ts, Rs, Es, bs, wpw = _traced_small(cfg, R, small=True)
    assert wpw == per_warp
    np.testing.assert_array_equal(ts["xi"], tg["xi"])
    np.testing.assert_array_equal(ts["Rprop"], tg["Rprop"])
    np.testing.assert_array_equal(ts["u_metropolis"], tg["u_metropolis"])
    assert _rel(ts["logpsi"], tg["logpsi"]) < 1e-12 and _rel(ts["lap"], tg["lap"]) < 1e-12 and _rel(ts["V"], tg["V"]) < 1e-12
    assert _rel(ts["Fprop"], tg["Fprop"]) < 1e-12 and _rel(ts["Eprop"], tg["Eprop"]) < 1e-11 and _rel(ts["q"], tg["q"]) < 1e-9
    np.testing.assert_array_equal(ts["accepted"], tg["accepted"])
    np.testing.assert_array_equal(ts["n_desc"], tg["n_desc"])
    assert bs["n_moves"] == bg["n_moves"] == len(R) and bs["n_accepted"] == bg["n_accepted"]
    ep = port.eval(cfg, ts["Rprop"])
    for k, t in (("logpsi", "logpsi"), ("grad", "Fprop"), ("lap", "lap"), ("V", "V"), ("E", "Eprop")):
        assert _rel(ts[t], ep[k]) < 1e-10, k
    up = port.update(cfg, cfg["timeStep"], R, ts["xi"], ts["u_metropolis"])
    np.testing.assert_array_equal(up["accepted"], ts["accepted"])


def _hist_after_steps(cfg, R, lean_hist, steps=2):
    from qmc4_b200.system import System
    os.environ["QMC4_LEAN_HIST"] = "1" if lean_hist else "0"
    try:
        s = System.from_input(cfg)
        assert s.info("pair_hist_kernel") == (1 if lean_hist else 0)
        s.upload(R)
        pops = []
        for _ in range(steps):
            s.step(1, 1)
            pops.append(s.download())          # the post-branching population the estimator just accumulated over
        sums, w = s.estimator(0)
        s.close()
    finally:
        os.environ.pop("QMC4_LEAN_HIST", None)
    return sums, w, pops


@pytest.mark.parametrize("N,L,bins,minx", [(16, 2.6, 26, 0.0), (33, 3.2, 64, 0.1), (64, 4.0, 1000, 0.0), (100, 4.7, 200, 0.0)])
def test_pair_hist_kernel_vs_general_estimator_and_oracle(port, N, L, bins, minx):
    """g(r) through the fixed-point-coordinate kernel (kernels_pair.cu) == the general estimator kernel == the oracle's
    pairCorrelation::accumulate over the post-branching populations with unit weight"""
    W = 40
    meas = [{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": bins, "minx": minx, "maxx": L / 2}]
    cfg = cases.pbc(N, L, walkers=W, meas=meas)
    R = np.random.default_rng(3 * N).uniform(-L, L, size=(W, 3 * N))
    sg, wg, pg = _hist_after_steps(cfg, R, lean_hist=False)
    sl, wl, pl = _hist_after_steps(cfg, R, lean_hist=True)
    assert wl == wg == sum(len(p) for p in pl)
    for a, b in zip(pg, pl):
        np.testing.assert_array_equal(a, b)            # same trajectory: the estimator kernels do not touch the walkers
    assert _rel(sl, sg) < 1e-9
    ref = np.zeros(bins)
    for p in pl:
        rs, rw = port.histogram(cfg, p, 0, bins)
        ref += rs
    assert _rel(sl, ref) < 1e-9


@pytest.mark.parametrize("W", [2100, 40000])
def test_scan_over_many_ctas_upload_keeps_every_walker(W):
    """the scan needs several CTAs above 1024 walkers (40000: more CTAs than one warp of the finishing CTA sums -- round 2 lost
    every walker beyond 32768 because the published totals were cleared while thread 0 was still adding them): the population
    after an upload is the uploaded one, and a step without branching (svmc, main.cpp:393-395) keeps it"""
    from qmc4_b200.system import System
    cfg = cases.trap(2, walkers=W)
    cfg["method"] = "svmc"
    R = np.random.default_rng(W).normal(0, 1.0, size=(W, 6))
    s = System.from_input(cfg)
    s.upload(R)
    assert s.info("global_population") == W
    np.testing.assert_array_equal(s.download(), R)
    s.step(2, 1)
    assert s.block_results()["population"] == W and len(s.download()) == W
    s.close()


@pytest.mark.parametrize("W,pipeline", [(3000, 2), (3000, 1), (5300, 2), (40000, 2)])
def test_scan_over_many_ctas_branching_bookkeeping(W, pipeline):
    """branching across several scan CTAs: the new population is sum(nDescendants) (branching.cpp:16-63) and the walkers downloaded
    afterwards are the accepted / rejected configurations in stable prefix-sum order"""
    from qmc4_b200.system import System
    cfg = cases.trap(2, walkers=W)
    cfg["timeStep"] = 5e-2                       # a visible spread of nDescendants (0, 1, 2)
    R = np.random.default_rng(W).normal(0, 1.0, size=(W, 6))
    os.environ["QMC4_PIPELINE"] = str(pipeline)
    try:
        s = System.from_input(cfg)
        s.upload(R)
        tr = s.step_traced()
        Rn = s.download()
        b = s.block_results()
        s.close()
    finally:
        os.environ.pop("QMC4_PIPELINE", None)
    nd = tr["n_desc"]
    assert tr["n_before"] == W and tr["n_after"] == nd.sum() == len(Rn) == b["population"]
    assert set(np.unique(nd)) - {0, 1} and nd.min() >= 0             # branching did clone some walkers
    moved = np.where(tr["accepted"][:, None] != 0, tr["Rprop"], R)
    np.testing.assert_array_equal(Rn, np.repeat(moved, nd, axis=0))
