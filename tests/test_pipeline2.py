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
