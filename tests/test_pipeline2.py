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
