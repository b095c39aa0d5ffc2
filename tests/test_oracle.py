"""CPU tests (-m "not gpu"): the oracle (oracle/qmc_oracle.c) is pinned against
  (a) the golden vectors produced by the reference itself (tests/golden/make_golden.py),
  (b) the reference itself when oracle/_ref is present in this checkout,
  (c) the known-answer tests the reference's own suite holds for this path,
      restated against the oracle.
"""
import os

import numpy as np
import pytest

import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EVAL = np.load(os.path.join(GOLD, "eval_cases.npz"))
PIECES = np.load(os.path.join(GOLD, "pieces.npz"))
TOL = 1e-12     # oracle vs reference: same arithmetic, differences are rounding only


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("name", sorted(cases.eval_cases()))
def test_eval_matches_golden(port, name):
    cfg, R = cases.eval_cases()[name]
    np.testing.assert_array_equal(R, EVAL[name + "/R"])       # seeded inputs are reproducible
    o = port.eval(cfg, R)
    np.testing.assert_array_equal(o["valid"], EVAL[name + "/valid"])
    for k in ("logpsi", "grad", "lap", "V", "E"):
        assert rel(o[k], EVAL[name + "/" + k]) < TOL, k


@pytest.mark.parametrize("name", sorted(cases.eval_cases()))
def test_eval_matches_reference_itself(port, ref, name):
    cfg, R = cases.eval_cases()[name]
    a, b = ref.eval(cfg, R), port.eval(cfg, R)
    for k in ("logpsi", "grad", "lap", "V", "E"):
        assert rel(b[k], a[k]) < TOL, k
    np.testing.assert_array_equal(a["valid"], b["valid"])


def test_constraint_case_has_invalid_walkers():
    v = EVAL["three_body_hs_N12/valid"]
    assert 0 < v.sum() < len(v)


@pytest.mark.parametrize("name", ["pbc_gg_barrier_N33", "mixture", "trap_N10"])
def test_move_and_green_ratio(port, name):
    cfg, R = cases.eval_cases()[name]
    e0 = port.eval(cfg, R)
    xi = PIECES[name + "/xi"]
    Rn = port.drift_diffuse(cfg, cfg["timeStep"], R, e0["grad"], xi)
    assert rel(Rn, PIECES[name + "/Rnew"]) < 1e-15
    e1 = port.eval(cfg, Rn)
    q = port.green_ratio(cfg, cfg["timeStep"], Rn, e1["grad"], R, e0["grad"])
    assert rel(q, PIECES[name + "/q"]) < 1e-10


def _hist_cases():
    return {"mixture": (cases.mixture(), 2), "three_body_N16": (cases.three_body(16), 1),
            "pair_N33": (cases.pbc(33, 3.2, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 64, "minx": 0, "maxx": 1.6}]), 1)}


@pytest.mark.parametrize("name", sorted(_hist_cases()))
def test_histograms(port, name):
    cfg, nest = _hist_cases()[name]
    R = PIECES["hist_" + name + "/R"]
    for i in range(nest):
        s, w = port.histogram(cfg, R, i, cfg["measurements"][i]["bins"])
        assert w == PIECES["hist_%s/weight%d" % (name, i)][0]
        assert rel(s, PIECES["hist_%s/sums%d" % (name, i)]) < TOL
        assert s.sum() > 0


@pytest.mark.parametrize("name", sorted(cases.three_set_hist_cases()))
def test_three_set_and_scalar_observables(port, name):
    """trimerCorrelation / trimerCorrelationDis (histograms) and trimerhyperRadius / centerOfMassSquared
    (scalars: sum and count of a realScalarEstimator) against the fixtures produced by the reference"""
    cfg = cases.three_set_hist_cases()[name]
    R = PIECES["hist_" + name + "/R"]
    for i, m in enumerate(cfg["measurements"]):
        s, w = port.histogram(cfg, R, i, cases.meas_bins(m))
        assert w == PIECES["hist_%s/weight%d" % (name, i)][0] == len(R)
        assert rel(s, PIECES["hist_%s/sums%d" % (name, i)]) < TOL, m["kind"]
        assert s.sum() > 0


def test_three_set_observables_match_reference_itself(port, ref):
    for name, cfg in cases.three_set_hist_cases().items():
        R = np.random.default_rng(77).uniform(-2, 2, size=(5, 3 * sum(cfg["N"])))
        for i, m in enumerate(cfg["measurements"]):
            a, wa = ref.histogram(cfg, R, i, cases.meas_bins(m))
            b, wb = port.histogram(cfg, R, i, cases.meas_bins(m))
            assert wa == wb and rel(b, a) < TOL, (name, m["kind"])


def test_distinguishable_constraint_case_has_invalid_walkers():
    v = EVAL["mixture3_hs/valid"]
    assert 0 < v.sum() < len(v)


@pytest.mark.parametrize("tag", ["lower", "upper", "window"])
def test_branching_matches_reference(port, tag):
    """tst/dmc-test.cpp:15-153 (branching_lower/upper): population == sum nDescendants, each
    parent energy appears nDescendants times — here with the reference's own uniforms."""
    tau, mean, delta = PIECES["branch_%s/cfg" % tag]
    En, Eo = PIECES["branch_%s/Enew" % tag], PIECES["branch_%s/Eold" % tag]
    eT, eT2 = PIECES["branch_%s/shifts" % tag]
    assert abs(port.energy_shift(Eo, tau, mean, delta) - eT) <= 1e-12 * abs(eT)
    nd = port.n_descendants(En, Eo, eT, tau, PIECES["branch_%s/u" % tag])
    np.testing.assert_array_equal(nd, PIECES["branch_%s/nDesc" % tag])
    after = port.branch(nd, En)
    np.testing.assert_array_equal(after, PIECES["branch_%s/Eafter" % tag])
    assert len(after) == max(nd.sum(), 1)
    for i, e in enumerate(En):
        assert (after == e).sum() == nd[i]
    assert abs(port.energy_shift(after, tau, mean, delta) - eT2) <= 1e-12 * abs(eT2)


def test_branching_all_killed_keeps_one(port):
    after = port.branch(np.zeros(5, dtype=np.int32), np.arange(5.0))
    assert len(after) == 1 and after[0] == 0.0           # branching.cpp:29-32


def test_planner_matches_reference_and_balances(port):
    """tst/ptools-test.cpp:20-80: conservation and +-1 balance of the alias-method plan."""
    for p, s, a in zip(PIECES["plan/pops"], PIECES["plan/sources"], PIECES["plan/amounts"]):
        src, amt = port.plan(p)
        np.testing.assert_array_equal(src, s)
        np.testing.assert_array_equal(amt, a)
        new = p.copy()
        for r in range(len(p)):
            new[r] += amt[r]
            new[src[r]] -= amt[r]
        assert new.sum() == p.sum()
        assert np.all(np.abs(new - p.sum() // len(p)) <= 1)


# ---- the reference's own known-answer tests, restated against the oracle ------------------------
def test_ref_suite_oneBody_laplacian(port):
    """tst/wavefunction-test.cpp:21-58: 1-body Gaussian: lap(log psi) == -2*alpha*N*D exactly."""
    N, alpha = 100, 1.0
    cfg = {"N": [N], "lBox": [10., 10., 10.], "wavefunctions": [{"kind": "jastrow1b", "set": 0, "jastrow": {"kind": "gaussian", "alpha": alpha}}]}
    R = np.random.default_rng(1).uniform(-1, 1, size=(1, 3 * N))
    o = port.eval(cfg, R)
    assert o["lap"][0] == -2 * alpha * N * 3
    assert abs(o["logpsi"][0] - (-alpha) * (R * R).sum()) < 1e-5


def test_ref_suite_pair_order_and_minimum_image(port):
    """tst/wavefunction-test.cpp:60-101: k = i(i-1)/2+j, diff = min-image(r_i - r_j): checked through the
    gradient of a 2-body Gaussian (grad_i = -2a * sum_j diff_ij)."""
    N, L, a = 100, 10.0, 0.3
    cfg = {"N": [N], "lBox": [L] * 3, "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "gaussian", "alpha": a}}]}
    R = np.random.default_rng(2).uniform(-7, 7, size=(1, 3, N))
    o = port.eval(cfg, R)
    d = R[0][:, :, None] - R[0][:, None, :]
    d = d - np.round(d / L) * L
    g = -2 * a * d.sum(axis=2)
    assert rel(o["grad"].reshape(3, N), g) < 1e-12


def test_ref_suite_square_well_zero_energy(port):
    """tst/wavefunction-test.cpp:103-155: square-well Jastrow is the exact zero-energy 2-body solution:
    E_L == 0 to 1e-7, V == brute force."""
    lBox, V0, R0, Rm, alpha = 100.0, 2.4674011002723395, 1.0, 40.0, 0.5
    cfg = {"N": [2], "lBox": [lBox] * 3,
           "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0],
                              "jastrow": {"kind": "squareWell", "V0": V0, "R0": R0, "Rm": Rm, "alpha": alpha, "aInverse": 0.0, "cut_off": lBox / 2}}],
           "potentials": [{"kind": "squareWell2b", "sets": [0, 0], "V0": V0, "R0": R0}]}
    rng = np.random.default_rng(3)
    for scale in (Rm / (2. * np.sqrt(3)), 0.4):
        R = rng.uniform(-1, 1, size=(20, 6)) * scale
        o = port.eval(cfg, R)
        r = np.linalg.norm(R[:, 0::2] - R[:, 1::2], axis=1)
        np.testing.assert_allclose(o["V"], np.where(r <= R0, -V0, 0.0))
        assert np.abs(o["E"]).max() < 1e-7


def test_ref_suite_spline_reproduces_gaussian(port):
    """tst/jastrow-test.cpp:8-66: the 102-coefficient bSpline reproduces -x^2, -2x, -2 to 1e-4 on [0.1,1]
    (checked through log psi / gradient of a 2-particle system)."""
    coef = [-6.80202700e-05, 3.40101350e-05, -6.80202700e-05, -3.74111485e-04, -8.84263511e-04, -1.59847635e-03,
            -2.51674999e-03, -3.63908445e-03, -4.96547971e-03, -6.49593579e-03, -8.23045267e-03, -1.01690304e-02]
    step = 0.010101010101010102
    # extend with the exact B-spline coefficients of -x^2: c_i = -(h(i-1))^2 + h^2/3  (matches the table above)
    full = [-(step * (i - 1)) ** 2 + step * step / 3 for i in range(102)]
    assert np.allclose(full[:12], coef, atol=2e-9)
    cfg = {"N": [2], "lBox": [50.0] * 3,
           "wavefunctions": [{"kind": "jastrow2bUnDis", "sets": [0, 0],
                              "jastrow": {"kind": "bSpline", "stepSize": step, "valueRight": -1, "derivativeRight": -2, "coefficients": full}}]}
    xs = np.arange(0.1, 0.99, 1e-3)
    R = np.zeros((len(xs), 3, 2))
    R[:, 0, 0] = xs
    o = port.eval(cfg, R)
    assert np.abs(o["logpsi"] - (-xs * xs)).max() < 1e-4
    assert np.abs(o["grad"].reshape(-1, 3, 2)[:, 0, 0] - (-2 * xs)).max() < 1e-4
    assert np.abs(o["lap"] - 2 * (-2 + 2 * (-2 * xs) / xs)).max() < 1e-3


def test_update_step_consistency(port):
    """dmcDriver::update restated: accepted walkers carry the proposal, rejected ones the parent."""
    cfg, R = cases.eval_cases()["pbc_gg_barrier_N33"]
    rng = np.random.default_rng(9)
    xi = rng.normal(size=R.shape)
    u = np.full(len(R), 1e-300)              # log(u) very negative -> always accept
    up = port.update(cfg, 1e-3, R, xi, u)
    assert up["accepted"].all()
    e0 = port.eval(cfg, R)
    Rn = port.drift_diffuse(cfg, 1e-3, R, e0["grad"], xi)
    np.testing.assert_array_equal(up["R"], Rn)
    big = port.update(cfg, 10.0, R, xi * 50, np.full(len(R), 1.0 - 1e-16))   # absurd move, u ~ 1 -> mostly rejected
    rej = big["accepted"] == 0
    assert rej.any()
    np.testing.assert_array_equal(big["R"][rej], R[rej])
    np.testing.assert_array_equal(big["E"][rej], e0["E"][rej])


# ---- forward walking (correlationEstimator.cpp) ---------------------------------------------------------
FWD = np.load(os.path.join(GOLD, "forward_walking.npz")) if os.path.exists(os.path.join(GOLD, "forward_walking.npz")) else None


def _fw_genealogy(name):
    n = FWD[name + "/n"]
    R, anc = FWD[name + "/R"], FWD[name + "/anc"]
    edges = np.concatenate([[0], np.cumsum(n)])
    pops = [R[edges[i]:edges[i + 1]] for i in range(len(n))]
    ancs = [anc[edges[i]:edges[i + 1]] for i in range(len(n))]
    return pops, ancs


@pytest.mark.parametrize("name", sorted(cases.forward_walking_cases()))
def test_forward_walking_matches_golden(port, name):
    """storer ring + forward-walking estimator over a genealogy: the C restatement against what the
    reference's own realHistogramStorer / realHistogramForwardWalkingEstimator produced"""
    cfg = cases.forward_walking_cases()[name]
    iM, meas = cases.fw_measurement(cfg)
    pops, ancs = _fw_genealogy(name)
    for steps in meas["forwardWalkingSteps"]:
        sums, w = port.forward_walking(cfg, pops, ancs, iM, meas["recordSteps"], steps, cases.meas_bins(meas))
        assert w == float(FWD["%s/fw%d/weight" % (name, steps)]) and w > 0
        np.testing.assert_allclose(sums, FWD["%s/fw%d/sums" % (name, steps)], rtol=1e-13, atol=0)


def test_forward_walking_matches_reference_itself(port, ref):
    cfg = cases.forward_walking_cases()["pbc_pair"]
    iM, meas = cases.fw_measurement(cfg)
    pops, ancs = cases.synthetic_genealogy(cfg, 7, 5, 99)
    for k, steps in enumerate(meas["forwardWalkingSteps"]):
        a, wa = ref.forward_walking(cfg, pops, ancs, k, cases.meas_bins(meas))
        b, wb = port.forward_walking(cfg, pops, ancs, iM, meas["recordSteps"], steps, cases.meas_bins(meas))
        assert wa == wb
        np.testing.assert_allclose(b, a, rtol=1e-13, atol=0)


SFL = np.load(os.path.join(GOLD, "superfluid.npz")) if os.path.exists(os.path.join(GOLD, "superfluid.npz")) else None


@pytest.mark.parametrize("name", sorted(cases.superfluid_cases()))
def test_superfluid_fraction_matches_golden(port, name):
    """centre-of-mass ring + winding estimator over a genealogy: the C restatement against what the reference's own
    centerOfMassStorer / superfluidFractionEstimator produced (bit-equal: same operations in the same order)"""
    cfg = cases.superfluid_cases()[name]
    n = SFL[name + "/n"]
    edges = np.concatenate([[0], np.cumsum(n)])
    pops = [SFL[name + "/R"][edges[i]:edges[i + 1]] for i in range(len(n))]
    ancs = [SFL[name + "/anc"][edges[i]:edges[i + 1]] for i in range(len(n))]
    seen = 0
    for iM, m in enumerate(cfg["measurements"]):
        if m["kind"] != "superfluidFraction":
            continue
        sums, ws = port.superfluid_fraction(cfg, pops, ancs, iM)
        np.testing.assert_array_equal(ws, SFL["%s/%s/weights" % (name, m["label"])])
        np.testing.assert_array_equal(sums, SFL["%s/%s/sums" % (name, m["label"])])
        assert ws[0] == n.sum() and sums[0] == 0 and ws[-1] > 0
        seen += 1
    assert seen >= 1


def test_superfluid_fraction_matches_reference_itself(port, ref):
    cfg = cases.superfluid_cases()["mixture_rho"]
    pops, ancs = cases.synthetic_genealogy(cfg, 13, 5, 321)
    for iM, m in enumerate(cfg["measurements"]):
        if m["kind"] == "superfluidFraction":
            a, wa = ref.superfluid_fraction(cfg, pops, ancs, iM, m["bins"])
            b, wb = port.superfluid_fraction(cfg, pops, ancs, iM)
            np.testing.assert_array_equal(wa, wb)
            np.testing.assert_array_equal(a, b)


def test_bessel_k_pinned_on_scipy(port):
    """oracle/bessel_k.h (the Boost.Math stand-in behind the reference's dipolar_rep code in oracle/_ref, and the oracle's
    own restatement) against an independent implementation: scipy.special.kve / kv (AMOS).  Boost itself is absent, so this
    is the pin for that Jastrow's special function; the rest of jastrowDipolar.h is the reference's own compiled code."""
    import ctypes as C
    from scipy.special import kv, kve
    port.lib.orc_bessel_k.restype = C.c_double
    port.lib.orc_bessel_k.argtypes = [C.c_int, C.c_double]
    port.lib.orc_bessel_k01e.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    k0, k1 = C.c_double(0), C.c_double(0)
    for x in np.concatenate([np.logspace(-4, np.log10(700.0), 1500), np.linspace(0.05, 12, 400)]):
        port.lib.orc_bessel_k01e(float(x), C.byref(k0), C.byref(k1))
        assert abs(k0.value / kve(0, x) - 1) < 4e-15 and abs(k1.value / kve(1, x) - 1) < 4e-15, x
    for x in np.logspace(-3, np.log10(50.0), 300):
        for n in (0, 1, 2, 3, -1):
            assert abs(port.lib.orc_bessel_k(n, float(x)) / kv(abs(n), x) - 1) < 2e-14, (n, x)


def test_force_energy_matches_golden(port):
    """0.5 sum |grad log psi|^2 + V: the restatement against forceEnergy::operator() of the reference (energy.cpp:37-61);
    Eigen's reduction order differs from a plain loop, hence 1e-13 instead of bit-equality"""
    G = np.load(os.path.join(GOLD, "force_energy.npz"))
    names = sorted({k.split("/")[0] for k in G.files})
    assert len(names) >= 3
    for name in names:
        cfg, _ = cases.eval_cases()[name]
        got = port.force_energy(cfg, G[name + "/R"])
        np.testing.assert_allclose(got, G[name + "/forceEnergy"], rtol=1e-13, atol=0)
        ev = port.eval(cfg, G[name + "/R"])
        np.testing.assert_allclose(got, 0.5 * (ev["grad"] ** 2).sum(axis=1) + ev["V"], rtol=1e-13)


def test_forward_walking_needs_steps_below_record_steps(port):
    cfg = cases.forward_walking_cases()["pbc_pair"]
    pops, ancs = cases.synthetic_genealogy(cfg, 3, 4, 1)
    with pytest.raises(Exception):
        port.forward_walking(cfg, pops, ancs, 0, 4, 4, 16)


def test_pair_histogram_lower_bound(port):
    """histogramAccumulator::accumulate (accumulators.h:117) has no lower bound: minx - dx < x < minx truncates to bin 0, anything
    below writes before the array in the reference.  The restatement drops those samples (and so do the CUDA kernels): checked
    against a numpy restatement of pairCorrelation.cpp:17-39 with that rule, on walkers that do contain pairs below minx"""
    N, L, bins, minx = 20, 3.0, 16, 0.6
    maxx = L / 2
    cfg = cases.pbc(N, L, walkers=8, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": bins, "minx": minx, "maxx": maxx}])
    R = np.random.default_rng(12).uniform(-L / 2, L / 2, size=(8, 3 * N))
    sums, w = port.histogram(cfg, R, 0, bins)
    dx = (maxx - minx) / bins
    nf = 2 * L ** 3 / (dx * 4 * np.pi * N * N)
    exp, below = np.zeros(bins), 0
    for Rw in R:
        x = Rw.reshape(3, N)
        for i in range(N):
            for j in range(i):
                d = x[:, i] - x[:, j]
                d -= L * np.round(d / L)
                r = np.sqrt((d * d).sum())
                if r < maxx:
                    b = int((r - minx) / dx)          # C truncation
                    if r < minx - dx:
                        below += 1
                    if b >= 0:
                        exp[min(b, bins - 1)] += nf / r ** 2
    assert below > 0 and w == len(R)
    np.testing.assert_allclose(sums, exp, rtol=1e-12)
