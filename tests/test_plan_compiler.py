"""CPU tests (-m "not gpu") of the host logic behind the C ABI: the term-list compiler that turns the objects
the reference's factories would create into the POD kernel plan (qmc4_b200/csrc/plan.h, build_plan in
capi.cu), through a planning-only context (device -1: no CUDA call, no compute)."""
import math

import numpy as np
import pytest

import cases
from qmc4_b200.system import System, invalidInput

J_NONE, J_GAUSSIAN, J_GAUSSGAUSS = 0, 1, 2
V_NONE, V_STEP, V_DIPOLAR = 0, 1, 2
KSEL_GG_STEP, KSEL_GG_NONE, KSEL_GAUSS_NONE, KSEL_GENERIC = 0, 1, 2, 3


def plan(cfg):
    return System.from_input(cfg, device=-1)


def test_planning_context_cannot_compute():
    s = plan(cases.pbc(8, 4.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.eval(np.zeros((1, 24)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.upload(np.zeros((1, 24)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.step(1)
    s.close()


def test_config2_is_one_fused_pass_on_the_specialised_kernel():
    L = 4.0
    cfg = cases.pbc(64, L, walkers=16384, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 1000, "minx": 0, "maxx": L / 2}])
    s = plan(cfg)
    assert s.info("n_pair_passes") == 1 and s.info("n_one_body") == 0 and s.info("n_triplet_passes") == 0
    assert s.info("kernel_variant") == KSEL_GG_STEP and s.info("warps_per_walker") == 1 and s.info("groups_per_cta") == 8
    assert (s.info("pass0.jkind"), s.info("pass0.vkind"), s.info("pass0.hist")) == (J_GAUSSGAUSS, V_STEP, 0)
    assert s.info("pass0.v_val") == 2.0                                   # potentialBarrier: +V0 inside R0
    # bit-exact threshold: r2 <= t  <=>  sqrt(r2) <= R0 (potentialFunctors.h:26 compares the rounded norm)
    t = s.info("pass0.v_thr2")
    assert math.sqrt(t) <= 1.0 < math.sqrt(np.nextafter(t, np.inf))
    # pairCorrelation normalisation (pairCorrelation.cpp:41-59) and histogram geometry (accumulators.h:108)
    dx = (L / 2) / 1000
    assert s.info("est0.norm") == pytest.approx(2 * L ** 3 / (dx * 4 * math.pi * 64 * 64), rel=1e-15)
    assert s.info("est0.dxinv") == pytest.approx(1 / dx, rel=1e-15)
    assert s.info("est0.fx_scale") == 2.0 ** 36 * (L / 2) ** 2
    assert s.info("hist_stride") == 1000
    assert s.info("capacity") == 2 * (16384 + 1638) + 64
    # smem: 8 walkers x (R + one accumulator + reduction slots) + the per-CTA fixed-point histogram
    assert s.info("smem_bytes") == 8 * 8 * (192 * 2 + 12) + 1000 * 8
    s.close()


def test_gaussgauss_exp_constants():
    """exp(-alpha r^2) is evaluated as 2^n * P(r2 - n*c), c = ln2/(-alpha) split hi/lo, P with coefficients T_j(-alpha)^j"""
    r0, u0 = 1.3, -0.5
    s = plan(cases.pbc(8, 4.0, jastrow={"kind": "gaussiangaussian", "r0": r0, "u0": u0}))
    alpha = 1 / (2 * r0 * r0)                                             # jastrowGaussGauss.cpp:5-10
    p = [s.info("pass0.p%d" % i) for i in range(15)]
    assert p[0] == alpha and p[1] == u0
    assert p[2] == pytest.approx(-alpha * math.log2(math.e), rel=1e-15)
    c = math.log(2) / (-alpha)
    assert p[3] + p[4] == pytest.approx(c, rel=1e-15)
    assert np.float64(p[3]).view(np.uint64) & np.uint64(0xFFFFFFFF) == 0          # n*hi is exact for |n| < 2^31
    assert abs(p[4]) < abs(p[3]) * 2.0 ** -20
    T = [1.0, 1.0, 0.5, 0.16666666666648303, 0.041666666666651364, 0.008333333353717156, 0.0013888888905871347,
         0.0001984120875699232, 2.4801536409064087e-05, 2.7625102005388108e-06, 2.761379555451986e-07]
    for j in range(1, 11):
        assert p[4 + j] == pytest.approx(T[j] * (-alpha) ** j, rel=1e-15)
    # the polynomial itself: |P(f) - exp(f)| / exp(f) < 2e-15 on the reduced range
    f = np.linspace(-math.log(2) / 2, math.log(2) / 2, 2001)
    P = np.zeros_like(f)
    for j in range(10, -1, -1):
        P = P * f + T[j]
    assert np.abs(P / np.exp(f) - 1).max() < 2e-15
    s.close()


def test_trap_uses_gaussian_kernel_and_merges_the_one_body_potential():
    s = plan(cases.trap(10))
    assert s.info("kernel_variant") == KSEL_GAUSS_NONE and s.info("n_pair_passes") == 1 and s.info("n_one_body") == 1
    assert (s.info("onebody0.jkind"), s.info("onebody0.has_pot"), s.info("onebody0.half_omega2")) == (J_GAUSSIAN, 1, 0.5)
    assert s.info("pass0.p0") == 0.05 and s.info("pass0.vkind") == V_NONE and s.info("pass0.hist") == -1
    s.close()


def test_mixture_plan():
    s = plan(cases.mixture())
    assert s.info("n_pair_passes") == 3 and s.info("n_one_body") == 1 and s.info("n_estimators") == 2
    # a specialised pass is moved first: (0,0) gaussiangaussian + squareWell2b (-V0) + g00
    assert (s.info("pass0.setA"), s.info("pass0.setB"), s.info("pass0.jkind"), s.info("pass0.vkind")) == (0, 0, J_GAUSSGAUSS, V_STEP)
    assert s.info("pass0.v_val") == -1.5 and s.info("pass0.hist") == 1
    assert (s.info("pass1.setA"), s.info("pass1.setB"), s.info("pass1.vkind"), s.info("pass1.v_val")) == (1, 1, V_DIPOLAR, 0.3)
    assert (s.info("pass2.setA"), s.info("pass2.setB"), s.info("pass2.vkind"), s.info("pass2.v_val"), s.info("pass2.hist")) == (0, 1, V_STEP, -2.0, 0)
    # distinguishable g(r): no factor 2 (pairCorrelation.cpp:56)
    dx = 2.5 / 40
    assert s.info("est0.norm") == pytest.approx(5.0 ** 3 / (dx * 4 * math.pi * 12 * 20), rel=1e-15)
    assert s.info("est1.offset") == 40 and s.info("hist_stride") == 70
    s.close()


def test_potential_or_estimator_without_a_wavefunction_on_the_same_sets_gets_its_own_pass():
    cfg = cases.trap(6)
    cfg["potentials"].append({"kind": "potentialTwoBody", "sets": [0, 0], "functor": {"kind": "squareWell", "R0": 0.7, "V0": 3.0}})
    cfg["wavefunctions"] = cfg["wavefunctions"][:1]            # only the one-body Jastrow
    cfg["measurements"] = [{"kind": "pairCorrelation", "label": "g", "sets": [0, 0], "bins": 10, "minx": 0, "maxx": 1.0}]
    s = plan(cfg)
    assert s.info("n_pair_passes") == 1 and s.info("kernel_variant") == KSEL_GENERIC
    assert (s.info("pass0.jkind"), s.info("pass0.vkind"), s.info("pass0.v_val"), s.info("pass0.hist")) == (J_NONE, V_STEP, -3.0, 0)
    s.close()


def test_three_body_plan():
    s = plan(cases.three_body(32, hs=0.4))
    # undistinguishable triplets on <= 32 particles: the cyclic schedule (every triplet once), two warps per walker
    assert s.info("n_triplet_passes") == 1 and s.info("triplet_cyclic") == 1 and s.info("warps_per_walker") == 2
    assert (s.info("triplet0.jkind"), s.info("triplet0.has_pot"), s.info("triplet0.hist"), s.info("triplet0.has_constraint")) == (J_GAUSSGAUSS, 1, 0, 1)
    assert s.info("triplet0.v_val") == -1.0                                # squareWell functor: -V0
    N = 32
    assert s.info("est0.norm") == pytest.approx(1 / ((2.5 / 50) * 4 * math.pi * N ** 3), rel=1e-15)     # trimerCorrelation.cpp:128-143
    tables = 1024 * (1 + 2)                 # r2[32][32] + one pair-weight table per warp
    groups = 4                              # walkers per 256-thread CTA
    assert s.info("smem_bytes") == groups * 8 * (96 * 3 + 16 + tables) + 50 * 8
    s.close()
    # the three-visit pass remains for larger sets
    s = plan(cases.three_body(40, L=5.5))
    assert s.info("triplet_cyclic") == 0 and s.info("warps_per_walker") == 8
    s.close()


def test_structure_factor_grid():
    L = 4.0
    s = plan(cases.structure_factor(24, L))
    q0 = 2 * math.pi / L
    assert s.info("est0.q0") == pytest.approx(q0, rel=1e-15)
    dq = int((16 * 2 * math.pi / L) / (2 * math.pi * 16 / L)) * 2 * math.pi / L       # structureFactor.cpp:8-21
    assert s.info("est0.dq") == pytest.approx(dq if dq else q0, rel=1e-15)
    s.close()


@pytest.mark.parametrize("N,wpw", [(10, 1), (64, 1), (65, 2), (128, 2), (256, 4), (257, 8), (512, 8)])
def test_warps_per_walker_follow_the_tile_count(N, wpw):
    s = plan(cases.pbc(N, 6.0))
    assert s.info("warps_per_walker") == wpw and s.info("groups_per_cta") == 8 // wpw
    s.close()


def test_work_decomposition_override(monkeypatch):
    monkeypatch.setenv("QMC4_WPW", "4")
    s = plan(cases.pbc(64, 4.0))
    assert s.info("warps_per_walker") == 4
    s.close()


def test_population_control_constants():
    s = plan(cases.pbc(8, 4.0, walkers=1000, timeStep=2e-3))
    assert s.info("log_up") == pytest.approx(1 / 2e-3 * math.log(1000 / 1100), rel=1e-15)     # branching.cpp:84-96
    assert s.info("log_down") == pytest.approx(1 / 2e-3 * math.log(1000 / 900), rel=1e-15)
    s.close()


def test_derived_jastrow_coefficients_follow_the_reference_constructors():
    J = cases.JASTROWS["squareWell"]
    s = plan(cases.pbc(8, 5.0, jastrow=J))
    V0, R0, Rm, alpha, aInv, lBox = J["V0"], J["R0"], J["Rm"], J["alpha"], J["aInverse"], 2 * J["cut_off"]
    K0 = math.sqrt(V0)
    A = math.sin(K0 * R0) / (1 - R0 * aInv)                                # jastrowSquareWell.cpp:5-39
    C = -A / (Rm * Rm * alpha * (math.exp(-alpha * (lBox - Rm)) - math.exp(-alpha * Rm)))
    B = A * (1 / Rm - aInv) - C * (math.exp(-alpha * Rm) + math.exp(-alpha * (lBox - Rm)))
    got = [s.info("pass0.p%d" % i) for i in range(10)]
    want = [R0, Rm, lBox / 2, K0, aInv, alpha, lBox, B, C, math.log(B + 2 * C * math.exp(-alpha * lBox / 2))]
    assert got == pytest.approx(want, rel=1e-14)
    s.close()
    J = cases.JASTROWS["hardSphereGauss"]
    s = plan(cases.pbc(8, 5.0, jastrow=J))
    a, Rm, d = J["a"], J["Rm"], J["D"]
    al = a / (Rm * (Rm - a) * 2 * (d - Rm))                                # jastrowHardSphereGauss.cpp:4-10
    assert [s.info("pass0.p%d" % i) for i in range(5)] == pytest.approx([Rm, a, d, al, math.log((1 - a / Rm) * math.exp(al * (Rm - d) ** 2))], rel=1e-14)
    s.close()
    J = cases.JASTROWS["bSpline"]
    s = plan(cases.pbc(8, 5.0, jastrow=J))
    assert s.info("pass0.p0") == 1 / J["stepSize"] and s.info("pass0.p1") == (len(J["coefficients"]) - 3) * J["stepSize"]   # jastrowSpline.cpp:72
    s.close()


def test_errors():
    cfg = cases.pbc(8, 4.0)
    cfg["wavefunctions"][0]["sets"] = [0, 3]
    with pytest.raises((invalidInput, RuntimeError)):
        plan(cfg)
    cfg = cases.pbc(8, 4.0)
    cfg["potentials"][0]["sets"] = [0, 2]
    with pytest.raises(RuntimeError, match="invalid set"):
        plan(cfg)
    cfg = cases.pbc(8, 4.0)
    cfg["wavefunctions"] = [dict(cfg["wavefunctions"][0]) for _ in range(5)]
    with pytest.raises(RuntimeError, match="too many pair passes"):
        plan(cfg)
    with pytest.raises(RuntimeError, match="shared memory"):
        plan(cases.pbc(4000, 20.0))
    with pytest.raises(RuntimeError, match="Jastrow cannot be negative"):       # jastrowSquareWell.cpp:29-33 (the reference exits)
        plan(cases.pbc(8, 5.0, jastrow={"kind": "squareWell", "V0": 2.4674011002723395, "R0": 1.0, "Rm": 1.6, "alpha": 0.5, "aInverse": 0.7, "cut_off": 2.5}))


def test_superfluid_force_energy_and_dipolar_rep_are_planned_and_validated_without_a_device():
    """the round-1 continuation's host logic: measurement wiring, derived dipolar_rep constants, error semantics"""
    # superfluidFraction: the record ring shares the forward-walking event clock; link depth = max(max fw + 1, bins)
    s = plan(cases.superfluid_cases()["mixture_rho"])
    assert s.info("fw_link_depth") == 5 and [e["label"] for e in s.estimators] == ["rho11", "g01", "rho01", "rho00"]
    assert all(e.get("vector") for e in s.estimators if e["kind"] == "superfluidFraction")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.estimator(s.estimators[0]["index"])
    s.close()
    fw = cases.forward_walking_cases()["pbc_pair"]
    fw["measurements"].append({"kind": "superfluidFraction", "label": "rho", "sets": [0, 0], "bins": 3})
    s = plan(fw)
    assert s.info("fw_depth") == 4 and s.info("fw_link_depth") == 4       # max fw = 3 -> 4 events of rows and links
    s.close()
    bad = cases.superfluid_cases()["mixture_rho"]
    bad["measurements"][2]["bins"] = 7
    with pytest.raises(RuntimeError, match="same .bins"):                   # correlationEstimator.cpp:333: record size = data.size()/bins
        plan(bad)
    bad = cases.superfluid_cases()["pbc_rho"]
    bad["measurements"][0]["sets"] = [0, 1]
    with pytest.raises(RuntimeError, match="set out of range"):
        plan(bad)
    with pytest.raises(RuntimeError, match="vmc and superfluid"):           # main.cpp:279-282
        plan(dict(cases.superfluid_cases()["pbc_rho"], method="vmc"))
    # forceEnergy: attached for vmc / svmc only (main.cpp:318-335,395-408)
    cfg = cases.pbc(8, 4.0)
    cfg["measurements"] = [{"kind": "forceEnergy", "label": "forceEnergy"}]
    for method, on in (("dmc", False), ("svmc", True), ("vmc", True)):
        s = plan(dict(cfg, method=method))
        assert s.force_energy_enabled is on
        s.close()
    # dipolar_rep: D = sqrt(j["D"]), lBox = 2 cut_off, k = pi / lBox, logC = log(C) (jastrowDipolar.cpp:3-27); generic kernel
    s = plan(cases.pbc(10, 5.2, jastrow=cases.DIPOLAR_REP))
    J = cases.DIPOLAR_REP
    assert s.info("kernel_variant") == KSEL_GENERIC and s.info("pass0.jkind") == 12
    got = [s.info("pass0.p%d" % i) for i in range(6)]
    exp = [math.sqrt(J["D"]), J["cut_off"], J["matching_point"], J["alpha"], math.pi / (2 * J["cut_off"]), math.log(J["C"])]
    assert got == exp
    s.close()
