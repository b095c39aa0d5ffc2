"""TEST INFRASTRUCTURE — loaders for the two CPU checkers.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product
(qmc4_b200/) never does.

  ref  : oracle/_ref/libqmc4ref.so — the reference's own, unmodified dmc/*.cpp compiled in place
         (oracle/Makefile `ref`) behind the C shim oracle/ref_shim.cpp.
  port : oracle/liboracle.so — plain-C restatement of the hot path (oracle/qmc_oracle.c), pinned
         against `ref` and the golden vectors by tests/test_oracle.py.

Both take the reference's input.json text and walkers in the reference's per-walker layout
(species-major, then data[d*N_s + i]).
"""
import ctypes as C
import json
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libqmc4ref.so")
REF_BIN = os.path.join(_HERE, "_ref", "dmc_ref")
PORT_LIB = os.path.join(_HERE, "liboracle.so")

_dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
_ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))


def have_ref():
    return os.path.exists(REF_LIB)


def have_port():
    return os.path.exists(PORT_LIB)


def _dofs(cfg):
    return 3 * sum(cfg["N"])


class _Base:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        getattr(self.lib, prefix + "_last_error").restype = C.c_char_p

    def _ck(self, rc):
        if rc:
            raise RuntimeError(getattr(self.lib, self.prefix + "_last_error")().decode())

    def _fn(self, name):
        return getattr(self.lib, self.prefix + "_" + name)

    def eval(self, cfg, R):
        txt = json.dumps(cfg).encode()
        d = _dofs(cfg)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, d)
        nW = R.shape[0]
        out = dict(logpsi=np.zeros(nW), grad=np.zeros((nW, d)), lap=np.zeros(nW), V=np.zeros(nW), E=np.zeros(nW),
                   valid=np.zeros(nW, dtype=np.int32))
        self._ck(self._fn("eval")(txt, _dp(R), nW, _dp(out["logpsi"]), _dp(out["grad"]), _dp(out["lap"]), _dp(out["V"]),
                                  _dp(out["E"]), _ip(out["valid"])))
        return out

    def green_ratio(self, cfg, tau, Rnew, Fnew, Rold, Fold):
        txt = json.dumps(cfg).encode()
        d = _dofs(cfg)
        a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1, d) for x in (Rnew, Fnew, Rold, Fold)]
        nW = a[0].shape[0]
        q = np.zeros(nW)
        self._ck(self._fn("green_ratio")(txt, C.c_double(tau), nW, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(q)))
        return q

    def drift_diffuse(self, cfg, tau, Rold, Fold, xi_layout):
        """xi_layout: normal deviates in WALKER LAYOUT (species, d, i); reordered here to the
        reference's draw order (species, i, d) (dmc/moves/vmcMoves.cpp:22-33)."""
        txt = json.dumps(cfg).encode()
        d = _dofs(cfg)
        Rold = np.ascontiguousarray(Rold, dtype=np.float64).reshape(-1, d)
        Fold = np.ascontiguousarray(Fold, dtype=np.float64).reshape(-1, d)
        xi_layout = np.asarray(xi_layout, dtype=np.float64).reshape(-1, d)
        nW = Rold.shape[0]
        xi = np.zeros_like(xi_layout)
        off = 0
        for n in cfg["N"]:
            blk = xi_layout[:, off:off + 3 * n].reshape(nW, 3, n)
            xi[:, off:off + 3 * n] = blk.transpose(0, 2, 1).reshape(nW, 3 * n)
            off += 3 * n
        out = np.zeros_like(Rold)
        self._ck(self._fn("drift_diffuse")(txt, C.c_double(tau), nW, _dp(Rold), _dp(Fold), _dp(np.ascontiguousarray(xi)), _dp(out)))
        return out

    def histogram(self, cfg, R, iMeas, bins):
        txt = json.dumps(cfg).encode()
        d = _dofs(cfg)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, d)
        sums = np.zeros(bins)
        w = C.c_double(0)
        with tempfile.TemporaryDirectory() as tmp:
            self._ck(self._fn("histogram")(txt, tmp.encode(), _dp(R), R.shape[0], iMeas, _dp(sums), bins, C.byref(w)))
        return sums, w.value

    def forward_walking(self, cfg, pops, ancs, iEst, bins):
        """pops: list of [n_e, dofs] arrays (one per sampling event); ancs: list of int arrays (index into the
        previous event's population; ancs[0] ignored); iEst: index in the vector estimatorFactory::create returns"""
        txt = json.dumps(cfg).encode()
        d = _dofs(cfg)
        n = np.asarray([len(p) for p in pops], dtype=np.int32)
        R = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, d) for p in pops]))
        anc = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.int32) for a in ancs]))
        assert len(anc) == len(R)
        sums = np.zeros(bins)
        w = C.c_double(0)
        with tempfile.TemporaryDirectory() as tmp:
            self._ck(self._fn("forward_walking")(txt, tmp.encode(), _dp(R), _ip(n), _ip(anc), len(n), iEst, _dp(sums), bins, C.byref(w)))
        return sums, w.value

    def force_energy(self, cfg, R):
        """forceEnergy::operator() per configuration (energy.cpp:37-61)"""
        txt = json.dumps(cfg).encode()
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, _dofs(cfg))
        out = np.zeros(R.shape[0])
        self._ck(self._fn("force_energy")(txt, _dp(R), R.shape[0], _dp(out)))
        return out

    def superfluid_fraction(self, cfg, pops, ancs, iEst, bins):
        """the reference's centerOfMassStorer + superfluidFractionEstimator over a genealogy -> (sums, weights)"""
        txt = json.dumps(cfg).encode()
        d = _dofs(cfg)
        n = np.asarray([len(p) for p in pops], dtype=np.int32)
        R = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, d) for p in pops]))
        anc = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.int32) for a in ancs]))
        assert len(anc) == len(R)
        sums, ws = np.zeros(bins), np.zeros(bins)
        with tempfile.TemporaryDirectory() as tmp:
            self._ck(self._fn("superfluid_fraction")(txt, tmp.encode(), _dp(R), _ip(n), _ip(anc), len(n), iEst, _dp(sums), _dp(ws), bins))
        return sums, ws

    def plan(self, populations):
        p = np.asarray(populations, dtype=np.int32)
        src = np.zeros(len(p), dtype=np.int32)
        amt = np.zeros(len(p), dtype=np.int32)
        self._ck(self._fn("plan")(_ip(p), len(p), _ip(src), _ip(amt)))
        return src, amt

    def branch(self, tau, mean, delta, seed, Enew, Eold):
        Enew = np.ascontiguousarray(Enew, dtype=np.float64)
        Eold = np.ascontiguousarray(Eold, dtype=np.float64)
        n = len(Enew)
        cap = 8 * n + 64
        eT, eT2 = C.c_double(0), C.c_double(0)
        nd = np.zeros(n, dtype=np.int32)
        u = np.zeros(n)
        newpop = C.c_int(0)
        Eafter = np.zeros(cap)
        self._ck(self._fn("branch")(C.c_double(tau), C.c_double(mean), C.c_double(delta), int(seed), n, _dp(Enew), _dp(Eold),
                                    C.byref(eT), _ip(nd), _dp(u), C.byref(newpop), _dp(Eafter), cap, C.byref(eT2)))
        return dict(energyShift=eT.value, nDesc=nd, u=u, newPop=newpop.value, Eafter=Eafter[:newpop.value],
                    energyShiftAfter=eT2.value)


class Ref(_Base):
    def __init__(self):
        if not have_ref():
            raise FileNotFoundError(REF_LIB + " — run `make -C oracle ref` where /root/reference exists")
        super().__init__(REF_LIB, "ref")


class _Model(C.Structure):
    _fields_ = [("lBox", C.c_double * 3), ("nSpecies", C.c_int), ("N", C.POINTER(C.c_int)), ("nTerms", C.c_int),
                ("ti", C.POINTER(C.c_int)), ("tp", C.POINTER(C.c_double))]


_J = {"gaussian": (1, ("alpha",)), "gaussiangaussian": (2, ("r0", "u0")),
      "squareWell": (3, ("V0", "R0", "Rm", "alpha", "aInverse", "cut_off")), "hardSphereGauss": (4, ("a", "Rm", "D")),
      "bSpline": (5, ("stepSize", "valueRight", "derivativeRight")), "poschTeller": (6, ("R0", "C", "alpha", "Rm", "cut_off")),
      "hardSphere": (7, ("a", "k", "cut_off")),
      "squareWellBoundState": (8, ("V0", "R0", "Rm", "cut_off", "aInverse", "eb", "alpha")),
      "delta_phonons": (9, ("k", "delta", "beta", "z", "cut_off")), "delta_in_trap": (10, ("a",)),
      "delta_bound_state_phonons": (11, ("xI", "k", "A", "beta", "cut_off")),
      "dipolar_rep": (12, ("D", "cut_off", "matching_point", "alpha", "C"))}
_WF = {"jastrow1b": 1, "jastrow2bUnDis": 2, "jastrow2bDis": 3, "jastrow3bUnDis": 4, "jastrow3bDis": 5}
_POT = {"harmonic1b": 1, "squareWell2b": 2, "dipolar2b": 3, "poschTeller2b": 4, "potentialTwoBody/squareWell": 5,
        "potentialTwoBody/potentialBarrier": 6, "potentialThreeBodyUnDis/squareWell": 7,
        "potentialThreeBodyUnDis/potentialBarrier": 8, "potentialThreeBodyDis/squareWell": 9,
        "potentialThreeBodyDis/potentialBarrier": 10}
_EST = {"pairCorrelation": 1, "trimerCorrelationUnDis": 2, "structureFactor": 3, "trimerCorrelation": 4,
        "trimerCorrelationDis": 5, "trimerhyperRadius": 6, "centerOfMassSquared": 7}
SCALAR_OBSERVABLES = ("trimerhyperRadius", "centerOfMassSquared")


def flatten(cfg):
    """input.json (dict) -> flat term list of qmc_oracle.h, in the reference's term order."""
    ti, tp = [], []

    def add(cls, kind, sets, jk, params):
        sets = (list(sets) + [0, 0, 0])[:3]
        ti.extend([cls, kind] + sets + [jk, len(params), len(tp)])
        tp.extend(float(x) for x in params)

    for w in cfg.get("wavefunctions", []):
        jname = w["jastrow"]["kind"]
        jk, keys = _J[jname]
        params = [w["jastrow"][k] for k in keys]
        if jname == "bSpline":
            params += list(w["jastrow"]["coefficients"])
        kind = _WF[w["kind"]]
        sets = [w["set"]] if w["kind"] == "jastrow1b" else w["sets"]
        if kind in (4, 5):
            params.append(w.get("hardSphereRadius", -1.0))
        add(1, kind, sets, jk, params)
    for p in cfg.get("potentials", []):
        name = p["kind"] + ("/" + p["functor"]["kind"] if "functor" in p else "")
        kind = _POT[name]
        if kind == 1:
            add(2, kind, [p["set"]], 0, [p["omega"]])
        elif kind == 2:
            add(2, kind, p["sets"], 0, [p["V0"], p["R0"]])
        elif kind == 3:
            add(2, kind, p["sets"], 0, [p["D"]])
        elif kind == 4:
            add(2, kind, p["sets"], 0, [p["R0"]])
        else:
            add(2, kind, p["sets"], 0, [p["functor"]["R0"], p["functor"]["V0"]])
    for e in cfg.get("measurements", []):
        if e["kind"] in ("energy", "forceEnergy", "superfluidFraction"):      # no observable term: see Port.superfluid_fraction
            continue
        kind = _EST[e["kind"]]
        if kind == 3:
            add(3, kind, e["sets"], e["bins"], [e["maxq"], e["lBox"], e["channel"]])
        elif kind in (6, 7):      # scalar observables: one "bin" holding the sum
            add(3, kind, [e["set"]] if kind == 7 else e["sets"], 1, [0.0, 1.0])
        else:
            add(3, kind, e["sets"], e["bins"], [e["minx"], e["maxx"]])
    return np.asarray(ti, dtype=np.int32), np.asarray(tp if tp else [0.0], dtype=np.float64)


class Port(_Base):
    """Plain-C restatement (oracle/qmc_oracle.c)."""

    def __init__(self):
        if not have_port():
            raise FileNotFoundError(PORT_LIB + " — run `make -C oracle port`")
        super().__init__(PORT_LIB, "orc")
        self.lib.orc_energy_shift.restype = C.c_double
        self.lib.orc_energy_shift.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double, C.c_double]
        self.lib.orc_n_descendants.argtypes = [C.c_double] * 5

    def _model(self, cfg):
        ti, tp = flatten(cfg)
        N = np.asarray(cfg["N"], dtype=np.int32)
        m = _Model()
        for d in range(3):
            m.lBox[d] = cfg["lBox"][d]
        m.nSpecies = len(N)
        m.N = _ip(N)
        m.nTerms = len(ti) // 8
        m.ti = _ip(ti)
        m.tp = _dp(tp)
        m._keep = (ti, tp, N)
        return m

    def eval(self, cfg, R):
        m = self._model(cfg)
        d = _dofs(cfg)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, d)
        nW = R.shape[0]
        out = dict(logpsi=np.zeros(nW), grad=np.zeros((nW, d)), lap=np.zeros(nW), V=np.zeros(nW), E=np.zeros(nW),
                   valid=np.zeros(nW, dtype=np.int32))
        self._ck(self.lib.orc_eval_model(C.byref(m), _dp(R), nW, _dp(out["logpsi"]), _dp(out["grad"]), _dp(out["lap"]),
                                         _dp(out["V"]), _dp(out["E"]), _ip(out["valid"])))
        return out

    def green_ratio(self, cfg, tau, Rnew, Fnew, Rold, Fold):
        m = self._model(cfg)
        d = _dofs(cfg)
        a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1, d) for x in (Rnew, Fnew, Rold, Fold)]
        q = np.zeros(a[0].shape[0])
        self._ck(self.lib.orc_green_ratio_model(C.byref(m), C.c_double(tau), len(q), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(q)))
        return q

    @staticmethod
    def draw_order(cfg, xi_layout):
        d = _dofs(cfg)
        xi_layout = np.asarray(xi_layout, dtype=np.float64).reshape(-1, d)
        nW = xi_layout.shape[0]
        xi = np.zeros_like(xi_layout)
        off = 0
        for n in cfg["N"]:
            blk = xi_layout[:, off:off + 3 * n].reshape(nW, 3, n)
            xi[:, off:off + 3 * n] = blk.transpose(0, 2, 1).reshape(nW, 3 * n)
            off += 3 * n
        return np.ascontiguousarray(xi)

    def drift_diffuse(self, cfg, tau, Rold, Fold, xi_layout):
        m = self._model(cfg)
        d = _dofs(cfg)
        Rold = np.ascontiguousarray(Rold, dtype=np.float64).reshape(-1, d)
        Fold = np.ascontiguousarray(Fold, dtype=np.float64).reshape(-1, d)
        xi = self.draw_order(cfg, xi_layout)
        out = np.zeros_like(Rold)
        self._ck(self.lib.orc_drift_diffuse_model(C.byref(m), C.c_double(tau), Rold.shape[0], _dp(Rold), _dp(Fold), _dp(xi), _dp(out)))
        return out

    def update(self, cfg, tau, Rold, xi_layout, u_m):
        """dmcDriver::update for every walker given the random numbers it consumes."""
        m = self._model(cfg)
        d = _dofs(cfg)
        Rold = np.ascontiguousarray(Rold, dtype=np.float64).reshape(-1, d)
        nW = Rold.shape[0]
        xi = self.draw_order(cfg, xi_layout)
        u_m = np.ascontiguousarray(u_m, dtype=np.float64)
        out = dict(R=np.zeros((nW, d)), F=np.zeros((nW, d)), logpsi=np.zeros(nW), E=np.zeros(nW), accepted=np.zeros(nW, dtype=np.int32))
        self._ck(self.lib.orc_update_model(C.byref(m), C.c_double(tau), nW, _dp(Rold), _dp(xi), _dp(u_m), _dp(out["R"]), _dp(out["F"]),
                                           _dp(out["logpsi"]), _dp(out["E"]), _ip(out["accepted"])))
        return out

    def tabulate(self, jastrow, minx, maxx, n):
        """jastrow<T>::print for the radial functor given as the input's "jastrow" object"""
        jk, keys = _J[jastrow["kind"]]
        params = [jastrow[k] for k in keys] + (list(jastrow["coefficients"]) if jastrow["kind"] == "bSpline" else [])
        p = np.asarray(params, dtype=np.float64)
        out = np.zeros((int(n), 4))
        self._ck(self.lib.orc_tabulate(jk, _dp(p), len(p), C.c_double(minx), C.c_double(maxx), int(n), _dp(out)))
        return out

    def vmc_step(self, cfg, sigma, Rold, xi_layout, u_m):
        """vmcDriver::step for every chain given the random numbers it consumes."""
        m = self._model(cfg)
        d = _dofs(cfg)
        Rold = np.ascontiguousarray(Rold, dtype=np.float64).reshape(-1, d)
        nW = Rold.shape[0]
        xi = self.draw_order(cfg, xi_layout)
        u_m = np.ascontiguousarray(u_m, dtype=np.float64)
        out = dict(R=np.zeros((nW, d)), logpsi=np.zeros(nW), E=np.zeros(nW), accepted=np.zeros(nW, dtype=np.int32))
        self._ck(self.lib.orc_vmc_step_model(C.byref(m), C.c_double(sigma), nW, _dp(Rold), _dp(xi), _dp(u_m), _dp(out["R"]),
                                             _dp(out["logpsi"]), _dp(out["E"]), _ip(out["accepted"])))
        return out

    def histogram(self, cfg, R, iMeas, bins):
        m = self._model(cfg)
        d = _dofs(cfg)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, d)
        sums = np.zeros(bins)
        w = C.c_double(0)
        self._ck(self.lib.orc_histogram_model(C.byref(m), _dp(R), R.shape[0], iMeas, _dp(sums), bins, C.byref(w)))
        return sums, w.value

    def forward_walking(self, cfg, pops, ancs, iHist, recordSteps, fwSteps, bins):
        """iHist: index among the measurements of the input (the target histogram estimator)"""
        m = self._model(cfg)
        d = _dofs(cfg)
        n = np.asarray([len(p) for p in pops], dtype=np.int32)
        R = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, d) for p in pops]))
        anc = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.int32) for a in ancs]))
        assert len(anc) == len(R)
        sums = np.zeros(bins)
        w = C.c_double(0)
        self._ck(self.lib.orc_forward_walking_model(C.byref(m), _dp(R), _ip(n), _ip(anc), len(n), iHist, int(recordSteps),
                                                    int(fwSteps), _dp(sums), bins, C.byref(w)))
        return sums, w.value

    def force_energy(self, cfg, R):
        m = self._model(cfg)
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, _dofs(cfg))
        out = np.zeros(R.shape[0])
        self._ck(self.lib.orc_force_energy_model(C.byref(m), _dp(R), R.shape[0], _dp(out)))
        return out

    def superfluid_fraction(self, cfg, pops, ancs, iMeas):
        """iMeas: index among the measurements of the input (a "superfluidFraction" one) -> (sums, weights)"""
        m = self._model(cfg)
        d = _dofs(cfg)
        sf = [x for x in cfg["measurements"] if x["kind"] == "superfluidFraction"]
        stored = []
        for x in sf:                                          # createSuperfluidFractionStorer: sets in order of first appearance
            for s_ in x["sets"]:
                if s_ not in stored:
                    stored.append(s_)
        meas = cfg["measurements"][iMeas]
        bins = int(meas["bins"])
        n = np.asarray([len(p) for p in pops], dtype=np.int32)
        R = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, d) for p in pops]))
        anc = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.int32) for a in ancs]))
        st = np.asarray(stored, dtype=np.int32)
        sums, ws = np.zeros(bins), np.zeros(bins)
        self._ck(self.lib.orc_superfluid_fraction_model(C.byref(m), _dp(R), _ip(n), _ip(anc), len(n), _ip(st), len(st), int(sf[0]["bins"]),
                                                        stored.index(meas["sets"][0]), stored.index(meas["sets"][1]), _dp(sums), _dp(ws), bins))
        return sums, ws

    def plan(self, populations):
        p = np.asarray(populations, dtype=np.int32)
        src = np.zeros(len(p), dtype=np.int32)
        amt = np.zeros(len(p), dtype=np.int32)
        self._ck(self.lib.orc_plan(_ip(p), len(p), _ip(src), _ip(amt)))
        return src, amt

    def energy_shift(self, E, tau, mean, delta):
        E = np.ascontiguousarray(E, dtype=np.float64)
        return self.lib.orc_energy_shift(_dp(E), len(E), tau, mean, delta)

    def n_descendants(self, Enew, Eold, eT, tau, u):
        return np.asarray([self.lib.orc_n_descendants(float(a), float(b), float(eT), float(tau), float(c))
                           for a, b, c in zip(Enew, Eold, u)], dtype=np.int32)

    def branch(self, nDesc, E):
        nDesc = np.ascontiguousarray(nDesc, dtype=np.int32)
        E = np.ascontiguousarray(E, dtype=np.float64)
        cap = int(max(nDesc.sum(), len(E))) + 8
        out = np.zeros(cap)
        n = self.lib.orc_branch(_ip(nDesc), _dp(E), len(E), _dp(out), cap)
        if n < 0:
            raise RuntimeError("orc_branch: capacity")
        return out[:n]
