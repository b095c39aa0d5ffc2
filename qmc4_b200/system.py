"""Host-side mirror of the reference's factory layer for tests and bench.py.

`System.from_input(j)` takes the reference's unchanged input.json document (a dict) and issues one
C-ABI call per object the reference factories would create, using the same id scheme
(`createId`, dmc/abstractFactory.cpp:3-20: own "kind" + "/" + ids of structured children in
key-sorted order) and the same error types as dmc/qmcExceptions.h.  The production host driver is
the C++ one under qmc4_b200/host/ (same mapping); this module is the Python plumbing around the
same C ABI.  It never computes physics.
"""
import ctypes as C
import json

import numpy as np

from . import capi


class invalidInput(Exception):
    pass


class factoryIdNotRecorded(Exception):
    pass


class missingKey(Exception):
    pass


def create_id(j):
    """createId (dmc/abstractFactory.cpp:3-20); nlohmann objects iterate in key-sorted order."""
    ident = j.get("kind", "") if isinstance(j, dict) else ""
    items = [j[k] for k in sorted(j)] if isinstance(j, dict) else list(j)
    for el in items:
        if isinstance(el, (dict, list)):
            sub = create_id(el)
            if sub != "":
                ident += "/" + sub
    return ident


def _get(j, key):
    if key not in j:
        raise missingKey(key)
    return j[key]


def _delta_phonons(j):
    """jastrow_delta.cpp:5-39: k, delta, beta, z, cut_off are read in this order, then g, which must be a number or the
    string "inf" (it only matters to process(), which the reference never calls)"""
    p = [_get(j, k) for k in ("k", "delta", "beta", "z", "cut_off")]
    g = _get(j, "g")
    if isinstance(g, str) and g != "inf":
        raise invalidInput("g is not a number and is not infinite")
    return p


def _delta_bound_state(j):
    """jastrow_delta.cpp:76-86 reads xI, g, cut_off, k, A, beta (all required; g is stored, never used)"""
    v = {k: _get(j, k) for k in ("xI", "g", "cut_off", "k", "A", "beta")}
    return [v["xI"], v["k"], v["A"], v["beta"], v["cut_off"]]


# radial functor name -> (C kind, parameter extractor) ; mirrors the JSON constructors
_JASTROWS = {
    "gaussian": (capi.J_GAUSSIAN, lambda j: [_get(j, "alpha")]),
    "gaussiangaussian": (capi.J_GAUSSGAUSS, lambda j: [_get(j, "r0"), _get(j, "u0")]),
    "squareWell": (capi.J_SQUAREWELL, lambda j: [_get(j, k) for k in ("V0", "R0", "Rm", "alpha", "aInverse", "cut_off")]),
    "hardSphereGauss": (capi.J_HARDSPHEREGAUSS, lambda j: [_get(j, k) for k in ("a", "Rm", "D")]),
    "bSpline": (capi.J_BSPLINE, lambda j: [_get(j, "stepSize"), _get(j, "valueRight"), _get(j, "derivativeRight")] + list(_get(j, "coefficients"))),
    "poschTeller": (capi.J_POSCHTELLER, lambda j: [_get(j, k) for k in ("R0", "C", "alpha", "Rm", "cut_off")]),
    "hardSphere": (capi.J_HARDSPHERE, lambda j: [_get(j, k) for k in ("a", "k", "cut_off")]),
    "squareWellBoundState": (capi.J_SQUAREWELL_BOUND, lambda j: [_get(j, k) for k in ("V0", "R0", "Rm", "cut_off", "aInverse", "eb", "alpha")]),
    "delta_phonons": (capi.J_DELTA_PHONONS, lambda j: _delta_phonons(j)),
    "delta_in_trap": (capi.J_DELTA_IN_TRAP, lambda j: [_get(j, "a")]),
    "dipolar_rep": (capi.J_DIPOLAR_REP, lambda j: [_get(j, k) for k in ("D", "cut_off", "matching_point", "alpha", "C")]),   # jastrowDipolar.cpp:3-13
    "delta_bound_state_phonons": (capi.J_DELTA_BOUND_STATE_PHONONS, lambda j: _delta_bound_state(j)),
}
_WAVES = {"jastrow1b": capi.WF_JASTROW1B, "jastrow2bUnDis": capi.WF_JASTROW2B_UNDIS,
          "jastrow2bDis": capi.WF_JASTROW2B_DIS, "jastrow3bUnDis": capi.WF_JASTROW3B_UNDIS,
          "jastrow3bDis": capi.WF_JASTROW3B_DIS}
_POTENTIALS = {
    "harmonic1b": (capi.POT_HARMONIC1B, lambda j: ([_get(j, "set")], [_get(j, "omega")])),
    "squareWell2b": (capi.POT_SQUAREWELL2B, lambda j: (_get(j, "sets"), [_get(j, "V0"), _get(j, "R0")])),
    "dipolar2b": (capi.POT_DIPOLAR2B, lambda j: (_get(j, "sets"), [_get(j, "D")])),
    "poschTeller2b": (capi.POT_POSCHTELLER2B, lambda j: (_get(j, "sets"), [_get(j, "R0")])),
    "potentialTwoBody/squareWell": (capi.POT_TWOBODY_SQUAREWELL, lambda j: (_get(j, "sets"), [_get(j["functor"], "R0"), _get(j["functor"], "V0")])),
    "potentialTwoBody/potentialBarrier": (capi.POT_TWOBODY_BARRIER, lambda j: (_get(j, "sets"), [_get(j["functor"], "R0"), _get(j["functor"], "V0")])),
    "potentialThreeBodyUnDis/squareWell": (capi.POT_THREEBODY_UNDIS_SQUAREWELL, lambda j: (_get(j, "sets"), [_get(j["functor"], "R0"), _get(j["functor"], "V0")])),
    "potentialThreeBodyUnDis/potentialBarrier": (capi.POT_THREEBODY_UNDIS_BARRIER, lambda j: (_get(j, "sets"), [_get(j["functor"], "R0"), _get(j["functor"], "V0")])),
    "potentialThreeBodyDis/squareWell": (capi.POT_THREEBODY_DIS_SQUAREWELL, lambda j: (_get(j, "sets"), [_get(j["functor"], "R0"), _get(j["functor"], "V0")])),
    "potentialThreeBodyDis/potentialBarrier": (capi.POT_THREEBODY_DIS_BARRIER, lambda j: (_get(j, "sets"), [_get(j["functor"], "R0"), _get(j["functor"], "V0")])),
}
# histogram observables over three sets / scalar observables (estimatorsFactory.h:60-95: scalar observables get a
# realScalarEstimator, which only reads "label"; here they are 1-bin estimators holding sum and count)
_THREE_SET_HISTOGRAMS = {"trimerCorrelation": capi.EST_TRIMERCORRELATION, "trimerCorrelationDis": capi.EST_TRIMERCORRELATION_DIS}
_SCALARS = {"trimerhyperRadius": (capi.EST_TRIMERHYPERRADIUS, lambda j: list(_get(j, "sets"))[:3]),
            "centerOfMassSquared": (capi.EST_CENTEROFMASSSQUARED, lambda j: [_get(j, "set")])}


def _dp(a):
    return a.ctypes.data_as(capi.c_double_p)


def _ip(a):
    return a.ctypes.data_as(capi.c_int_p)


class System:
    """One context = one GPU = one rank (the reference's one MPI rank)."""

    def __init__(self, lBox, N, device=0):
        self.lib = capi.load()
        self.N = [int(n) for n in N]
        self.dofs = 3 * sum(self.N)
        self.ctx = C.c_void_p()
        lb = np.asarray(lBox, dtype=np.float64)
        if lb.size != 3:
            raise invalidInput("Input file implies dimensionality different from 3")
        n = np.asarray(self.N, dtype=np.int32)
        rc = self.lib.qmc_create(C.byref(self.ctx), device, _dp(lb), len(self.N), _ip(n))
        if rc:
            raise RuntimeError(self.lib.qmc_last_error(None).decode())
        self.estimators = []   # (label, kind, bins, minx, maxx)
        self.hidden_estimators = []   # targets of forward-walking estimators (not reported, estimatorsFactory.cpp:47)
        self.rank, self.nranks = 0, 1

    # -- error plumbing ---------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            msg = self.lib.qmc_last_error(self.ctx).decode()
            if "Inalid initial" in msg or "set" in msg and "jastrow" in msg or "should be distinguishible" in msg:
                raise invalidInput(msg)
            raise RuntimeError(msg)

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.qmc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- factories (wavefunctionFactory.h:35-49, potentialFactory.h:43-55, estimatorsFactory.cpp:39-133)
    def add_wavefunction(self, jw):
        ident = create_id(jw)
        kind, _, jname = ident.partition("/")
        if kind not in _WAVES or jname not in _JASTROWS:
            raise factoryIdNotRecorded(ident)
        jk, extract = _JASTROWS[jname]
        params = np.asarray(extract(jw["jastrow"]), dtype=np.float64)
        sets = [_get(jw, "set")] if kind == "jastrow1b" else list(_get(jw, "sets"))
        if kind == "jastrow3bUnDis":
            sets = sets[:1]
        if kind == "jastrow3bDis":
            sets = sets[:3]
        sets = np.asarray(sets, dtype=np.int32)
        hs = float(jw["hardSphereRadius"]) if "hardSphereRadius" in jw else -1.0
        self._ck(self.lib.qmc_add_wavefunction_term(self.ctx, _WAVES[kind], _ip(sets), len(sets), jk, _dp(params), len(params), hs))

    def add_potential(self, jp):
        ident = create_id(jp)
        if ident not in _POTENTIALS:
            raise factoryIdNotRecorded(ident)
        pk, extract = _POTENTIALS[ident]
        sets, params = extract(jp)
        sets = np.asarray(sets, dtype=np.int32)
        params = np.asarray(params, dtype=np.float64)
        self._ck(self.lib.qmc_add_potential_term(self.ctx, pk, _ip(sets), len(sets), _dp(params), len(params)))

    def add_measurement(self, jm):
        kind = _get(jm, "kind")
        if kind in ("energy", "forceEnergy"):       # skipped by the factory (estimatorsFactory.cpp:122)
            return None
        idx = C.c_int(-1)
        if kind == "pairCorrelation":
            sets = np.asarray(_get(jm, "sets"), dtype=np.int32)
            self._ck(self.lib.qmc_add_estimator(self.ctx, capi.EST_PAIRCORRELATION, _ip(sets), 2, int(jm["bins"]),
                                                float(jm["minx"]), float(jm["maxx"]), None, 0, C.byref(idx)))
        elif kind == "trimerCorrelationUnDis":
            sets = np.asarray(_get(jm, "sets")[:1], dtype=np.int32)
            self._ck(self.lib.qmc_add_estimator(self.ctx, capi.EST_TRIMERCORRELATION_UNDIS, _ip(sets), 1, int(jm["bins"]),
                                                float(jm["minx"]), float(jm["maxx"]), None, 0, C.byref(idx)))
        elif kind in _THREE_SET_HISTOGRAMS:
            sets = np.asarray(list(_get(jm, "sets"))[:3], dtype=np.int32)
            self._ck(self.lib.qmc_add_estimator(self.ctx, _THREE_SET_HISTOGRAMS[kind], _ip(sets), len(sets), int(jm["bins"]),
                                                float(jm["minx"]), float(jm["maxx"]), None, 0, C.byref(idx)))
        elif kind in _SCALARS:
            ek, getsets = _SCALARS[kind]
            sets = np.asarray(getsets(jm), dtype=np.int32)
            self._ck(self.lib.qmc_add_estimator(self.ctx, ek, _ip(sets), len(sets), 1, 0.0, 1.0, None, 0, C.byref(idx)))
        elif kind == "superfluidFraction":
            # estimatorsFactory.cpp:80-116: centre-of-mass storer shared by all such measurements + one winding estimator each
            sets = list(_get(jm, "sets"))
            if len(sets) != 2:
                raise invalidInput("Superfluid fraction should have two sets")
            self._ck(self.lib.qmc_add_superfluid_fraction(self.ctx, int(sets[0]), int(sets[1]), int(_get(jm, "bins")), C.byref(idx)))
            self.estimators.append(dict(label=_get(jm, "label"), kind=kind, bins=int(jm["bins"]), index=idx.value, scalar=False,
                                        vector=True, minx=0.0, maxx=0.0))
            return idx.value
        elif kind == "structureFactor":
            sets = np.asarray(_get(jm, "sets"), dtype=np.int32)
            extra = np.asarray([jm["maxq"], jm["lBox"], jm["channel"]], dtype=np.float64)
            self._ck(self.lib.qmc_add_estimator(self.ctx, capi.EST_STRUCTUREFACTOR, _ip(sets), 2, int(jm["bins"]),
                                                0.0, 0.0, _dp(extra), 3, C.byref(idx)))
        else:
            raise factoryIdNotRecorded(kind)
        plain = dict(label=jm.get("label", kind), kind=kind, bins=int(jm.get("bins", 1)), index=idx.value, scalar=kind in _SCALARS,
                     minx=float(jm.get("minx", 0.0)), maxx=float(jm.get("maxx", 0.0)))
        if "forwardWalkingSteps" in jm:
            # estimatorsFactory.cpp:47-74: one forward-walking estimator "<label>_fw<steps>" per entry INSTEAD of
            # the plain one; the storer exists when recordSteps > 0 (estimatorsFactory.h:112-121)
            fw = np.asarray([int(x) for x in jm["forwardWalkingSteps"]], dtype=np.int32)
            first = C.c_int(-1)
            self._ck(self.lib.qmc_add_forward_walking(self.ctx, idx.value, int(_get(jm, "recordSteps")), _ip(fw), len(fw),
                                                      C.byref(first)))
            plain["hidden"] = True                 # still accumulated on the device, not reported by the reference
            self.hidden_estimators.append(plain)
            for k, steps in enumerate(fw):
                self.estimators.append(dict(plain, label="%s_fw%d" % (plain["label"], steps), index=first.value + k,
                                            forwardWalkingSteps=int(steps), targetLabel=plain["label"], hidden=False))
            return idx.value
        self.estimators.append(plain)
        return idx.value

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.qmc_comm_init(self.ctx, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        if self.lib.qmc_comm_unique_id(buf):
            raise RuntimeError(self.lib.qmc_last_error(None).decode())
        return buf.raw

    def set_dmc(self, timeStep, walkers, delta_walkers=None, seed=0, branching=True, capacity=0):
        if delta_walkers is None:
            delta_walkers = int(walkers * 0.1)          # main.cpp:364
        self._ck(self.lib.qmc_set_dmc(self.ctx, float(timeStep), int(walkers), int(delta_walkers), int(seed),
                                      1 if branching else 0, int(capacity)))

    def set_vmc(self, sigma, chains=1, seed=0, capacity=0):
        self._ck(self.lib.qmc_set_vmc(self.ctx, float(sigma), int(chains), int(seed), int(capacity)))

    @classmethod
    def from_input(cls, j, device=0, comm=None, capacity=0):
        """j: the reference's input document (dict).  comm: None or (unique_id, rank, nranks)."""
        if isinstance(j, str):
            j = json.loads(j)
        lBox = _get(j, "lBox")
        if len(lBox) != 3:
            raise invalidInput("Input file implies dimensionality different from 3")
        s = cls(lBox, _get(j, "N"), device=device)
        for jw in _get(j, "wavefunctions"):
            s.add_wavefunction(jw)
        for jp in j.get("potentials", []):
            s.add_potential(jp)
        for jm in j.get("measurements", []):
            s.add_measurement(jm)
        if comm is not None:
            s.comm_init(*comm)
        method = j.get("method", "dmc")
        walkers = int(j.get("walkers", 1))
        # main.cpp:318-335,395-408: the forceEnergy estimator is attached for vmc and svmc only (ignored for dmc, as there)
        s.force_energy_enabled = method in ("vmc", "svmc") and any(m.get("kind") == "forceEnergy" for m in j.get("measurements", []))
        if s.force_energy_enabled:
            s._ck(s.lib.qmc_enable_force_energy(s.ctx))
        if method == "vmc":          # vmcDriver(&psi, timeStep) (main.cpp:314): "walkers" independent chains on the GPU
            s.set_vmc(j.get("timeStep", 1e-3), walkers, seed=int(j.get("seed", 0)), capacity=capacity)
            s.input = j
            return s
        s.set_dmc(j.get("timeStep", 1e-3), walkers, j.get("delta_walkers"), seed=int(j.get("seed", 0)),
                  branching=(method != "svmc"), capacity=capacity)
        s.input = j
        return s

    # -- data path -----------------------------------------------------------------------------
    def eval(self, R):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, self.dofs)
        nW = R.shape[0]
        out = dict(logpsi=np.zeros(nW), grad=np.zeros((nW, self.dofs)), lap=np.zeros(nW), V=np.zeros(nW),
                   E=np.zeros(nW), valid=np.zeros(nW, dtype=np.int32))
        self._ck(self.lib.qmc_eval(self.ctx, _dp(R), nW, _dp(out["logpsi"]), _dp(out["grad"]), _dp(out["lap"]),
                                   _dp(out["V"]), _dp(out["E"]), _ip(out["valid"])))
        return out

    def tabulate_wavefunction(self, index, minx, maxx, n):
        """jastrow<T>::print (jastrow.h:42-63) evaluated on the device: n rows {x, u, u', u''}"""
        out = np.zeros((int(n), 4))
        self._ck(self.lib.qmc_tabulate_wavefunction(self.ctx, int(index), float(minx), float(maxx), int(n), _dp(out)))
        return out

    def accumulate_configs(self, R):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, self.dofs)
        self._ck(self.lib.qmc_accumulate_configs(self.ctx, _dp(R), R.shape[0]))

    def upload(self, R):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, self.dofs)
        self._ck(self.lib.qmc_upload_walkers(self.ctx, _dp(R), R.shape[0]))

    def download(self, with_energy=False):
        n = C.c_int64(0)
        self._ck(self.lib.qmc_download_walkers(self.ctx, None, 0, C.byref(n), None))
        R = np.zeros((n.value, self.dofs))
        E = np.zeros(n.value)
        if n.value:
            self._ck(self.lib.qmc_download_walkers(self.ctx, _dp(R), n.value, C.byref(n), _dp(E)))
        return (R, E) if with_energy else R

    def step(self, nSteps=1, sampleEvery=1):
        self._ck(self.lib.qmc_step(self.ctx, int(nSteps), int(sampleEvery)))

    def step_traced(self, accumulate=False):
        n0, n1 = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.qmc_download_walkers(self.ctx, None, 0, C.byref(n0), None))
        n = n0.value
        d = self.dofs
        arrs = dict(xi=np.zeros((n, d)), Rprop=np.zeros((n, d)), Fprop=np.zeros((n, d)), logpsi=np.zeros(n), lap=np.zeros(n),
                    V=np.zeros(n), Eprop=np.zeros(n), q=np.zeros(n), u_metropolis=np.zeros(n), u_branch=np.zeros(n),
                    accepted=np.zeros(n, np.int32), valid=np.zeros(n, np.int32), n_desc=np.zeros(n, np.int32),
                    parent=np.zeros(n, np.int32))
        t = capi.Trace()
        for k, a in arrs.items():
            setattr(t, k, _ip(a) if a.dtype == np.int32 else _dp(a))
        self._ck(self.lib.qmc_step_traced(self.ctx, 1 if accumulate else 0, C.byref(t), n, C.byref(n0), C.byref(n1)))
        arrs["n_before"], arrs["n_after"] = n0.value, n1.value
        return arrs

    def block_results(self):
        r = capi.BlockResult()
        self._ck(self.lib.qmc_block_results(self.ctx, C.byref(r)))
        return {f: getattr(r, f) for f, _ in capi.BlockResult._fields_}

    def estimator(self, index):
        e = [x for x in self.estimators + self.hidden_estimators if x["index"] == index][0]
        sums = np.zeros(e["bins"])
        if e.get("vector"):           # realVectorAccumulator: one weight per bin (superfluidFraction)
            ws = np.zeros(e["bins"])
            self._ck(self.lib.qmc_get_vector_estimator(self.ctx, index, _dp(sums), _dp(ws), e["bins"]))
            return sums, ws
        w = C.c_double(0)
        self._ck(self.lib.qmc_get_estimator(self.ctx, index, _dp(sums), e["bins"], C.byref(w)))
        return sums, w.value

    def force_energy(self):
        """(sum, count) of forceEnergy::operator() over the walkers of the sampling steps since the last clear"""
        a, n = C.c_double(0), C.c_double(0)
        self._ck(self.lib.qmc_get_force_energy(self.ctx, C.byref(a), C.byref(n)))
        return a.value, n.value

    def clear_estimators(self):
        self._ck(self.lib.qmc_clear_estimators(self.ctx))

    def rebalance(self):
        m = C.c_int64(0)
        self._ck(self.lib.qmc_rebalance(self.ctx, C.byref(m)))
        return m.value

    def set_migration(self, every=1, band=0.02):
        """in-step walker migration of qmc_step: plan every `every` steps, dead band `band` x mean population"""
        self._ck(self.lib.qmc_set_migration(self.ctx, int(every), float(band)))

    def synchronize(self):
        self._ck(self.lib.qmc_synchronize(self.ctx))

    def info(self, key):
        v = C.c_double(0)
        self._ck(self.lib.qmc_get_info(self.ctx, key.encode(), C.byref(v)))
        return v.value


def plan_rebalance(populations):
    """The alias-method planner (ptools.cpp:93-161) through the C ABI (host-only, no GPU needed)."""
    lib = capi.load()
    p = np.asarray(populations, dtype=np.int32)
    src = np.zeros(len(p), dtype=np.int32)
    amt = np.zeros(len(p), dtype=np.int32)
    if lib.qmc_plan_rebalance(_ip(p), len(p), _ip(src), _ip(amt)):
        raise RuntimeError("qmc_plan_rebalance failed")
    return src, amt
