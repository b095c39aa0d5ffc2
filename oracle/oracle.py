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


class Port(_Base):
    def __init__(self):
        if not have_port():
            raise FileNotFoundError(PORT_LIB + " — run `make -C oracle port`")
        super().__init__(PORT_LIB, "orc")
