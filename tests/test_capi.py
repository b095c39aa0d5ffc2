"""CPU tests (-m "not gpu") of the drop-in boundary: the C-ABI library loads, exports every symbol
include/qmc4_b200.h declares, refuses to run without a GPU (no CPU fallback), and its host-side
pieces (planner, id scheme) match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from qmc4_b200 import capi
from qmc4_b200.system import System, create_id, factoryIdNotRecorded, plan_rebalance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qmc4_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(qmc_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 24
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(capi.EXPORTS) == declared
    assert lib.qmc_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="CUDA device"):
        System([4.0, 4.0, 4.0], [8])


def test_oracle_is_not_reachable_from_the_product():
    """the product package must not import, link or execute anything under oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qmc4_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("oracle tolerance", ""), os.path.join(dirpath, f)


def test_planner_matches_oracle(port):
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 9))
        p = rng.integers(0, 200, size=n)
        src, amt = plan_rebalance(p)
        s2, a2 = port.plan(p)
        np.testing.assert_array_equal(src, s2)
        np.testing.assert_array_equal(amt, a2)
        new = p.copy()
        for r in range(n):
            new[r] += amt[r]
            new[src[r]] -= amt[r]
        assert new.sum() == p.sum() and np.all(np.abs(new - p.sum() // n) <= 1)


def test_create_id_scheme():
    """createId (dmc/abstractFactory.cpp:3-20): kind + '/' + ids of structured children, key-sorted."""
    assert create_id({"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "gaussiangaussian", "r0": 1, "u0": 1}}) == "jastrow2bUnDis/gaussiangaussian"
    assert create_id({"kind": "potentialTwoBody", "sets": [0, 0], "functor": {"kind": "squareWell", "R0": 1, "V0": 1}}) == "potentialTwoBody/squareWell"
    assert create_id({"kind": "harmonic1b", "omega": 1.0, "set": 0}) == "harmonic1b"
    assert create_id({"kind": "jastrow1b", "set": 0, "jastrow": {"kind": "gaussian", "alpha": 0.5}}) == "jastrow1b/gaussian"


def test_unknown_id_raises_like_the_reference_factory():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a context")
    s = System([4.0] * 3, [4])
    with pytest.raises(factoryIdNotRecorded):
        s.add_wavefunction({"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "nonexistent"}})
