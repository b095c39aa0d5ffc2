"""CPU-suite coverage of the CUDA kernels' logic through tests/emu (TEST INFRASTRUCTURE).

tests/emu compiles the repository's CUDA sources with g++ against a stand-in <cuda_runtime.h> and
runs every kernel with one OS thread per CUDA thread (real barriers, shuffles, shared memory,
atomics).  This file re-runs the fast part of the `-m gpu` parity suite against that build in a
subprocess, so a container without a GPU still checks walker groups, pair tiles, parent maps, the
prefix-sum scan, Metropolis/branching and the fixed-point histograms against the oracle and the
reference's golden vectors.  It proves logic, not performance, and it is not a fallback: the product
loader refuses the emulation build unless QMC4_B200_ALLOW_EMULATION=tests-only is set (checked below).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libqmc4b200_emu.so")

FAST = "not statistical and not population_control and not determinism and not normal_deviates"


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "all"])
    return EMU_LIB


def _env(lib, allow=True):
    env = dict(os.environ)
    env["QMC4_B200_LIB"] = lib
    env.pop("QMC4_B200_ALLOW_EMULATION", None)
    if allow:
        env["QMC4_B200_ALLOW_EMULATION"] = "tests-only"
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    return env


def test_product_loader_refuses_the_emulation_build(emu_lib):
    """pointing the package at the emulation build is an ImportError, not a silent CPU path"""
    r = subprocess.run([sys.executable, "-c", "import qmc4_b200.capi as c; c.load()"], env=_env(emu_lib, allow=False),
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0
    assert "emulation build" in r.stderr and "no CPU fallback" in r.stderr


def test_shipped_library_has_no_emulation_marker():
    import ctypes
    lib = os.path.join(ROOT, "qmc4_b200", "libqmc4b200.so")
    if not os.path.exists(lib):
        pytest.skip("CUDA library not built")
    assert not hasattr(ctypes.CDLL(lib), "qmc_emulated")


def test_parity_suite_on_emulated_kernels(emu_lib):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                        "-m", "gpu", "-k", FAST, "-p", "no:cacheprovider"],
                       env=_env(emu_lib), cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail + r.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail


@pytest.mark.parametrize("nranks,N,wpr,wpw", [(2, 16, 24, ""), (4, 40, 10, "2"), (8, 64, 6, "8")])
def test_multirank_steps_and_rebalance_on_emulated_ranks(emu_lib, nranks, N, wpr, wpw):
    """ranks = threads over an in-process NCCL stand-in: global E_T / block results identical on every
    rank, rebalance conserves the walker multiset and equalises the populations (ptools.cpp:93-161)"""
    env = _env(emu_lib)
    env["QMC4_NCCL_LIB"] = os.path.join(EMU_DIR, "libnccl_emu.so")
    if wpw:
        env["QMC4_WPW"] = wpw
    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "multirank_driver.py"), str(nranks), str(N), str(wpr)],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(os.environ.get("QMC4_TEST_EMULATED_SLOW") != "1", reason="minutes of CPU: set QMC4_TEST_EMULATED_SLOW=1")
def test_experimental_pipeline2_bitwise_on_emulated_kernels(emu_lib):
    env = _env(emu_lib)
    env["QMC4_TEST_EXPERIMENTAL"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_experimental_pipeline2.py"), "-q",
                        "-m", "gpu", "-p", "no:cacheprovider"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
