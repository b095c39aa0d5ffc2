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
# the CPU suite has to stay within minutes: the emulation is ~10^4 x slower than the B200, so the slowest cases of the
# never-on-the-critical-path families run on hardware only (and here with QMC4_TEST_EMULATED_SLOW=1)
SLOW = os.environ.get("QMC4_TEST_EMULATED_SLOW") == "1"
LIGHT = FAST if SLOW else (FAST + " and not host_driver_three_set and not ring_restarts and not measurement_file and not force_energy_file"
                           " and not bitwise_reproducible and not impurity_hyper and not (forward_walking_vs_oracle and 3)"
                           " and not (agrees_with_the_general and (256 or 512 or 160 or 100)) and not baseline_sizes and not is_bitwise"
                           " and not branching_bookkeeping and not (upload_keeps_every_walker and 40000)")


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


@pytest.mark.parametrize("suite", ["test_gpu_parity.py", "test_pipeline2.py", "test_zz_gpu_new_paths.py"])
def test_parity_suite_on_emulated_kernels(emu_lib, suite):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", suite), "-q", "-x",
                        "-m", "gpu", "-k", LIGHT, "-p", "no:cacheprovider"],
                       env=dict(_env(emu_lib), QMC4_DMC_BIN=os.path.join(EMU_DIR, "dmc_b200_emu"), QMC4_EMU_DEVICES="1",
                                QMC4_NCCL_LIB=os.path.join(EMU_DIR, "libnccl_emu.so")),
                       cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail + r.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail


@pytest.mark.parametrize("nranks,N,wpr,wpw", [(2, 16, 24, ""), (8, 16, 6, "")] + ([(4, 40, 10, "2"), (8, 64, 6, "8"), (2, 2, 1300, "")] if SLOW else []))
def test_multirank_steps_and_rebalance_on_emulated_ranks(emu_lib, nranks, N, wpr, wpw):
    """ranks = threads over an in-process NCCL stand-in: global E_T / block results identical on every
    rank, rebalance conserves the walker multiset and equalises the populations (ptools.cpp:93-161).  The slow list adds
    1300 walkers per rank: in-step migration (n_adjust) through the multi-CTA scan"""
    env = _env(emu_lib)
    env["QMC4_NCCL_LIB"] = os.path.join(EMU_DIR, "libnccl_emu.so")
    if wpw:
        env["QMC4_WPW"] = wpw
    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "multirank_driver.py"), str(nranks), str(N), str(wpr)],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nranks,sample_every,mid", [(2, 1, 0)] + ([(4, 2, 0), (3, 3, 1)] if SLOW else []))
def test_forward_walking_records_migrate_between_emulated_ranks(emu_lib, nranks, sample_every, mid):
    """forward walking + superfluidFraction on several ranks: qmc_rebalance ships the lineage records of the tail walkers
    (chain of ancestor slots -> message -> private lineage on the receiver); the rank-reduced estimators must equal the
    oracle's per-walker rings over the GLOBAL genealogy rebuilt from the ranks' traces.  mid = 1: rebalances also happen
    between two sampling events, where the composed ancestor map (not the fresh parent map) is what travels"""
    env = _env(emu_lib)
    env["QMC4_NCCL_LIB"] = os.path.join(EMU_DIR, "libnccl_emu.so")
    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "multirank_fw_driver.py"), str(nranks), str(sample_every), str(mid)],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(os.environ.get("QMC4_TEST_EMULATED_SLOW") != "1", reason="minutes of CPU: set QMC4_TEST_EMULATED_SLOW=1")
def test_pipeline2_bitwise_on_emulated_kernels(emu_lib):
    env = _env(emu_lib)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_pipeline2.py"), "-q",
                        "-m", "gpu", "-p", "no:cacheprovider"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def _tsan_inputs():
    """small host-driver inputs that together reach every kernel family"""
    import copy
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    small = dict(stepsPerBlock=3, nBlocks=2, correlationSteps=0)
    fw = cases.forward_walking_cases()
    pair = copy.deepcopy(fw["pbc_pair"])
    pair["measurements"].append({"kind": "superfluidFraction", "label": "rho", "sets": [0, 0], "bins": 4})
    vmc = copy.deepcopy(cases.vmc_case())
    vmc["measurements"].append({"kind": "forceEnergy", "label": "forceEnergy"})
    out = {
        "fw_pair_superfluid": dict(pair, walkers=10, delta_walkers=2, stepsPerBlock=3, nBlocks=2, correlationSteps=1),
        "fw_trimer": dict(fw["three_body_trimer"], walkers=8, delta_walkers=2, **small),
        "fw_mixture": dict(fw["mixture_g01"], walkers=8, delta_walkers=2, stepsPerBlock=4, nBlocks=2, correlationSteps=0),
        "fw_impurity_scalar": dict(fw["impurity_hyper"], walkers=8, delta_walkers=2, **small),
        "triplets_dis": dict(cases.mixture3(), walkers=6, delta_walkers=2, timeStep=1e-3, **small),
        "triplets_dis_hard_sphere": dict(cases.mixture3(hs=1.3), walkers=6, delta_walkers=2, timeStep=1e-3, **small),
        "vmc_force_energy": dict(vmc, walkers=6, stepsPerBlock=3, nBlocks=2, correlationSteps=1),
        "generic_dipolar_rep": dict(cases.pbc(10, 5.2, walkers=6, jastrow=cases.DIPOLAR_REP), delta_walkers=2, timeStep=1e-3, **small),
        "structure_factor": dict(cases.structure_factor(), walkers=6, delta_walkers=2, **small),
        # round-2 kernels: tournament pair kernel with a ragged last block + its fixed-point g(r) kernel, the packed small-system
        # kernel, and a population that needs several scan CTAs (the finishing CTA of the scan once cleared the published totals
        # while its thread 0 was still adding them: exactly the access pair tsan reports)
        "pair_kernel_g_of_r": dict(cases.pbc(70, 4.1, walkers=6, meas=[{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 50,
                                                                         "minx": 0.1, "maxx": 2.05}]), delta_walkers=2, **small),
        "small_trap": dict(cases.trap(5, walkers=12), delta_walkers=3, initialConditionGenerator={"lBox": [2.0, 2.0, 2.0]}, **small),
        "scan_many_ctas": dict(cases.trap(2, walkers=1100), delta_walkers=100, initialConditionGenerator={"lBox": [2.0, 2.0, 2.0]},
                               stepsPerBlock=2, nBlocks=1, correlationSteps=0),
    }
    return out


@pytest.mark.skipif(os.environ.get("QMC4_TEST_EMULATED_SANITIZERS") != "1", reason="~15 minutes of CPU: set QMC4_TEST_EMULATED_SANITIZERS=1")
@pytest.mark.parametrize("san", ["tsan", "asan"])
def test_kernels_under_sanitizers(tmp_path, san):
    """ThreadSanitizer / AddressSanitizer over the emulation (tests/emu/Makefile `tsan`, `asan`).  tsan: CUDA threads are OS
    threads and every barrier, shuffle and atomic is a std::atomic operation, so an access pair the GPU does not order shows up
    as a data race.  asan: every device allocation and every launch's shared memory is its own heap block, so an out-of-range
    index is an error.  Run for all kernel families through the host driver, on 1 emulated GPU, on 2 (NCCL stand-in, record
    migration) and with pipeline 2."""
    import json
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "-j4", san])
    d_san = os.path.join(EMU_DIR, san)

    def run(name, cfg, *extra, **env):
        d = tmp_path / name
        d.mkdir()
        json.dump(cfg, open(d / "input.json", "w"))
        e = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0", ASAN_OPTIONS="detect_leaks=0", QMC4_EMU_DEVICES="1",
                 QMC4_NCCL_LIB=os.path.join(d_san, "libnccl_emu.so"))
        e.update(env)
        p = subprocess.run([os.path.join(d_san, "dmc_b200_emu"), "input.json", *extra], cwd=d, env=e, capture_output=True, text=True, timeout=1800)
        assert p.returncode == 0, name + p.stderr[-1500:]
        assert "Sanitizer" not in p.stderr, name + "\n" + p.stderr[:3000]

    inputs = _tsan_inputs()
    for name, cfg in inputs.items():
        run(name, cfg)
    run("two_gpus_record_migration", inputs["fw_pair_superfluid"], "--gpus", "2", QMC4_EMU_DEVICES="2")
    run("pipeline1", inputs["structure_factor"], QMC4_PIPELINE="1")


def _run_host(tmp_path, cfg, *extra, devices=1):
    import json
    with open(os.path.join(tmp_path, "input.json"), "w") as f:
        json.dump(cfg, f)
    env = dict(os.environ, QMC4_EMU_DEVICES=str(devices), QMC4_NCCL_LIB=os.path.join(EMU_DIR, "libnccl_emu.so"))
    return subprocess.run([os.path.join(EMU_DIR, "dmc_b200_emu"), "input.json", *extra], cwd=tmp_path, env=env,
                          capture_output=True, text=True, timeout=900)


def test_host_driver_files_on_emulated_kernels(emu_lib, tmp_path):
    """dmc_b200 (host C++ over the C ABI) end to end: block report, energy.dat, <label>.dat, checkpoint, restart"""
    import json
    import re
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    cfg = cases.dmc_cases()["pbc_N16"]
    cfg = dict(cfg, nBlocks=3, walkers=24, delta_walkers=3, stepsPerBlock=6)
    p = _run_host(str(tmp_path), cfg)
    assert p.returncode == 0, p.stderr
    out = re.sub(r"\x1b\[[0-9;]*m", "", p.stdout)
    assert out.count("Block: ") == 3 and "Acc. Ratio: " in out and "totalRunningTime :" in out
    lines = open(tmp_path / "energy.dat").read().split("\n")
    assert lines[0] == "energy" and len([x for x in lines[1:] if x.strip()]) == 3
    assert re.fullmatch(r"-?\d\.\d{12}e[+-]\d\d", lines[1])
    m = cfg["measurements"][0]
    blocks = open(tmp_path / (m["label"] + ".dat")).read().strip().split("\n\n")
    rows = blocks[0].split("\n")
    assert rows[0] == "x " + m["label"] and len(blocks) == 3 and len(rows) == m["bins"] + 1
    g = np.array([[float(r.split()[1]) for r in b.split("\n") if not r.startswith("x")] for b in blocks])
    assert np.isfinite(g).all() and g.sum() > 0
    ck = json.load(open(tmp_path / "configurations" / "walkers-Rank0.json"))
    assert sorted(ck["configurations"][0][0]) == ["x", "y", "z"]
    p2 = _run_host(str(tmp_path), dict(cfg, configurations="configurations", nBlocks=1))
    assert p2.returncode == 0 and "WARNING" not in p2.stdout
    pops = [int(x) for x in re.findall(r"Curr. Walkers: (\d+)", re.sub(r"\x1b\[[0-9;]*m", "", p2.stdout))]
    assert abs(pops[0] - len(ck["configurations"])) <= 0.5 * len(ck["configurations"])


def test_host_driver_two_gpu_threads_on_emulated_kernels(emu_lib, tmp_path):
    """--gpus 2: one host thread per (emulated) GPU, NCCL unique id shared in-process, per-rank checkpoints"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    cfg = dict(cases.dmc_cases()["trap_N10"], nBlocks=3, walkers=40, delta_walkers=4, stepsPerBlock=10)
    p = _run_host(str(tmp_path), cfg, "--gpus", "2", devices=2)
    assert p.returncode == 0, p.stderr
    e = np.loadtxt(tmp_path / "energy.dat", skiprows=1)
    assert len(e) == 3 and np.isfinite(e).all() and abs(e.mean() - 15.15) < 5
    assert os.path.exists(tmp_path / "configurations" / "walkers-Rank1.json")
