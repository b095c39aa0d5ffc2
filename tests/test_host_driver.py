"""GPU tests (-m gpu) of the C++ host driver qmc4_b200/host/dmc_b200: the reference's CLI, input schema,
block report and output files, running on the GPU through the C ABI."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# QMC4_DMC_BIN: tests/test_emulated_kernels.py points this at tests/emu/dmc_b200_emu to run the same checks without a GPU
BIN = os.environ.get("QMC4_DMC_BIN", os.path.join(ROOT, "qmc4_b200", "host", "dmc_b200"))
GOLD = os.path.join(ROOT, "tests", "golden")


def run(cfg, cwd, *extra):
    with open(os.path.join(cwd, "input.json"), "w") as f:
        json.dump(cfg, f)
    p = subprocess.run([BIN, "input.json", *extra], cwd=cwd, capture_output=True, text=True, timeout=600)
    return p


def _blocked_error(x):
    x = np.asarray(x, float)
    errs = []
    while len(x) >= 8:
        errs.append(x.std(ddof=1) / np.sqrt(len(x)))
        x = 0.5 * (x[: len(x) // 2 * 2: 2] + x[1: len(x) // 2 * 2: 2])
    return max(errs)


@pytest.mark.gpu
def test_unchanged_input_json_runs_and_matches_reference_statistics(tmp_path):
    fx = json.load(open(os.path.join(GOLD, "dmc_reference_runs.json")))["pbc_N16"]
    cfg = fx["input"]
    p = run(cfg, str(tmp_path))
    assert p.returncode == 0, p.stderr
    out = re.sub(r"\x1b\[[0-9;]*m", "", p.stdout)
    assert out.count("Block: ") == cfg["nBlocks"] and "Acc. Ratio: " in out and "Curr. Walkers: " in out
    assert "totalRunningTime :" in out and "outputTime :" in out
    # energy.dat: header + one %.12e record per block (estimators.cpp:35-46,89-98)
    lines = open(tmp_path / "energy.dat").read().split("\n")
    assert lines[0] == "energy"
    e = np.array([float(x) for x in lines[1:] if x.strip()])
    assert len(e) == cfg["nBlocks"] and re.fullmatch(r"-?\d\.\d{12}e[+-]\d\d", lines[1])
    e_ref = np.asarray(fx["energy_per_block"])[15:]
    sig = np.hypot(_blocked_error(e[15:]), _blocked_error(e_ref))
    assert abs(e[15:].mean() - e_ref.mean()) < 2 * sig
    # pair.dat: "x pair" header, per block `bins` lines "i*dx value" + blank line (estimators.cpp:80-87,29-33)
    blocks = open(tmp_path / "pair.dat").read().strip().split("\n\n")
    rows = blocks[0].split("\n")
    assert rows[0] == "x pair" and len(blocks) == cfg["nBlocks"]
    m = cfg["measurements"][0]
    first = np.array([[float(v) for v in r.split()] for r in rows[1:]])
    assert first.shape == (m["bins"], 2)
    np.testing.assert_allclose(first[:, 0], np.arange(m["bins"]) * (m["maxx"] - m["minx"]) / m["bins"], rtol=1e-5)
    g = np.array([[float(r.split()[1]) for r in b.split("\n") if not r.startswith("x")] for b in blocks])[15:]
    g_ref = np.asarray(fx["pair_per_block"])[15:]
    sg = np.hypot([_blocked_error(g[:, i]) for i in range(g.shape[1])], [_blocked_error(g_ref[:, i]) for i in range(g.shape[1])])
    z = np.abs(g.mean(0) - g_ref.mean(0)) / np.maximum(sg, 1e-12)
    assert (z < 2).mean() > 0.8 and z.max() < 4.5, z
    # checkpoint in the reference's format (tools.cpp:27-53) and restart from it
    ck = json.load(open(tmp_path / "configurations" / "walkers-Rank0.json"))
    w0 = ck["configurations"][0][0]
    assert sorted(w0) == ["x", "y", "z"] and len(w0["x"]) == cfg["N"][0]
    cfg2 = dict(cfg, configurations="configurations", nBlocks=2)
    p2 = run(cfg2, str(tmp_path))
    assert p2.returncode == 0 and "WARNING" not in p2.stdout
    pops = [int(x) for x in re.findall(r"Curr. Walkers: (\d+)", re.sub(r"\x1b\[[0-9;]*m", "", p2.stdout))]
    assert abs(pops[0] - len(ck["configurations"])) < 0.2 * len(ck["configurations"])


def test_errors_follow_the_reference(tmp_path):
    """factory and input errors are raised before any device work: checked on CPU"""
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "qmc4_b200", "host")])
    cfg = cases.pbc(8, 4.0, walkers=8)
    cfg["wavefunctions"][0]["jastrow"]["kind"] = "nonexistent"
    p = run(cfg, str(tmp_path))
    assert p.returncode != 0 and "jastrow2bUnDis/nonexistent" in p.stderr        # factoryIdNotRecorded(id)
    cfg = cases.pbc(8, 4.0, walkers=8)
    del cfg["correlationSteps"]
    p = run(cfg, str(tmp_path))
    assert p.returncode != 0 and "correlationSteps" in p.stderr                  # required key (main.cpp:287)
    cfg = cases.pbc(8, 4.0, walkers=8, method="pigs")
    p = run(cfg, str(tmp_path))
    assert p.returncode != 0 and "method 'pigs'" in p.stderr                     # only dmc / svmc / vmc are device paths
    cfg = cases.pbc(8, 4.0, walkers=8)
    cfg["lBox"] = [4.0, 4.0]
    p = run(cfg, str(tmp_path))
    assert p.returncode != 0 and "dimensionality" in p.stderr                    # main.cpp:173-177
    cfg = cases.pbc(8, 4.0, walkers=8)
    cfg["wavefunctions"][0]["sets"] = [0, 1]
    cfg["N"] = [4, 4]
    p = run(cfg, str(tmp_path))
    assert p.returncode != 0 and "setA != setB" in p.stderr                      # jastrowWavefunctionTwoBody.h:100
    json.dump({"kind": "gaussiangaussian", "r0": 1.0, "u0": -0.5}, open(tmp_path / "jastrow.json", "w"))
    cfg = cases.pbc(8, 4.0, walkers=8)
    cfg["wavefunctions"][0]["jastrow"] = "@jastrow.json"                          # "@file" links (input.cpp:7-34)
    cfg["potentials"][0]["functor"]["kind"] = "nonexistent"
    p = run(cfg, str(tmp_path))
    assert p.returncode != 0 and "potentialTwoBody/nonexistent" in p.stderr      # link resolved, ids built in key order


@pytest.mark.gpu
def test_multi_gpu_threads(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg = cases.dmc_cases()["trap_N10"]
    cfg = dict(cfg, nBlocks=20, walkers=800)
    p = run(cfg, str(tmp_path), "--gpus", "2")
    assert p.returncode == 0, p.stderr
    e = np.loadtxt(tmp_path / "energy.dat", skiprows=1)
    assert len(e) == 20 and abs(e[8:].mean() - 15.15) < 0.5
    assert os.path.exists(tmp_path / "configurations" / "walkers-Rank1.json")
