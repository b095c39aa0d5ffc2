"""CPU test (-m "not gpu") of the bench.py contract that can be checked without a GPU: the reference arm
(`--impl reference`) runs the reference's CPU implementation (oracle/_ref, or the C restatement when the
binary is absent) on a bounded sample and prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "dmc_walker_moves_per_sec" and line["unit"] == "walker-moves/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "walker-moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("config5")          # the configuration the north-star targets are quoted on
    assert line["config"]["cpu_walkers_per_process"] >= 8 and line["config"]["dmc_steps_per_bench_step"] >= 1


def test_workloads_follow_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    for cfgno, N in ((1, 10), (2, 64), (3, 256), (4, 32), (5, 512)):
        cfg = bench.workload(cfgno, bench.DEFAULT_WALKERS[cfgno], 1)
        assert cfg["N"] == [N] and cfg["correlationSteps"] == 0 and cfg["timeStep"] == 1e-3
        pairs, trip, flops, nbytes = bench.work_per_move(cfg)
        assert pairs == N * (N - 1) / 2 and nbytes == 96 * N + 48
        assert flops >= 64 * pairs
    assert bench.workload(2, 16384, 8)["walkers"] == 8 * 16384          # weak scaling: global target grows with the ranks


def test_committed_workload_files_are_what_bench_feeds_both_arms():
    """bench/configs/config<i>.json == bench.workload(i): the documents are fed unchanged to both arms"""
    import json
    sys.path.insert(0, ROOT)
    import bench as B
    for c in (1, 2, 3, 4, 5):
        ng = 8 if c == 5 else 1
        want = json.loads(json.dumps(B.workload(c, B.DEFAULT_WALKERS[c], ng)))
        got = json.load(open(os.path.join(ROOT, "bench", "configs", "config%d.json" % c)))
        assert got == want, c
    sv = json.load(open(os.path.join(ROOT, "bench", "configs", "config1_svmc.json")))
    assert sv == dict(json.loads(json.dumps(B.workload(1, B.DEFAULT_WALKERS[1], 1))), method="svmc")
