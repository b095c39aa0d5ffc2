#!/usr/bin/env python
"""bench.py — DMC walker-moves/s (and pair-evals/s) of the B200 hot path, next to the reference's
CPU path timed on this box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|5|1|4] [--impl reference]

One "step" = one DMC step (dmcDriver::step + accumulate: propagate + branching + E_T + estimators)
over the rank's whole walker population.  Default workload: BASELINE config 2 (N=64, periodic box,
gaussiangaussian pair Jastrow + potentialBarrier, 16384 walkers PER GPU (weak scaling), energy +
pairCorrelation estimators every step).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 64.0        # algorithmic FP64 flop per pair evaluation
FLOP_PER_COORD = 16.0       # drift, diffusion, Green ratio, |grad|^2
FLOP_PER_TRIPLET = 59.0


def workload(cfgno, walkers_per_gpu, ngpu):
    """The reference-schema input.json of BASELINE config `cfgno`."""
    W = walkers_per_gpu * ngpu
    base = {"method": "dmc", "timeStep": 1e-3, "walkers": W, "delta_walkers": int(0.1 * W), "stepsPerBlock": 100,
            "nBlocks": 1, "correlationSteps": 0, "seed": 1234 + cfgno, "measurements": []}
    gg = {"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "gaussiangaussian", "r0": 1.0, "u0": -0.5}}
    barrier = {"kind": "potentialTwoBody", "sets": [0, 0], "functor": {"kind": "potentialBarrier", "R0": 1.0, "V0": 2.0}}
    if cfgno == 1:
        base.update(N=[10], lBox=[1e4] * 3,
                    wavefunctions=[{"kind": "jastrow1b", "set": 0, "jastrow": {"kind": "gaussian", "alpha": 0.5}},
                                   {"kind": "jastrow2bUnDis", "sets": [0, 0], "jastrow": {"kind": "gaussian", "alpha": 0.05}}],
                    potentials=[{"kind": "harmonic1b", "omega": 1.0, "set": 0}])
    elif cfgno in (2, 3, 5):
        N = {2: 64, 3: 256, 5: 512}[cfgno]
        L = {2: 4.0, 3: 6.3496, 5: 8.0}[cfgno]
        base.update(N=[N], lBox=[L] * 3, wavefunctions=[gg], potentials=[barrier])
        if cfgno == 2:
            base["measurements"] = [{"kind": "pairCorrelation", "label": "pair", "sets": [0, 0], "bins": 1000, "minx": 0, "maxx": L / 2}]
        if cfgno == 3:
            base["measurements"] = [{"kind": "structureFactor", "label": "sq", "sets": [0, 0], "bins": 64, "maxq": 64 * 2 * np.pi / L, "lBox": L, "channel": 1}]
    elif cfgno == 4:
        base.update(N=[32], lBox=[5.0] * 3,
                    wavefunctions=[gg, {"kind": "jastrow3bUnDis", "sets": [0], "jastrow": {"kind": "gaussiangaussian", "r0": 1.5, "u0": -0.2}}],
                    potentials=[{"kind": "potentialThreeBodyUnDis", "sets": [0, 0, 0], "functor": {"kind": "squareWell", "R0": 1.5, "V0": 1.0}}],
                    measurements=[{"kind": "trimerCorrelationUnDis", "label": "trimer", "sets": [0], "bins": 500, "minx": 0, "maxx": 2.5}])
    else:
        raise SystemExit("unknown config")
    return base


DEFAULT_WALKERS = {1: 1000, 2: 16384, 3: 32768, 4: 16384, 5: 12500}
NAMES = {1: "config1 trap N=10 1k walkers", 2: "config2 PBC N=64 16384 walkers/GPU energy+pairCorrelation",
         3: "config3 PBC N=256 32768 walkers/GPU structureFactor", 4: "config4 three-body N=32 16384 walkers/GPU trimerCorrelation",
         5: "config5 PBC N=512 12500 walkers/GPU"}


def work_per_move(cfg):
    N = sum(cfg["N"])
    pairs = N * (N - 1) / 2 if any(w["kind"].startswith("jastrow2b") for w in cfg["wavefunctions"]) else 0
    trip = N * (N - 1) * (N - 2) / 6 if any(w["kind"].startswith("jastrow3b") for w in cfg["wavefunctions"]) else 0
    flops = FLOP_PER_PAIR * pairs + FLOP_PER_COORD * 3 * N + FLOP_PER_TRIPLET * trip
    bytes_ = 96.0 * N + 48.0
    return pairs, trip, flops, bytes_


def initial_walkers(cfg, nW, seed):
    rng = np.random.default_rng(seed)
    N = sum(cfg["N"])
    L = np.asarray(cfg["lBox"])
    if L[0] > 1e3:
        return rng.normal(0, np.sqrt(0.5), size=(nW, 3 * N))
    return (rng.uniform(-0.5, 0.5, size=(nW, 3, N)) * L[None, :, None]).reshape(nW, 3 * N)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms DURING the timed region."""

    def __init__(self, index):
        self.proc, self.index = None, index
        self.path = os.path.join(tempfile.gettempdir(), "qmc4_clocks_%d_%d.csv" % (os.getpid(), index))

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-f", self.path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            time.sleep(0.3)
        except OSError:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            try:
                rows = [[x.strip() for x in line.split(",")] for line in open(self.path) if line.strip()]
                os.remove(self.path)
            except OSError:
                rows = []

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if len(r) > 1 and num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if len(r) > 2 and num(r[2]) is not None]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples under load = upper half of the observed SM clocks (the sampler also sees the idle gaps between steps)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_arm(cfg, seconds_budget, max_procs=None):
    """The reference's own CPU implementation (oracle/_ref/dmc_ref = unmodified dmc/main.cpp + dmc_lib),
    one single-rank process per core with seed+core — its own decomposition (main.cpp:383,391) —
    on a bounded sample of the workload.  Falls back to the C restatement when the binary is absent."""
    from oracle import oracle as orc
    cpu_ids = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    cores = len(cpu_ids)
    if max_procs:
        cores = min(cores, max_procs)
    N = sum(cfg["N"])
    pairs, trip, _, _ = work_per_move(cfg)
    if os.path.exists(orc.REF_BIN):
        est_us = 0.12 * (pairs + 6 * trip) + 25.0 * N / 10 + 5           # planning estimate per walker-move
        # the reference caches a 32*P-byte distance table per walker in two containers (4.2 MB per
        # walker at N=512): bound one process to ~256 MB and all processes to a quarter of the free host memory
        per_walker = 2 * (32.0 * max(pairs, 1) + 200.0 * N) + 4096
        w_cpu = int(max(8, min(64, 256e6 // per_walker)))
        try:
            avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
            cores = int(max(1, min(cores, 0.25 * avail // (per_walker * w_cpu * 1.5 + 64e6))))
        except (OSError, IndexError, ValueError):
            pass
        steps = int(max(4, min(400, seconds_budget * 1e6 / (est_us * w_cpu))))
        spb = max(1, steps // 4)
        c = dict(cfg)
        c.update(walkers=w_cpu, delta_walkers=max(1, w_cpu // 10), stepsPerBlock=spb, nBlocks=4, correlationSteps=0)
        c.pop("configurations", None)
        tmp = tempfile.mkdtemp(prefix="qmc4ref_")
        procs = []
        t0 = time.time()
        for k in range(cores):
            d = os.path.join(tmp, "r%d" % k)
            os.makedirs(d)
            ck = dict(c, seed=int(c["seed"]) + k)
            json.dump(ck, open(os.path.join(d, "input.json"), "w"))
            cmd = [orc.REF_BIN, "input.json"]
            if hasattr(os, "sched_setaffinity"):
                cmd = ["taskset", "-c", str(cpu_ids[k])] + cmd
            procs.append(subprocess.Popen(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
        moves, rate = 0.0, 0.0
        for p in procs:
            out = re.sub(r"\x1b\[[0-9;]*m", "", p.communicate()[0])
            pops = [float(l.split()[-1]) for l in out.splitlines() if "Curr. Walkers" in l]
            run_ms = [float(l.split()[2]) for l in out.splitlines() if l.startswith("totalRunningTime")]
            out_ms = [float(l.split()[2]) for l in out.splitlines() if l.startswith("outputTime")]
            if not pops or not run_ms:
                continue
            traj = [w_cpu] + pops
            mv = sum(0.5 * (a + b) * spb for a, b in zip(traj[:-1], traj[1:]))
            ms = max(run_ms[0] - (out_ms[0] if out_ms else 0.0), 1.0)       # driver.cpp:19-20,45-57 timers (1 ms resolution)
            moves += mv
            rate += mv / (ms * 1e-3)
        wall = time.time() - t0
        return {"value": rate, "unit": "walker-moves/s", "cores": cores, "kind": "reference",
                "sample": "%d procs x %d walkers x %d steps of the same input.json (seed+core), reference timers totalRunningTime-outputTime; wall %.1fs" % (cores, w_cpu, spb * 4, wall),
                "pair_evals_per_s": rate * pairs}
    port = orc.Port()
    nW = 32
    R = initial_walkers(cfg, nW, 1)
    xi = np.random.default_rng(2).normal(size=R.shape)
    u = np.random.default_rng(3).uniform(size=nW)
    t0 = time.time()
    reps = 0
    while time.time() - t0 < max(2.0, seconds_budget / 3):
        port.update(cfg, cfg["timeStep"], R, xi, u)
        reps += 1
    dt = time.time() - t0
    rate = reps * nW / dt / 2.0     # update() evaluates old and new configuration: 2 evaluations per call
    return {"value": rate * 2.0, "unit": "walker-moves/s", "cores": 1, "kind": "port",
            "sample": "%d x dmcDriver::update on %d walkers through oracle/qmc_oracle.c (scalar, 1 thread)" % (reps, nW),
            "pair_evals_per_s": rate * 2.0 * pairs}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: BASELINE size)")
    ap.add_argument("--equil", type=int, default=200, help="untimed DMC steps before warm-up (SURVEY §8d)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--watchdog", type=int, default=1500, help="seconds after which a hung run dumps every thread's stack to stderr and exits")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    steps, warmup = args.steps, max(args.warmup, 3)
    wpg = args.walkers or DEFAULT_WALKERS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ngpu = max(args.gpus, world)
    cfg = workload(args.config, wpg, ngpu)
    pairs, trip, flops_per_move, bytes_per_move = work_per_move(cfg)
    config = {"workload": NAMES[args.config], "N": cfg["N"], "walkers_per_gpu": wpg, "global_walkers": wpg * ngpu,
              "timeStep": cfg["timeStep"], "estimators": [m["kind"] for m in cfg["measurements"]] + ["energy"],
              "sample_every": 1, "equilibration_steps": args.equil}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(cfg, seconds_budget=12.0 * max(1, min(steps, 10)) / 10)
        line = {"impl": "reference", "metric": "dmc_walker_moves_per_sec", "value": cb["value"], "unit": "walker-moves/s",
                "n_gpus": ngpu, "steps": steps, "warmup": warmup, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "pair_evals_per_s": cb["pair_evals_per_s"], "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "walker-moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from qmc4_b200 import parallel
    from qmc4_b200.system import System
    rank, world, local = parallel.init_process_group()
    torch.cuda.set_device(local)
    s = parallel.make_comm(lambda comm: System.from_input(cfg, device=local, comm=comm))
    dofs = s.dofs
    R0 = initial_walkers(cfg, wpg, 1000 + rank)
    s.upload(R0)
    s.step(args.equil, 1)
    s.block_results()
    if world > 1:
        s.rebalance()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        if flush is not None:
            flush.fill_(1)
            torch.cuda.synchronize()
        s.step(1, 1)
    s.block_results()
    s.info("propagate_timers_reset")
    s.info("time_kernels_on")
    l0 = s.info("launches")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    moves = 0.0
    ms_total = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.fill_(1)                      # L2 flush: 256 MiB written between timed steps
            torch.cuda.synchronize()
        s.step(1, 1)                            # CUDA events around the step on the library's stream
    barrier()
    b = s.block_results()
    clocks = sampler.stop() if rank == 0 else None
    ms_total, moves = b["device_ms"], b["n_moves"]
    launches = s.info("launches") - l0
    prop_ms = s.info("propagate_ms_total") / max(1.0, s.info("propagate_launches_timed"))
    s.info("time_kernels_off")
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # moves are already global (block results are all-reduced)
    ms_total = float(t.item())
    value = moves / (ms_total * 1e-3)
    pop = b["population"]

    # ---- e2e: the same step through the C ABI with HOST buffers (pinned), copies inside the timed region
    e2e_steps = max(3, min(steps, 10))
    Rh, _ = s.download(with_energy=True)
    nW = len(Rh)
    pin_in = torch.from_numpy(Rh.copy()).pin_memory()
    pin_out = torch.empty((int(s.info("capacity")), dofs), dtype=torch.float64).pin_memory()
    pin_E = torch.empty(int(s.info("capacity")), dtype=torch.float64).pin_memory()
    import ctypes as C
    from qmc4_b200 import capi
    lib = s.lib
    barrier()
    t0 = time.perf_counter()
    e2e_moves = 0
    for _ in range(e2e_steps):
        rc = lib.qmc_upload_walkers(s.ctx, C.cast(pin_in.data_ptr(), capi.c_double_p), nW)
        assert rc == 0, lib.qmc_last_error(s.ctx)
        s.step(1, 1)
        br = s.block_results()
        n = C.c_int64(0)
        rc = lib.qmc_download_walkers(s.ctx, C.cast(pin_out.data_ptr(), capi.c_double_p), pin_out.shape[0], C.byref(n),
                                      C.cast(pin_E.data_ptr(), capi.c_double_p))
        assert rc == 0, lib.qmc_last_error(s.ctx)
        e2e_moves += br["n_moves"]
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = {"value": e2e_moves / float(t.item()), "unit": "walker-moves/s", "h2d_bytes_per_step": int(nW * dofs * 8),
           "d2h_bytes_per_step": int(nW * (dofs + 1) * 8 + 72),
           "note": "per step: qmc_upload_walkers(pinned host R) [+ initial evaluation] -> qmc_step(1) -> qmc_block_results -> qmc_download_walkers(host R,E)"}

    if rank != 0:
        s.close()
        return
    peak = s.info("fp64_peak_tflops")
    moves_per_launch = moves / steps / world
    achieved = moves_per_launch * flops_per_move / (prop_ms * 1e-3) * 1e-12
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel/config (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[str(args.config)]["traffic_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "kernel": "propagate_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "algorithmic_flop_per_launch": moves_per_launch * flops_per_move, "avg_launch_ms": prop_ms,
                "peak_source": "FP64 DFMA pipe measured in-process (qmc_get_info fp64_peak_tflops); MEASURED_PEAKS.json carries HBM and bf16 only",
                "hbm": {"achieved_gbs": moves_per_launch * bytes_per_move / (prop_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                        "algorithmic_bytes_per_launch": moves_per_launch * bytes_per_move}}
    line = {"metric": "dmc_walker_moves_per_sec", "value": value, "unit": "walker-moves/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(config, l2="flushed between timed steps (256 MiB fill)" if flush is not None else "not flushed"),
            "pair_evals_per_s": value * pairs, "triplet_evals_per_s": value * trip, "population": pop,
            "acceptance": b["n_accepted"] / max(b["n_moves"], 1), "energy_per_particle": b["energy_sum"] / max(b["energy_count"], 1) / sum(cfg["N"]),
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "kernel": {"warps_per_walker": s.info("warps_per_walker"), "grid": s.info("grid"), "smem_bytes": s.info("smem_bytes"),
                       "regs": s.info("regs_propagate"), "variant": s.info("kernel_variant")}}
    s.close()
    if not args.no_cpu_baseline and world == 1:          # rank 0 at N=1 only (the N>1 runs compare against the same number)
        line["cpu_baseline"] = cpu_reference_arm(cfg, seconds_budget=12.0)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
