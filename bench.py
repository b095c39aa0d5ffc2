#!/usr/bin/env python
"""bench.py — DMC walker-moves/s (and pair-evals/s) of the B200 hot path, next to the reference's
CPU path timed on this box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 5|2|3|4|1] [--scaling weak|strong] [--impl reference]

Default workload = BASELINE config 5, the configuration both north-star targets are quoted on: N=512 bosons in a
periodic box, gaussiangaussian pair Jastrow + potentialBarrier, 100 000 walkers over 8 GPUs = 12 500 walkers PER GPU
(weak scaling; `--scaling strong` keeps 100 000 walkers in total).  One bench "step" = DMC_STEPS_PER_STEP[config]
DMC steps (dmcDriver::step + accumulate: move + branching + E_T + estimators + walker migration) over the whole
population, so that one timed unit is >= 50 ms of device work.  Rank 0 prints ONE JSON line; at N=1 it also carries
the other four BASELINE shapes in `secondary` (same measurement, smaller samples).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
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


DEFAULT_CONFIG = 5
DEFAULT_WALKERS = {1: 1000, 2: 16384, 3: 32768, 4: 16384, 5: 12500}
STRONG_TOTAL = {1: 1000, 2: 16384, 3: 32768, 4: 16384, 5: 100000}          # --scaling strong: the BASELINE population, split over the GPUs
DMC_STEPS_PER_STEP = {1: 400, 2: 100, 3: 10, 4: 10, 5: 10}                  # one bench step >= 50 ms of device work
NAMES = {1: "config1 trap N=10 1k walkers", 2: "config2 PBC N=64 16384 walkers/GPU energy+pairCorrelation",
         3: "config3 PBC N=256 32768 walkers/GPU structureFactor", 4: "config4 three-body N=32 16384 walkers/GPU trimerCorrelation",
         5: "config5 PBC N=512 100k walkers / 8 GPUs (12500 walkers/GPU)"}


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


def cpu_sample_shape(cfg):
    """(walkers per process, bytes per walker) of the bounded CPU sample: the reference caches a 32*P-byte distance table per
    walker in two containers (4.2 MB per walker at N=512), so one process is bounded to ~256 MB."""
    pairs, _, _, _ = work_per_move(cfg)
    N = sum(cfg["N"])
    per_walker = 2 * (32.0 * max(pairs, 1) + 200.0 * N) + 4096
    return int(max(8, min(64, 256e6 // per_walker))), per_walker


def describe(cfgno, cfg, wpg, ngpu, scaling, equil, dmc_steps, flushed=True):
    """The `config` object — identical on both arms (bench.py and bench.py --impl reference)."""
    return {"workload": NAMES[cfgno], "N": cfg["N"], "walkers_per_gpu": wpg, "global_walkers": wpg * ngpu, "scaling": scaling,
            "timeStep": cfg["timeStep"], "estimators": [m["kind"] for m in cfg["measurements"]] + ["energy"],
            "sample_every": 1, "equilibration_steps": equil, "dmc_steps_per_bench_step": dmc_steps,
            "cpu_walkers_per_process": cpu_sample_shape(cfg)[0],
            "l2": "flushed between timed bench steps (256 MiB fill)" if flushed else "not flushed (working set >> L2)"}


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
    dropped = ""
    if any(m.get("kind") == "structureFactor" for m in cfg.get("measurements", [])):
        # the reference builds structureFactor for DIMENSIONS == 1 only (structureFactor.h:8): its CPU sample runs without it
        cfg = dict(cfg, measurements=[m for m in cfg["measurements"] if m.get("kind") != "structureFactor"])
        dropped = "; structureFactor measurement dropped (the reference builds it in 1-D only)"
    if os.path.exists(orc.REF_BIN):
        est_us = 0.12 * (pairs + 6 * trip) + 25.0 * N / 10 + 5           # planning estimate per walker-move
        w_cpu, per_walker = cpu_sample_shape(cfg)
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
                "sample": "%d procs x %d walkers x %d steps of the same input.json (seed+core), reference timers totalRunningTime-outputTime; wall %.1fs%s" % (cores, w_cpu, spb * 4, wall, dropped),
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


# ---------------------------------------------------------------------------------------------------
def load_traffic(cfgno):
    """dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the move kernel of this config (profiles/)"""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))[str(cfgno)]["traffic_bytes_per_launch"]
        except Exception:
            continue
    return None


def measure(s, cfg, cfgno, steps, warmup, dmc_steps, equil, flush, barrier, world, rank, torch, dist, sampler=None):
    """Equilibrate, warm up, time `steps` bench steps of `dmc_steps` DMC steps each with CUDA events on the library's
    stream (qmc_step), max over ranks.  Returns the measurement dict of this rank (global quantities where stated)."""
    pairs, trip, flops_per_move, bytes_per_move = work_per_move(cfg)
    verbose = bool(os.environ.get("QMC4_BENCH_VERBOSE"))
    if verbose:         # population of this rank along the equilibration (diagnostics on stderr)
        done = 0
        while done < equil:
            n = min(10, equil - done)
            s.step(n, 1)
            done += n
            b = s.block_results()
            print("[bench rank %d] equil step %d: population %d, E/N %.6g, migrated out/in %d/%d" % (
                rank, done, b["population"], b["energy_sum"] / max(b["energy_count"], 1) / sum(cfg["N"]),
                s.info("migrated_out"), s.info("migrated_in")), file=sys.stderr, flush=True)
    else:
        s.step(equil, 1)
    s.block_results()
    for _ in range(warmup):
        if flush is not None:
            flush.fill_(1)
            torch.cuda.synchronize()
        s.step(dmc_steps, 1)
    s.block_results()
    s.info("propagate_timers_reset")
    # launch-bound shape (config 1: two ~10 us kernels per DMC step): the timed region replays CUDA graphs of step pairs, which
    # per-launch event pairs would prevent; the move kernel is timed in a separate pass right after the timed region instead
    separate_timing = cfgno == 1
    if not separate_timing:
        s.info("time_kernels_on")
    l0 = s.info("launches")
    m0 = (s.info("migration_events"), s.info("migrated_out"), s.info("migrated_in"))
    if sampler is not None:
        sampler.start()
    barrier()
    for _ in range(steps):
        if flush is not None:
            flush.fill_(1)                      # L2 flush: 256 MiB written between timed bench steps
            torch.cuda.synchronize()
        s.step(dmc_steps, 1)                    # CUDA events around the DMC steps on the library's stream
    barrier()
    b = s.block_results()                       # n_moves etc. are global (all-reduced) with a communicator
    clocks = sampler.stop() if sampler is not None else None
    launches = s.info("launches") - l0
    graph_replays = s.info("graph_replays")
    if separate_timing:
        s.info("time_kernels_on")
        s.step(min(dmc_steps, 200), 1)
        s.block_results()
    prop_ms = s.info("propagate_ms_total") / max(1.0, s.info("propagate_launches_timed"))
    s.info("time_kernels_off")
    ms_total = b["device_ms"]
    local_pop = b["population"]
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    moves = b["n_moves"]
    value = moves / (ms_total * 1e-3)
    # the move kernel of THIS rank processed about its own population per launch
    moves_per_launch = float(local_pop)
    achieved = moves_per_launch * flops_per_move / (prop_ms * 1e-3) * 1e-12 if prop_ms > 0 else None
    out = {"value": value, "ms_per_step": ms_total / steps, "ms_per_dmc_step": ms_total / steps / dmc_steps, "moves": moves,
           "population": local_pop, "acceptance": b["n_accepted"] / max(b["n_moves"], 1),
           "energy_per_particle": b["energy_sum"] / max(b["energy_count"], 1) / sum(cfg["N"]), "launches": int(launches),
           "prop_ms": prop_ms, "achieved_tflops": achieved, "moves_per_launch": moves_per_launch, "clocks": clocks,
           "graph_replays": graph_replays, "launch_timing": "separate pass after the timed region" if separate_timing else "event pairs inside the timed region",
           "pair_evals_per_s": value * pairs, "triplet_evals_per_s": value * trip,
           "migration": {"events": s.info("migration_events") - m0[0], "walkers_out": s.info("migrated_out") - m0[1],
                         "walkers_in": s.info("migrated_in") - m0[2], "every": s.info("migrate_every"), "band": s.info("migrate_band")}}
    return out


def roofline_block(s, cfg, cfgno, m, peak, hbm_peak):
    pairs, trip, flops_per_move, bytes_per_move = work_per_move(cfg)
    step_flops = m["moves"] * flops_per_move            # whole timed region, all ranks
    r = {"bound": "fp64" if cfgno != 1 else "hbm", "kernel": "move_kernel (drift-diffusion + pair/triplet evaluation + Green ratio + Metropolis)",
         "achieved": m["achieved_tflops"], "peak": peak, "unit": "TFLOP/s",
         "frac": (m["achieved_tflops"] / peak) if (peak and m["achieved_tflops"]) else None, "traffic": load_traffic(cfgno),
         "algorithmic_flop_per_launch": m["moves_per_launch"] * flops_per_move, "avg_launch_ms": m["prop_ms"], "launch_timing": m["launch_timing"],
         "flop_per_pair": FLOP_PER_PAIR, "flop_per_coordinate": FLOP_PER_COORD, "flop_per_triplet": FLOP_PER_TRIPLET,
         "peak_source": "FP64 DFMA pipe measured in-process (qmc_get_info fp64_peak_tflops: 8 independent DFMA chains/thread, 148x8 CTAs); "
                        "MEASURED_PEAKS.json carries HBM and bf16 only",
         "hbm": {"achieved_gbs": m["moves_per_launch"] * bytes_per_move / (m["prop_ms"] * 1e-3) * 1e-9 if m["prop_ms"] > 0 else None,
                 "peak_gbs": hbm_peak, "algorithmic_bytes_per_launch": m["moves_per_launch"] * bytes_per_move}}
    if cfgno == 1:       # latency/HBM-bound shape: report the HBM fraction as the roofline, FP64 beside it
        r.update(achieved=r["hbm"]["achieved_gbs"], peak=hbm_peak, unit="GB/s",
                 frac=(r["hbm"]["achieved_gbs"] / hbm_peak) if (hbm_peak and r["hbm"]["achieved_gbs"]) else None,
                 fp64_frac=(m["achieved_tflops"] / peak) if (peak and m["achieved_tflops"]) else None,
                 peak_source="MEASURED_PEAKS.json hbm_gbs")
    return r


def run_config(cfgno, args, torch, dist, parallel, System, rank, world, local, steps, warmup, primary, method=None):
    ngpu = max(args.gpus, world)
    dmc_steps = args.dmc_steps or DMC_STEPS_PER_STEP[cfgno]
    if args.scaling == "strong":
        wpg = max(1, STRONG_TOTAL[cfgno] // ngpu)
    else:
        wpg = DEFAULT_WALKERS[cfgno]
    if primary and args.walkers:
        wpg = args.walkers
    equil = args.equil if primary else min(args.equil, 100)
    cfg = workload(cfgno, wpg, ngpu)
    if method:
        cfg["method"] = method          # "svmc": the same walkers without branching (main.cpp:393-395)
    config = describe(cfgno, cfg, wpg, ngpu, args.scaling, equil, dmc_steps, flushed=not args.no_flush)
    config["method"] = cfg["method"]
    s = parallel.make_comm(lambda comm: System.from_input(cfg, device=local, comm=comm))
    if args.migrate_every or args.migrate_band is not None:
        s.set_migration(args.migrate_every or 1, 0.02 if args.migrate_band is None else args.migrate_band)
    dofs = s.dofs
    s.upload(initial_walkers(cfg, wpg, 1000 + rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(local) if (rank == 0 and primary and not os.environ.get("QMC4_BENCH_NO_SAMPLER")) else None
    m = measure(s, cfg, cfgno, steps, warmup, dmc_steps, equil, flush, barrier, world, rank, torch, dist, sampler)
    out = {"config": config, "m": m, "cfg": cfg}

    if primary:
        # ---- conservation checks of the migration (N > 1): the ranks' populations add up to the all-reduced one; an explicit
        #      qmc_rebalance leaves the global multiset fingerprint unchanged
        if world > 1:
            pop = torch.tensor([m["population"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(pop)
            pops = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(pops, torch.tensor([m["population"]], dtype=torch.float64, device="cuda"))
            mig = torch.tensor([m["migration"]["events"], m["migration"]["walkers_out"], m["migration"]["walkers_in"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(mig)
            ck0 = torch.tensor([int(s.info("population_checksum"))], dtype=torch.int64, device="cuda")
            dist.all_reduce(ck0)
            moved = s.rebalance()
            ck1 = torch.tensor([int(s.info("population_checksum"))], dtype=torch.int64, device="cuda")
            dist.all_reduce(ck1)
            mv = torch.tensor([abs(moved)], dtype=torch.float64, device="cuda")
            dist.all_reduce(mv)
            out["conservation"] = {
                "sum_of_rank_populations": float(pop.item()), "all_reduced_population": s.info("global_population"),
                "rank_populations": [float(p.item()) for p in pops],
                "in_step_migration": {"events_sum_over_ranks": float(mig[0].item()), "walkers_out": float(mig[1].item()),
                                      "walkers_in": float(mig[2].item()), "every": m["migration"]["every"], "band": m["migration"]["band"]},
                "explicit_rebalance": {"walkers_moved": float(mv.item()) / 2, "multiset_fingerprint_before": int(ck0.item()) % (1 << 52),
                                       "multiset_fingerprint_after": int(ck1.item()) % (1 << 52)}}
            # frozen-population check of the IN-STEP migration: with timeStep = 1e-300 a move changes nothing in fp64 and every
            # walker has exactly one descendant, so only the migration moves walkers (unequal start, dead band 0): the global
            # multiset fingerprint must not change while the populations equalise
            starts = [max(8, min(wpg, 2048) - 37 * r) for r in range(world)]
            # target population = what is uploaded and a window wider than it: E_T never gets the population-control term
            sz = parallel.make_comm(lambda comm: System.from_input(dict(cfg, timeStep=1e-300, walkers=sum(starts), delta_walkers=sum(starts)),
                                                                   device=local, comm=comm))
            sz.set_migration(1, 0.0)
            nz = starts[rank]
            sz.upload(initial_walkers(cfg, nz, 5000 + rank))
            sz.step(1, 1)
            fz0 = torch.tensor([int(sz.info("population_checksum"))], dtype=torch.int64, device="cuda")
            dist.all_reduce(fz0)
            sz.step(4, 1)
            fz1 = torch.tensor([int(sz.info("population_checksum"))], dtype=torch.int64, device="cuda")
            dist.all_reduce(fz1)
            pz = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(pz, torch.tensor([sz.block_results()["population"]], dtype=torch.float64, device="cuda"))
            mz = torch.tensor([sz.info("migrated_out"), sz.info("migrated_in")], dtype=torch.float64, device="cuda")
            dist.all_reduce(mz)
            sz.close()
            pz = [float(p.item()) for p in pz]
            out["conservation"]["frozen_population_in_step"] = {
                "start_populations": starts, "end_populations": pz,
                "walkers_out": float(mz[0].item()), "walkers_in": float(mz[1].item()),
                "multiset_fingerprint_before": int(fz0.item()) % (1 << 52), "multiset_fingerprint_after": int(fz1.item()) % (1 << 52)}
            fz_ok = (int(fz0.item()) % (1 << 52) == int(fz1.item()) % (1 << 52) and max(pz) - min(pz) <= 1 and mz[0].item() == mz[1].item() > 0)
            out["conservation"]["ok"] = fz_ok and bool(out["conservation"]["sum_of_rank_populations"] == out["conservation"]["all_reduced_population"]
                                             and mig[1].item() == mig[2].item()
                                             and out["conservation"]["explicit_rebalance"]["multiset_fingerprint_before"]
                                             == out["conservation"]["explicit_rebalance"]["multiset_fingerprint_after"])
        # ---- e2e: the same bench step through the C ABI with HOST buffers (pinned), copies inside the timed region
        e2e_steps = max(3, min(steps, 5))
        Rh, _ = s.download(with_energy=True)
        nW = len(Rh)
        pin_in = torch.from_numpy(Rh.copy()).pin_memory()
        capw = int(s.info("capacity"))
        pin_out = torch.empty((capw, dofs), dtype=torch.float64).pin_memory()
        pin_E = torch.empty(capw, dtype=torch.float64).pin_memory()
        import ctypes as C
        from qmc4_b200 import capi
        lib = s.lib
        barrier()
        t0 = time.perf_counter()
        e2e_moves = 0
        for _ in range(e2e_steps):
            rc = lib.qmc_upload_walkers(s.ctx, C.cast(pin_in.data_ptr(), capi.c_double_p), nW)
            assert rc == 0, lib.qmc_last_error(s.ctx)
            s.step(dmc_steps, 1)
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
        out["e2e"] = {"value": e2e_moves / float(t.item()), "unit": "walker-moves/s", "h2d_bytes_per_step": int(nW * dofs * 8),
                      "d2h_bytes_per_step": int(nW * (dofs + 1) * 8 + 72),
                      "note": "per bench step: qmc_upload_walkers(pinned host R) [+ initial evaluation] -> qmc_step(%d DMC steps) -> qmc_block_results -> qmc_download_walkers(host R,E)" % dmc_steps}
    if rank == 0:
        out["kernel"] = {"move_kernel": "pair_move_kernel (kernels_pair.cu)" if s.info("pair_kernel") else
                         "small_move_kernel (kernels_small.cu, %d walkers per warp)" % s.info("small_walkers_per_warp") if s.info("small_kernel") else
                         "move_kernel (kernels_v2.cu)", "grid": s.info("grid"), "smem_bytes": s.info("smem_bytes"),
                         "regs": s.info("regs_propagate"), "pipeline": s.info("pipeline"),
                         "warps_per_walker": s.info("pair_kernel_warps_per_walker") if s.info("pair_kernel") else s.info("warps_per_walker")}
        if primary:
            out["peak"] = s.info("fp64_peak_tflops")
    s.close()
    del flush
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=DEFAULT_CONFIG)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: BASELINE size)")
    ap.add_argument("--equil", type=int, default=200, help="untimed DMC steps before warm-up (SURVEY §8d)")
    ap.add_argument("--dmc-steps", type=int, default=0, help="DMC steps per bench step (default: per config, >= 50 ms of device work)")
    ap.add_argument("--migrate-every", type=int, default=0)
    ap.add_argument("--migrate-band", type=float, default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--watchdog", type=int, default=1500, help="seconds after which a hung run dumps every thread's stack to stderr and exits")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    steps, warmup = args.steps, max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ngpu = max(args.gpus, world)
    dmc_steps = args.dmc_steps or DMC_STEPS_PER_STEP[args.config]
    wpg = args.walkers or (max(1, STRONG_TOTAL[args.config] // ngpu) if args.scaling == "strong" else DEFAULT_WALKERS[args.config])
    cfg = workload(args.config, wpg, ngpu)
    pairs, trip, flops_per_move, bytes_per_move = work_per_move(cfg)
    config = describe(args.config, cfg, wpg, ngpu, args.scaling, args.equil, dmc_steps, flushed=not args.no_flush)

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(cfg, seconds_budget=12.0 * max(1, min(steps, 10)) / 10)
        line = {"impl": "reference", "metric": "dmc_walker_moves_per_sec", "value": cb["value"], "unit": "walker-moves/s",
                "n_gpus": ngpu, "steps": steps, "warmup": warmup, "ms_per_step": None, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
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
    res = run_config(args.config, args, torch, dist, parallel, System, rank, world, local, steps, warmup, primary=True)
    m = res["m"]
    if rank != 0:
        return
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    peak = res["peak"]
    s_dummy = None
    roofline = roofline_block(s_dummy, cfg, args.config, m, peak, hbm_peak)
    # the same algorithmic flop count over the WHOLE step (scan, estimators, migration, collectives included), all ranks
    roofline["whole_step_frac"] = (m["moves"] * flops_per_move / (m["ms_per_step"] * steps * 1e-3) * 1e-12) / (peak * world) if peak else None
    roofline["fp64_peak_sm_mhz"] = m["clocks"]["sm_mhz"] if m["clocks"] else None
    line = {"metric": "dmc_walker_moves_per_sec", "value": m["value"], "unit": "walker-moves/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": res["config"], "ms_per_dmc_step": m["ms_per_dmc_step"],
            "pair_evals_per_s": m["pair_evals_per_s"], "triplet_evals_per_s": m["triplet_evals_per_s"], "population": m["population"],
            "acceptance": m["acceptance"], "energy_per_particle": m["energy_per_particle"],
            "gpu_launches": m["launches"], "clocks": m["clocks"], "e2e": res["e2e"], "roofline": roofline, "kernel": res["kernel"]}
    if "conservation" in res:
        line["migration"] = res["conservation"]
    if world == 1 and not args.no_secondary and args.config == DEFAULT_CONFIG:
        # the other BASELINE shapes, same measurement on smaller samples (N=1 only; parity for them is in tests/)
        sec = {}
        for c2, meth in ((2, None), (3, None), (4, None), (1, None), (1, "svmc")):
            key = "config%d" % c2 + ("_" + meth if meth else "")
            try:
                r2 = run_config(c2, args, torch, dist, parallel, System, rank, world, local, max(3, min(steps, 5)), 3, primary=False, method=meth)
                m2 = r2["m"]
                rf = roofline_block(None, r2["cfg"], c2, m2, peak, hbm_peak)
                sec[key] = {"config": r2["config"], "value": m2["value"], "unit": "walker-moves/s", "ms_per_step": m2["ms_per_step"],
                                        "ms_per_dmc_step": m2["ms_per_dmc_step"], "pair_evals_per_s": m2["pair_evals_per_s"],
                                        "triplet_evals_per_s": m2["triplet_evals_per_s"], "acceptance": m2["acceptance"],
                                        "population": m2["population"], "gpu_launches": m2["launches"], "cuda_graph_replays": m2["graph_replays"],
                                        "roofline": rf, "kernel": r2["kernel"]}
                if not args.no_cpu_baseline:
                    sec[key]["cpu_baseline"] = cpu_reference_arm(r2["cfg"], seconds_budget=10.0)
            except Exception as e:          # a secondary shape must never take the headline down
                sec[key] = {"error": repr(e)[:300]}
        line["secondary"] = sec
    if not args.no_cpu_baseline and world == 1:          # rank 0 at N=1 only (the N>1 runs compare against the same number)
        line["cpu_baseline"] = cpu_reference_arm(cfg, seconds_budget=12.0)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
