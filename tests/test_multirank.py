"""CPU tests (-m "not gpu") of the N>1 host logic with torch.distributed/gloo, world_size 2:
unique-id distribution, initial population split and one rebalance round executed with real
point-to-point messages following the plan the C ABI computes (qmc_plan_rebalance)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from qmc4_b200 import parallel
    from qmc4_b200.system import plan_rebalance
    r, w, _ = parallel.init_process_group("gloo")
    assert (r, w) == (rank, world)
    # (1) the 128-byte communicator id created on rank 0 reaches every rank unchanged
    uid = bytes(range(128)) if rank == 0 else b""
    got = parallel.broadcast_bytes(uid, 128, 0)
    assert got == bytes(range(128))
    # (2) initial split: walkers / nProcesses (main.cpp:383)
    assert parallel.split_walkers(1001, rank, world) == 500
    # (3) one rebalance round: allgather populations, plan, ship TAIL walkers, append on the receiver
    dofs = 6
    mine = [70, 30][rank]
    walkers = np.arange(mine * dofs, dtype=np.float64).reshape(mine, dofs) + 1000 * rank
    pops = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(pops, torch.tensor([mine]))
    pops = [int(p.item()) for p in pops]
    src, amt = plan_rebalance(pops)
    reqs = []
    sent = 0
    for dst in range(world):
        if amt[dst] > 0 and src[dst] == rank:
            tail = torch.from_numpy(walkers[mine - sent - amt[dst]: mine - sent].copy())
            reqs.append(dist.isend(tail, dst))
            sent += amt[dst]
    recv = None
    if amt[rank] > 0:
        recv = torch.zeros(int(amt[rank]), dofs, dtype=torch.float64)
        dist.recv(recv, int(src[rank]))
    for rq in reqs:
        rq.wait()
    walkers = walkers[: mine - sent]
    if recv is not None:
        walkers = np.concatenate([walkers, recv.numpy()])
    new_pops = parallel.apply_plan(pops, src, amt)
    assert len(walkers) == new_pops[rank]
    tot = torch.tensor([float(walkers.sum()), float(len(walkers))], dtype=torch.float64)
    dist.all_reduce(tot)
    q.put((rank, len(walkers), float(tot[0]), float(tot[1])))
    dist.destroy_process_group()


def test_two_rank_rebalance_with_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert [r[1] for r in res] == [50, 50]                    # balanced
    total = sum(np.arange(n * 6).sum() + 1000 * rk * n * 6 for rk, n in ((0, 70), (1, 30)))
    assert res[0][2] == res[1][2] == total                    # no walker lost or duplicated
    assert res[0][3] == 100
