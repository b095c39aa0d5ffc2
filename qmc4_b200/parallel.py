"""torch.distributed plumbing for the multi-GPU path (one process per GPU).

torch.distributed is used ONLY to rendezvous and to ship the 128-byte NCCL unique id to every
rank; the data path (E_T all-reduce per step, walker migration) runs inside the C library over
its own NCCL communicator (qmc_comm_init / qmc_rebalance).  The same helpers work with the gloo
backend on CPU for the host-logic tests."""
import os

import numpy as np


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_bytes(payload: bytes, nbytes: int, src: int = 0) -> bytes:
    """rank `src` provides `payload`; every rank returns it (works on nccl and gloo)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def split_walkers(total: int, rank: int, nranks: int) -> int:
    """initial per-rank population: walkers / nProcesses, integer division (dmc/main.cpp:383)"""
    return total // nranks


def make_comm(system_factory):
    """Create the per-rank System with an NCCL communicator spanning all ranks.
    system_factory(comm) -> System, where comm is None (1 rank) or (unique_id, rank, nranks)."""
    rank, world, local = env_rank()
    if world == 1:
        return system_factory(None)
    from . import capi
    import ctypes as C
    uid = b""
    if rank == 0:
        lib = capi.load()
        buf = C.create_string_buffer(128)
        if lib.qmc_comm_unique_id(buf):
            raise RuntimeError(lib.qmc_last_error(None).decode())
        uid = buf.raw
    uid = broadcast_bytes(uid, 128, 0)
    return system_factory((uid, rank, world))


def apply_plan(populations, sources, amounts):
    """populations after executing a rebalance plan (host-side bookkeeping, used by tests)"""
    new = np.asarray(populations).copy()
    for r in range(len(new)):
        new[r] += amounts[r]
        new[sources[r]] -= amounts[r]
    return new
