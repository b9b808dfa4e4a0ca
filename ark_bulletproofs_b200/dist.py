"""Multi-GPU plumbing.

Standalone MSMs shard by index range: every rank produces one 64-byte partial point, the partials are
all-gathered (NCCL over NVLink on GPUs; gloo in the CPU tests) and added on every rank.

Proving / verification shard the generators cyclically (bp_ctx_set_collective): the library calls back into
`torch_allgather` (or `ThreadGroup.allgather` for several contexts inside one process) with the 64-byte
partial points of each sharded MSM. No other data-path collective exists on this path."""
import threading
import ctypes

import torch
import torch.distributed as dist

from . import _lib, codec


def allgather_sum_points(curve: str, raw: bytes, ident: bool, device=None, group=None):
    """raw: this rank's 64-byte affine partial (ignored when ident). Returns (raw_sum, is_identity)."""
    world = dist.get_world_size(group)
    mine = torch.frombuffer(bytearray((bytes(64) if ident else raw)), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    allb = b"".join(t.cpu().numpy().tobytes() for t in bufs)       # identity partials are (0,0) and are skipped by the adder
    out = ctypes.create_string_buffer(64)
    idn = ctypes.c_int(0)
    rc = _lib.load().bp_points_sum_curve(codec.CURVE_IDS[curve], allb, world, out, ctypes.byref(idn))
    if rc != 0:
        raise _lib.BpError(rc, "bp_points_sum_curve")
    return out.raw, bool(idn.value)


def shard_range(n: int, rank: int, world: int):
    """Contiguous index range of `rank` when n terms are split over `world` ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def torch_allgather(device=None, group=None):
    """all-gather callback for Context.set_collective over torch.distributed (NCCL when `device` is a CUDA device)."""
    world = dist.get_world_size(group)

    def allgather(send: bytes) -> bytes:
        mine = torch.frombuffer(bytearray(send), dtype=torch.uint8)
        if device is not None:
            mine = mine.to(device)
        out = torch.empty(world * len(send), dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out.cpu().numpy().tobytes()
    return allgather


class ThreadGroup:
    """In-process stand-in for a process group: `world` threads, one bp_ctx each (possibly on the same GPU),
    exchanging through a barrier. Used by the single-GPU tests of the sharded prover."""

    def __init__(self, world: int):
        self.world = world
        self.slots = [b""] * world
        self.barrier = threading.Barrier(world)

    def allgather_for(self, rank: int):
        def allgather(send: bytes) -> bytes:
            self.slots[rank] = send
            self.barrier.wait(timeout=120)
            out = b"".join(self.slots)
            self.barrier.wait(timeout=120)      # nobody overwrites a slot before everyone has read it
            return out
        return allgather

    def run(self, fn):
        """fn(rank, allgather) on `world` threads; returns the list of results, re-raises the first exception."""
        res, err = [None] * self.world, [None] * self.world

        def work(r):
            try:
                res[r] = fn(r, self.allgather_for(r))
            except BaseException as e:      # noqa: BLE001
                err[r] = e
                self.barrier.abort()
        th = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        return res
