"""Multi-GPU plumbing: an MSM shards by index range; every rank produces one
64-byte partial point, the partials are all-gathered (NCCL over NVLink on GPUs; gloo in the CPU
tests) and added on every rank. No other data-path collective exists on this path."""
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
