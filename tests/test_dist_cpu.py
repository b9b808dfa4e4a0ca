"""world_size-2 gloo test (CPU) of the N>1 MSM path: shard by index range, all-gather the 64-byte
partial points, add them (ark_bulletproofs_b200/dist.py -- the code bench.py --gpus N runs with
NCCL). The per-rank partial MSM is computed by the CPU oracle here, so the test checks the host
logic (sharding, gather, point addition incl. identity partials), not the kernels."""
import os
import random
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist

    import bp_oracle as O
    import c_oracle
    from ark_bulletproofs_b200 import codec
    from ark_bulletproofs_b200.dist import allgather_sum_points, shard_range
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    cv = O.SECQ256K1
    rnd = random.Random(11)
    pts, P = [], None
    for _ in range(n):
        P = O.pt_add(cv, P, cv.G)
        pts.append(P)
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    for case in range(2):
        if case == 1:                      # rank 1's shard sums to the identity
            lo1, hi1 = shard_range(n, 1, world)
            for i in range(lo1, hi1):
                sc[i] = 0
        lo, hi = shard_range(n, rank, world)
        part = c_oracle.msm_bytes(0, codec.enc_points(pts[lo:hi], "secq256k1"), codec.enc_scalars(sc[lo:hi], "secq256k1"), hi - lo, 1)
        ident = codec.dec_point(part, "secq256k1") is None
        raw, idn = allgather_sum_points("secq256k1", part, ident)
        got = None if idn else codec.dec_point(raw, "secq256k1")
        want = O.msm(cv, pts, sc)
        q.put((rank, case, got == want))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 64])
def test_sharded_msm_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(4)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, _, ok in res), res


def test_shard_range_covers():
    from ark_bulletproofs_b200.dist import shard_range
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
