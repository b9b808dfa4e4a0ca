"""world_size-2 gloo test (CPU) of the cyclic generator sharding used by the multi-GPU prover (SURVEY.md 8(e),
ark_bulletproofs_b200/dist.py, csrc/r1cs.cuh ipa_create): rank g owns G_i, H_i with i = g (mod P); every round
each rank folds its shard locally and contributes a partial L / R point (with its share of c_L, c_R on Q); the
partials are all-gathered through the same callback the library uses (dist.torch_allgather) and added. The
per-rank arithmetic is the CPU oracle's, so this checks the sharding algebra and the collective plumbing, not
the kernels: the sharded run must reproduce the oracle's InnerProductProof::create exactly."""
import os
import random
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sharded_ipa_rank(O, cv, t, Q, Gf, Hf, G_loc, H_loc, a, b, rank, world, allgather, sum_points):
    """One rank of the sharded IPA: G_loc/H_loc = cyclic shard, a/b/Gf/Hf replicated. Returns (L, R, a, b)."""
    r = cv.r
    n = len(a)
    t.append_message(b"dom-sep", b"ipp v1")
    t.append_u64(b"n", n)
    Ls, Rs = [], []
    # generators stay unfolded once the partners live on different ranks: track the expansion coefficients
    a, b, G_loc, H_loc = list(a), list(b), list(G_loc), list(H_loc)
    fold_G = [Gf[j * world + rank] for j in range(len(G_loc))]       # per-element factor of each local stage generator
    fold_H = [Hf[j * world + rank] for j in range(len(H_loc))]
    stage_n = None                                                    # global length at which folding stopped
    coefG, coefH = None, None
    while n != 1:
        h = n // 2
        hl = h // world
        cL = sum(a[i] * b[h + i] for i in range(rank, h, world)) % r
        cR = sum(a[h + i] * b[i] for i in range(rank, h, world)) % r
        if hl >= 1 and stage_n is None:
            # fold mode: local halves [0,hl) / [hl,2hl); global index of local j is j*world + rank
            gi = [j * world + rank for j in range(hl)]
            Lp = O.msm(cv, G_loc[hl:2 * hl] + H_loc[:hl] + [Q],
                       [a[i] * fold_G[hl + j] % r for j, i in enumerate(gi)] + [b[h + i] * fold_H[j] % r for j, i in enumerate(gi)] + [cL])
            Rp = O.msm(cv, G_loc[:hl] + H_loc[hl:2 * hl] + [Q],
                       [a[h + i] * fold_G[j] % r for j, i in enumerate(gi)] + [b[i] * fold_H[hl + j] % r for j, i in enumerate(gi)] + [cR])
        else:
            if stage_n is None:
                stage_n = n
                coefG, coefH = list(fold_G), list(fold_H)
            # no-fold mode: every local stage generator belongs to exactly one of L / R at the current level
            Lpts, Lsc, Rpts, Rsc = [Q], [cL], [Q], [cR]
            for tl in range(len(G_loc)):
                ip = (tl * world + rank) % n
                if ip >= h:
                    Lpts.append(G_loc[tl]); Lsc.append(a[ip - h] * coefG[tl] % r)
                    Rpts.append(H_loc[tl]); Rsc.append(b[ip - h] * coefH[tl] % r)
                else:
                    Rpts.append(G_loc[tl]); Rsc.append(a[h + ip] * coefG[tl] % r)
                    Lpts.append(H_loc[tl]); Lsc.append(b[h + ip] * coefH[tl] % r)
            Lp, Rp = O.msm(cv, Lpts, Lsc), O.msm(cv, Rpts, Rsc)
        both = allgather(enc(O, cv, Lp) + enc(O, cv, Rp))
        Lp = sum_points([both[k * 128:k * 128 + 64] for k in range(world)])
        Rp = sum_points([both[k * 128 + 64:k * 128 + 128] for k in range(world)])
        Ls.append(Lp); Rs.append(Rp)
        O.append_point(cv, t, b"L", Lp)
        O.append_point(cv, t, b"R", Rp)
        u = O.challenge_scalar(cv, t, b"u")
        ui = pow(u, -1, r)
        a = [(a[i] * u + ui * a[h + i]) % r for i in range(h)]
        b = [(b[i] * ui + u * b[h + i]) % r for i in range(h)]
        if stage_n is None:
            G_loc = [O.pt_add(cv, O.pt_mul(cv, ui * fold_G[j] % r, G_loc[j]), O.pt_mul(cv, u * fold_G[hl + j] % r, G_loc[hl + j])) for j in range(hl)]
            H_loc = [O.pt_add(cv, O.pt_mul(cv, u * fold_H[j] % r, H_loc[j]), O.pt_mul(cv, ui * fold_H[hl + j] % r, H_loc[hl + j])) for j in range(hl)]
            fold_G, fold_H = [1] * hl, [1] * hl
        else:
            for tl in range(len(G_loc)):
                hi = ((tl * world + rank) % n) >= h
                coefG[tl] = coefG[tl] * (u if hi else ui) % r
                coefH[tl] = coefH[tl] * (ui if hi else u) % r
        n = h
    return Ls, Rs, a[0], b[0]


def enc(O, cv, P):
    from ark_bulletproofs_b200 import codec
    return bytes(64) if P is None else codec.enc_point(P, "secq256k1")


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ctypes

    import torch.distributed as dist

    import bp_oracle as O
    from ark_bulletproofs_b200 import _lib, codec
    from ark_bulletproofs_b200.dist import torch_allgather
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    cv = O.SECQ256K1
    rnd = random.Random(5)
    bp = O.BulletproofGens(cv, n, 1)
    G, H = bp.G(n), bp.H(n)
    Q = O.pt_mul(cv, 12345, cv.G)
    a = [rnd.randrange(cv.r) for _ in range(n)]
    b = [rnd.randrange(cv.r) for _ in range(n)]
    Gf = [rnd.randrange(1, cv.r) for _ in range(n)]
    Hf = [rnd.randrange(1, cv.r) for _ in range(n)]
    lib = _lib.load()

    def sum_points(raws):
        out = ctypes.create_string_buffer(64)
        idn = ctypes.c_int(0)
        assert lib.bp_points_sum_curve(codec.CURVE_IDS["secq256k1"], b"".join(raws), len(raws), out, ctypes.byref(idn)) == 0
        return None if idn.value else codec.dec_point(out.raw, "secq256k1")
    got = sharded_ipa_rank(O, cv, O.Transcript(b"shardtest"), Q, Gf, Hf, G[rank::world], H[rank::world], a, b, rank, world,
                           torch_allgather(), sum_points)
    want = O.ipa_create(cv, O.Transcript(b"shardtest"), Q, Gf, Hf, G, H, a, b)
    q.put((rank, got == (want.L_vec, want.R_vec, want.a, want.b)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [2, 8])
def test_sharded_ipa_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_thread_group_allgather():
    from ark_bulletproofs_b200.dist import ThreadGroup
    tg = ThreadGroup(4)
    out = tg.run(lambda r, ag: [ag(bytes([r]) * 3), ag(bytes([10 + r]))])
    assert all(o == [b"\x00\x00\x00\x01\x01\x01\x02\x02\x02\x03\x03\x03", bytes([10, 11, 12, 13])] for o in out)
