"""Batch verification throughput: `count` proofs of the chain circuit with 2^lg multipliers,
distinct witnesses, verified by one batch_verify call (verifier.rs:604-691). With WORLD_SIZE > 1 (torchrun) the proofs
are sharded over the ranks (bp_batch_verify_partial), the partial points all-gathered and summed.
    python tools/batch_verify_quick.py <lg> <count> [distinct]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_bulletproofs_b200 import Context, codec  # noqa: E402
from ark_bulletproofs_b200 import r1cs as R  # noqa: E402
from ark_bulletproofs_b200.dist import allgather_sum_points  # noqa: E402

curve = "secq256k1"
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 12
count = int(sys.argv[2]) if len(sys.argv) > 2 else 64
distinct = int(sys.argv[3]) if len(sys.argv) > 3 else min(count, 8)
nctx = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # contexts (host threads) per GPU: one context's host work overlaps another's kernels
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(curve, local)
N = 1 << lg
r = codec.MODULI[curve][1]
gens = R.Gens(ctx, N)
# `distinct` different proofs (seeds [4;32]+index), repeated to fill the batch: the verifier's work per instance is the same
made = []
for d in range(distinct):
    wit = R.ChaChaRng(bytes([(4 + d) % 256] * 32))
    x0_raw = wit.scalars_raw(curve, 1)
    ks_raw = wit.scalars_raw(curve, N)
    rng = R.ChaChaRng(bytes(range(32)))
    p = R.Prover(ctx, gens, R.Transcript(b"ChainCircuit"))
    com, var = p.commit(codec.dec_fe(x0_raw, r), rng.scalar(curve))
    p.chain_circuit_raw(var, N, ks_raw, x0_raw)
    made.append((p.prove(rng), com, ks_raw))


def instances(idx):
    out = []
    for i in idx:
        proof, com, ks_raw = made[i % distinct]
        v = R.Verifier(ctx, R.Transcript(b"ChainCircuit"))
        vv = v.commit(com)
        v.chain_circuit_raw(vv, N, ks_raw, None)
        out.append((v, proof))
    return out


import threading  # noqa: E402

mine = list(range(rank, count, world))
arng = R.ChaChaRng(bytes([5] * 32))
alphas = [arng.scalar(curve) for _ in range(count)]
ctxs = [ctx] + [Context(curve, local) for _ in range(nctx - 1)]
gens_k = [gens] + [R.Gens(c, N) for c in ctxs[1:]]


def instances_on(c, idx):
    out = []
    for i in idx:
        proof, com, ks_raw = made[i % distinct]
        v = R.Verifier(c, R.Transcript(b"ChainCircuit"))
        vv = v.commit(com)
        v.chain_circuit_raw(vv, N, ks_raw, None)
        out.append((v, proof))
    return out


best = None
for rep in range(3):
    t0 = time.perf_counter()
    shares = [mine[k::nctx] for k in range(nctx)]
    insts = [instances_on(ctxs[k], shares[k]) for k in range(nctx)]
    t_build = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if world == 1 and nctx == 1:
        R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), insts[0], gens)
    else:
        parts = [None] * nctx

        def work(k):
            parts[k] = R.batch_verify_partial(ctxs[k], [alphas[i] for i in shares[k]], insts[k], gens_k[k])
        th = [threading.Thread(target=work, args=(k,)) for k in range(nctx)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        part = ctx.points_sum(parts) if nctx > 1 else parts[0]
        if world > 1:
            raw, idn = allgather_sum_points(curve, codec.enc_point(part, curve) if part is not None else bytes(64), part is None,
                                            device=torch.device("cuda", local))
            assert idn, "batch rejected"
        else:
            assert part is None, "batch rejected"
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
    if best is None or dt < best:
        best = dt
if rank == 0:
    print(json.dumps({"lg_n": lg, "proofs": count, "n_gpus": world, "contexts_per_gpu": nctx, "batch_verify_ms": round(best * 1e3, 2), "proofs_per_s": round(count / best, 1),
                      "ms_per_proof": round(best * 1e3 / count, 3), "verifier_build_ms_per_proof": round(t_build * 1e3 / max(len(mine), 1), 3)}), flush=True)
if world > 1:
    dist.destroy_process_group()
