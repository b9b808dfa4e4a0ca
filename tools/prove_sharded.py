"""Sharded prove/verify timing of the synthetic chain circuit across the GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/prove_sharded.py 16,20
One process per GPU; generators sharded cyclically (bp_ctx_set_collective), partial points all-gathered over NCCL.
Every rank must produce the same proof bytes; rank 0 prints one JSON line per size."""
import hashlib
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
from ark_bulletproofs_b200.dist import torch_allgather  # noqa: E402

curve = "secq256k1"
lgs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16]
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(curve, local)
ctx.set_timing(True)
mode = os.environ.get("BP_COLL", "nccl")      # nccl: library-owned ncclAllGather (bp_ctx_init_nccl); callback: torch.distributed through bp_ctx_set_collective
if world > 1:
    if mode == "nccl":
        ctx.init_nccl(rank, world)
    else:
        ctx.set_collective(rank, world, torch_allgather(device=torch.device("cuda", local)))
_gold_path = os.path.join(ROOT, "tests", "golden", "large.json")
GOLD = json.load(open(_gold_path)) if os.path.exists(_gold_path) else {}
r = codec.MODULI[curve][1]
for lg in lgs:
    N = 1 << lg
    t0 = time.perf_counter()
    gens = R.Gens(ctx, N)
    t_gens = time.perf_counter() - t0
    wit = R.ChaChaRng(bytes([3] * 32))
    x0_raw = wit.scalars_raw(curve, 1)
    ks_raw = wit.scalars_raw(curve, N)
    best = None
    for rep in range(3):
        rng = R.ChaChaRng(bytes(range(32)))
        p = R.Prover(ctx, gens, R.Transcript(b"ChainCircuit"))
        com, var = p.commit(codec.dec_fe(x0_raw, r), rng.scalar(curve))
        p.chain_circuit_raw(var, N, ks_raw, x0_raw)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        proof = p.prove(rng)
        t_prove = (time.perf_counter() - t0) * 1e3
        st_p = ctx.last_stage_ms()
        v = R.Verifier(ctx, R.Transcript(b"ChainCircuit"))
        vv = v.commit(com)
        v.chain_circuit_raw(vv, N, ks_raw, None)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        v.verify(proof, gens)
        t_verify = (time.perf_counter() - t0) * 1e3
        st_v = ctx.last_stage_ms()
        if best is None or t_prove < best[0]:
            best = (t_prove, t_verify, st_p, st_v)
    digest = hashlib.sha256(proof.to_bytes()).hexdigest()
    t = torch.tensor([best[0], best[1]], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        digs = [None] * world
        dist.all_gather_object(digs, digest)
        assert len(set(digs)) == 1, "ranks disagree on the proof bytes"
    gold = GOLD.get("chain_2p%d" % lg)
    if gold is not None:
        assert digest == gold["sha256"], "proof differs from the oracle's golden bytes (tests/golden/large.json)"
    if rank == 0:
        print(json.dumps({"lg_n": lg, "n_gpus": world, "collective": mode if world > 1 else None, "golden": gold is not None, "prove_ms": round(float(t[0]), 2), "verify_ms": round(float(t[1]), 2), "gens_s": round(t_gens, 2),
                          "proof_sha256": digest, "prove_stages": {k: v_ for k, v_ in best[2].items() if v_},
                          "verify_stages": {k: v_ for k, v_ in best[3].items() if v_}}), flush=True)
    del gens
if world > 1:
    dist.destroy_process_group()
