"""Small end-to-end pass over every kernel family for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
MSM (several sizes incl. the warp-partials and chunk levels), device generator generation, prove + verify (one-phase with
generator folding, two-phase k-shuffle with the geometric fold), batch verify, and a 2-context sharded prove."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bp_oracle as O  # noqa: E402  (checker for the small MSM only)
from ark_bulletproofs_b200 import Context, codec  # noqa: E402
from ark_bulletproofs_b200 import r1cs as R  # noqa: E402
from ark_bulletproofs_b200.dist import ThreadGroup  # noqa: E402

curve = "secq256k1"
cv = O.SECQ256K1
ctx = Context(curve, 0)
rnd = random.Random(1)
pts, P = [], None
for _ in range(200):
    P = O.pt_add(cv, P, cv.G)
    pts.append(P)
sc = [rnd.randrange(cv.r) for _ in range(200)]
assert ctx.msm(pts, sc) == O.msm(cv, pts, sc)
import torch  # noqa: E402
for lg in (10, 15):
    n = 1 << lg
    d_pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(d_pts.data_ptr(), n, 0)
    d_sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
    d_sc.view(-1, 32)[:, 31] &= 0x7F
    torch.cuda.synchronize()
    ctx.msm_device(d_pts.data_ptr(), d_sc.data_ptr(), n)
# round 2: the pipeline's own bucket sort (staged and oversize-bin paths, zero-digit tail) and the two-level bucket reduction
for kind in ("uniform", "all_equal", "half_zero"):
    n = 1 << 16
    d_pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(d_pts.data_ptr(), n, 0)
    d_sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
    d_sc[:, 31] &= 0x7F
    if kind == "all_equal":
        d_sc[:] = d_sc[0].clone()
    if kind == "half_zero":
        d_sc[1::2] = 0
    torch.cuda.synchronize()
    ctx.set_sort(1, 1)
    a = ctx.msm_device(d_pts.data_ptr(), d_sc.data_ptr(), n)
    ctx.set_sort(0, 1)
    assert ctx.msm_device(d_pts.data_ptr(), d_sc.data_ptr(), n) == a
ctx.set_sort(1, 1 << 22)
ctx.set_window(17)                          # 2^16 buckets per window: 1024 reduce segments -> msm_reduce_kernel<PAIR>, msm_window_partial_kernel
a = ctx.msm_device(d_pts.data_ptr(), d_sc.data_ptr(), 1 << 12)
ctx.set_two_level_reduce(False)
assert ctx.msm_device(d_pts.data_ptr(), d_sc.data_ptr(), 1 << 12) == a
ctx.set_two_level_reduce(True)
ctx.set_window(0)
r = codec.MODULI[curve][1]


def chain(ctx_, gens, N, nofold):
    ctx_.set_ipa_nofold_threshold(nofold)
    wit = R.ChaChaRng(bytes([3] * 32))
    x0_raw, ks_raw = wit.scalars_raw(curve, 1), wit.scalars_raw(curve, N)
    rng = R.ChaChaRng(bytes(range(32)))
    p = R.Prover(ctx_, gens, R.Transcript(b"ChainCircuit"))
    com, var = p.commit(codec.dec_fe(x0_raw, r), rng.scalar(curve))
    p.chain_circuit_raw(var, N, ks_raw, x0_raw)
    proof = p.prove(rng)
    v = R.Verifier(ctx_, R.Transcript(b"ChainCircuit"))
    v.chain_circuit_raw(v.commit(com), N, ks_raw, None)
    v.verify(proof, gens)
    return proof.to_bytes(), (com, ks_raw)


gens = R.Gens(ctx, 512)                     # device generator path (>= 256)
b1, _ = chain(ctx, gens, 512, 64)           # geometric fold rounds + no-fold tail
b2, _ = chain(ctx, gens, 300, 0)            # padded: general first round, folds to the end
k = 9
inp = [rnd.randrange(1 << 64) for _ in range(k)]
out = list(inp)
rnd.shuffle(out)
t = R.Transcript(b"ShuffleProofTest")
rng = R.ChaChaRng(bytes(range(32)))
p = R.Prover(ctx, gens, t)
coms_raw, vars_ = p.commit_batch_raw(codec.enc_scalars(inp + out, curve), rng.scalars_raw(curve, 2 * k), 2 * k)
p.shuffle_gadget_native(vars_[:k], vars_[k:])
proof = p.prove(rng)
inst = []
for _ in range(3):
    v = R.Verifier(ctx, R.Transcript(b"ShuffleProofTest"))
    vv = v.commit_batch_raw(coms_raw, 2 * k)
    v.shuffle_gadget_native(vv[:k], vv[k:])
    inst.append((v, proof))
R.batch_verify(ctx, R.ChaChaRng(bytes([5] * 32)), inst, gens)


def work(rank, allgather):
    c = Context(curve, 0)
    c.set_collective(rank, 2, allgather)
    g = R.Gens(c, 512)
    return chain(c, g, 512, 64)[0]


res = ThreadGroup(2).run(work)
assert res[0] == res[1] == b1
print("sanitize_small ok; launches:", ctx.launches)
