"""GPU parity at the sizes bench.py reports ("parity stops three orders of magnitude below the benchmark").

The synthetic bases of the bench are P_i = (start + i + 1) * G (bp_synth_points_device), so the MSM has a closed form
that the oracle evaluates with one big-integer sum and one scalar multiplication at any n:

    sum_i s_i * P_i = ((sum_i s_i * (start + i + 1)) mod r) * G

Checked here for the *automatic* plan (c = 20, 13 windows, L = 64, 64-bucket reduce segments at 2^24 -- the plan the
headline number runs) through both entry points, the device-resident bp_msm_device and the host-buffer bp_msm (streamed
in chunks above 6 M points), with SURVEY 8(d)'s adversarial scalar sets: uniform, all equal (one bucket per window: partial
sums k*G meet the next base (k+1)*G... and the doubling branch of the mixed addition), half zeros, all r - 1.
At n <= 2^20 the same inputs also go through the C restatement of ark's Pippenger (oracle/c/bp_ref.c).
"""
import ctypes

import pytest

import bp_oracle as O

pytestmark = pytest.mark.gpu

CURVE = "secq256k1"
R256 = 1 << 256


@pytest.fixture(scope="module")
def ctx():
    from ark_bulletproofs_b200 import Context
    return Context(CURVE, 0)


def _scalar_bytes(kind, n, seed):
    """n raw 32-byte Montgomery residues (the C ABI's scalar format) as a uint8 cuda tensor."""
    import torch
    r = O.SECQ256K1.r
    g = torch.Generator(device="cuda").manual_seed(seed)
    if kind in ("uniform", "half_zero"):
        sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
        sc[:, 31] &= 0x7F                      # < 2^255 < r
        if kind == "half_zero":
            sc[1::2, :] = 0
        return sc.reshape(-1)
    if kind == "all_equal":
        v = int.from_bytes(bytes(torch.randint(0, 256, (32,), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed)).tolist()), "little") % r
    elif kind == "r_minus_1":
        v = (r - 1) * R256 % r                 # Montgomery form of r - 1
    else:
        raise ValueError(kind)
    row = torch.tensor(list(v.to_bytes(32, "little")), dtype=torch.uint8, device="cuda")
    return row.repeat(n)


def _closed_form(sc, n, start):
    """((sum_i s_i (start + i + 1)) mod r) * G for raw Montgomery residues sc (uint8 cuda tensor, n x 32)."""
    import torch
    cv = O.SECQ256K1
    w = torch.arange(start + 1, start + n + 1, dtype=torch.int64, device="cuda")     # < 2^28
    b = sc.view(n, 32)
    total = 0
    for l in range(32):                         # byte column l: sum < 2^8 * 2^28 * 2^24 = 2^60
        total += int((b[:, l].to(torch.int64) * w).sum().item()) << (8 * l)
    s = total % cv.r * pow(R256, -1, cv.r) % cv.r
    return O.pt_mul(cv, s, cv.G) if s else None


def _dec(raw, ident):
    from ark_bulletproofs_b200 import codec
    return None if ident else codec.dec_point(raw, CURVE)


@pytest.mark.parametrize("lg_n", [12, 16, 20, 24])
@pytest.mark.parametrize("kind", ["uniform", "all_equal", "half_zero", "r_minus_1"])
def test_msm_closed_form(ctx, lg_n, kind):
    import torch
    n, start = 1 << lg_n, 3 * (1 << lg_n)
    pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(pts.data_ptr(), n, start)
    ctx.sync()
    sc = _scalar_bytes(kind, n, 1000 + lg_n)
    torch.cuda.synchronize()
    want = _closed_form(sc, n, start)
    # device-resident entry point, automatic plan
    got = _dec(*ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n))
    assert got == want, "bp_msm_device differs from the closed form (n = 2^%d, %s)" % (lg_n, kind)
    # host-buffer entry point (streamed above the chunk threshold)
    h_pts = torch.empty(n * 64, dtype=torch.uint8, pin_memory=True)
    h_sc = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
    h_pts.copy_(pts)
    h_sc.copy_(sc)
    torch.cuda.synchronize()
    out = ctypes.create_string_buffer(64)
    idn = ctypes.c_int(0)
    ctx._check(ctx.lib.bp_msm(ctx.h, h_pts.data_ptr(), h_sc.data_ptr(), n, out, ctypes.byref(idn)))
    assert _dec(out.raw, bool(idn.value)) == want, "bp_msm (host buffers) differs from the closed form (n = 2^%d, %s)" % (lg_n, kind)
    # bases resident on the GPU (bp_bases_upload), scalars streamed from the host
    hb = ctx.bases_upload(h_pts.data_ptr(), n)
    try:
        assert _dec(*ctx.msm_bases(hb, h_sc.data_ptr(), n)) == want, "bp_msm_bases differs from the closed form (n = 2^%d, %s)" % (lg_n, kind)
        if kind == "uniform":
            # a window of the resident table: points [off, off + m) with the first m scalars
            off, m = n // 4 + 3, n // 2
            w2 = _closed_form(sc[:m * 32], m, start + off)
            assert _dec(*ctx.msm_bases(hb, h_sc.data_ptr(), m, offset=off)) == w2
            with pytest.raises(Exception):
                ctx.msm_bases(hb, h_sc.data_ptr(), n, offset=1)          # out of range: BP_ERR_LEN
    finally:
        ctx.bases_free(hb)
    if lg_n == 20 and kind in ("uniform", "all_equal"):
        # forced small chunks: many chunks adding into one bucket array
        ctx.set_chunk(1 << 17)
        try:
            ctx._check(ctx.lib.bp_msm(ctx.h, h_pts.data_ptr(), h_sc.data_ptr(), n, out, ctypes.byref(idn)))
        finally:
            ctx.set_chunk(1 << 22)
        assert _dec(out.raw, bool(idn.value)) == want
    if lg_n <= 20 and kind in ("uniform", "half_zero"):
        # the same bytes through the C restatement of ark-ec's msm_bigint_wnaf
        import c_oracle
        ref = c_oracle.msm_ptr(0, h_pts.data_ptr(), h_sc.data_ptr(), n, c_oracle.num_threads())
        assert _dec(ref, ref == bytes(64)) == want, "oracle/c/bp_ref.c differs from the closed form"


@pytest.mark.parametrize("lg_n,rounds", [(20, 1), (20, 3), (24, 2), (24, 4)])
@pytest.mark.parametrize("kind", ["uniform", "all_equal", "half_zero"])
def test_msm_closed_form_affine_rounds(ctx, lg_n, rounds, kind):
    """The batched-affine pair rounds (bp_msm_set_affine_rounds) in front of the XYZZ accumulation, at the bench sizes."""
    import torch
    n, start = 1 << lg_n, 11
    pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(pts.data_ptr(), n, start)
    sc = _scalar_bytes(kind, n, 500 + lg_n + rounds)
    torch.cuda.synchronize()
    want = _closed_form(sc, n, start)
    ctx.set_affine_rounds(rounds, 1 << 16)
    try:
        assert _dec(*ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)) == want
    finally:
        ctx.set_affine_rounds(0, 1 << 22)


@pytest.mark.parametrize("c", [12, 16, 18])
def test_msm_closed_form_forced_windows(ctx, c):
    """Other window widths at 2^18 points (segment sizes 4..32 of the bucket reduction, L < 64)."""
    import torch
    n, start = 1 << 18, 17
    pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(pts.data_ptr(), n, start)
    sc = _scalar_bytes("uniform", n, 77 + c)
    torch.cuda.synchronize()
    want = _closed_form(sc, n, start)
    ctx.set_window(c)
    try:
        assert _dec(*ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)) == want
    finally:
        ctx.set_window(0)


def test_synth_points_are_multiples_of_g(ctx):
    """The closed form rests on bp_synth_points_device: spot-check P_i = (start + i + 1) * G against the oracle."""
    import torch
    from ark_bulletproofs_b200 import codec
    cv = O.SECQ256K1
    n, start = 4096, (1 << 27) + 5
    pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(pts.data_ptr(), n, start)
    ctx.sync()
    raw = bytes(pts.cpu().numpy())
    for i in (0, 1, 2, 31, 32, 1000, 4095):
        assert codec.dec_point(raw[64 * i:64 * i + 64], CURVE) == O.pt_mul(cv, start + i + 1, cv.G)


@pytest.mark.parametrize("curve,cap", [("secq256k1", 256), ("secq256k1", 1000), ("zorro", 256), ("zorro", 700), ("curve25519", 256), ("curve25519", 600)])
def test_device_generators_match_oracle(curve, cap):
    """BulletproofGens chains generated on the GPU (csrc/gens_kernels.cuh; capacity >= 256 takes the device path) against
    the oracle's restatement of src/generators.rs:71-121,196-221 -- not against the repo's own host generator. zorro: the x
    draw rejects about every second time (the host walks the keystream, the device takes the square roots); curve25519:
    point from y, then the cofactor."""
    from ark_bulletproofs_b200 import Context
    from ark_bulletproofs_b200 import r1cs as R
    c = Context(curve, 0)
    g = R.Gens(c, cap)
    bp = O.BulletproofGens(O.CURVES[curve], cap, 1)
    assert g.export(0, 0, cap) == bp.G(cap)
    assert g.export(1, 0, cap) == bp.H(cap)
    c.set_device_gens(False)
    g2 = R.Gens(c, cap)                       # the host generator gives the same tables
    assert g2.export(0, 0, cap) == bp.G(cap) and g2.export(1, 0, cap) == bp.H(cap)


@pytest.mark.parametrize("curve", ["zorro", "curve25519"])
def test_device_generators_large_match_host(curve):
    """2^14 generators per chain: device chain == threaded host chain (which tests/test_host_layer.py pins to the oracle)."""
    from ark_bulletproofs_b200 import Context
    from ark_bulletproofs_b200 import r1cs as R
    cap = 1 << 14
    c = Context(curve, 0)
    g = R.Gens(c, cap)
    _, _, G, H = R.generate_gens_host(curve, cap)
    assert g.export(0, 0, cap) == G and g.export(1, 0, cap) == H


# ---- byte-identical proofs at 2^16 / 2^20 multipliers (tests/golden/large.json, made by make_golden_large.py) ----------
import hashlib
import json
import os

_LARGE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "large.json")
LARGE = json.load(open(_LARGE_PATH)) if os.path.exists(_LARGE_PATH) else {}


def _prove_large(R, ctx, gens, g):
    """The scenario of tests/oracle_cases.py (SEED-A) through the C ABI with the native circuit builders."""
    import oracle_cases as C
    from ark_bulletproofs_b200 import codec
    curve, kind, params = g["curve"], g["kind"], g["params"]
    rng = R.ChaChaRng(bytes(range(32)))
    if kind == "chain":
        N = params["N"]
        wit = R.ChaChaRng(bytes([3] * 32))                 # bp_oracle.chain_circuit_witness
        x0_raw = wit.scalars_raw(curve, 1)
        ks_raw = wit.scalars_raw(curve, N)
        p = R.Prover(ctx, gens, R.Transcript(b"ChainCircuit"))
        com, var = p.commit(codec.dec_fe(x0_raw, codec.MODULI[curve][1]), rng.scalar(curve))
        p.chain_circuit_raw(var, N, ks_raw, x0_raw)
        proof = p.prove(rng)

        def verifier():
            v = R.Verifier(ctx, R.Transcript(b"ChainCircuit"))
            vv = v.commit(com)
            v.chain_circuit_raw(vv, N, ks_raw, None)
            return v
        com_bytes = O.ser_point(O.CURVES[curve], com, True)
        return proof, com_bytes, verifier
    assert kind == "shuffle"
    k = params["k"]
    inp, out = C.shuffle_values(k, params["seed"])
    vals_raw = codec.enc_scalars(inp + out, curve)

    def transcript():
        t = R.Transcript(b"ShuffleProofTest")
        t.append_message(b"dom-sep", b"ShuffleProof")
        t.append_u64(b"k", k)
        return t
    blinds_raw = rng.scalars_raw(curve, 2 * k)             # Fr::rand in commit order, then the same rng goes to prove()
    p = R.Prover(ctx, gens, transcript())
    coms_raw, vars_ = p.commit_batch_raw(vals_raw, blinds_raw, 2 * k)
    p.shuffle_gadget_native(vars_[:k], vars_[k:])
    proof = p.prove(rng)

    def verifier():
        v = R.Verifier(ctx, transcript())
        vv = v.commit_batch_raw(coms_raw, 2 * k)
        v.shuffle_gadget_native(vv[:k], vv[k:])
        return v
    cv = O.CURVES[curve]
    com_bytes = b"".join(O.ser_point(cv, codec.dec_point(coms_raw[64 * i:64 * i + 64], curve), True) for i in range(2 * k))
    return proof, com_bytes, verifier


@pytest.mark.parametrize("name", sorted(LARGE) or ["<no tests/golden/large.json>"])
def test_large_golden_proofs_byte_identical(name):
    if not LARGE:
        pytest.skip("tests/golden/large.json missing")
    from ark_bulletproofs_b200 import Context
    from ark_bulletproofs_b200 import r1cs as R
    g = LARGE[name]
    ctx = Context(g["curve"], 0)
    gens = R.Gens(ctx, g["gens_capacity"])
    proof, com_bytes, verifier = _prove_large(R, ctx, gens, g)
    raw = proof.to_bytes()
    assert hashlib.sha256(com_bytes).hexdigest() == g["commitments_sha256"]
    assert hashlib.sha256(raw).hexdigest() == g["sha256"]
    assert raw.hex() == g["proof_hex"]
    # the oracle's bytes verify on the GPU; a tampered copy does not
    verifier().verify(R.Proof.from_bytes(g["curve"], bytes.fromhex(g["proof_hex"])), gens)
    bad = bytearray(raw)
    bad[-32] ^= 1
    with pytest.raises(Exception):
        verifier().verify(R.Proof.from_bytes(g["curve"], bytes(bad)), gens)


@pytest.mark.parametrize("world", [2, 8])
def test_large_golden_sharded(world):
    """2^16 multipliers with the generators sharded over `world` contexts (threads on one GPU): golden bytes on every rank."""
    if "chain_2p16" not in LARGE:
        pytest.skip("tests/golden/large.json missing")
    from test_r1cs_gpu import _sharded
    g = LARGE["chain_2p16"]

    def run(R, ctx, rank):
        gens = R.Gens(ctx, g["gens_capacity"])
        proof, _, verifier = _prove_large(R, ctx, gens, g)
        raw = proof.to_bytes()
        verifier().verify(R.Proof.from_bytes(g["curve"], raw), gens)
        return hashlib.sha256(raw).hexdigest()
    for digest in _sharded(world, g["curve"], run):
        assert digest == g["sha256"]


@pytest.mark.parametrize("lg_n", [14, 16, 20])
@pytest.mark.parametrize("kind", ["uniform", "all_equal", "half_zero", "r_minus_1"])
def test_msm_bucket_sort_closed_form(ctx, lg_n, kind):
    """csrc/msm_sort.cuh at sizes below its default threshold: the pipeline's own two-pass bucket sort against the closed
    form and against the library sort, for every adversarial scalar set (all_equal / r_minus_1 put a whole window into one
    bucket: the global-memory path of the bins kernel; half_zero: the zero-digit tail)."""
    import torch
    n, start = 1 << lg_n, 5 * (1 << lg_n) + 7
    pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(pts.data_ptr(), n, start)
    ctx.sync()
    sc = _scalar_bytes(kind, n, 4000 + lg_n)
    torch.cuda.synchronize()
    want = _closed_form(sc, n, start)
    try:
        ctx.set_sort(1, 1)
        got1 = ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
        ctx.set_sort(0, 1)
        got0 = ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
    finally:
        ctx.set_sort(1, 1 << 22)
    assert _dec(*got1) == want
    assert got0 == got1


def test_msm_bucket_sort_ragged_sizes(ctx):
    """sizes that are no multiple of the tile or the bin size, through both sorts"""
    import torch
    for n in (32768 + 1, 40000 + 17, (1 << 17) - 3):
        pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        ctx.synth_points_device(pts.data_ptr(), n, 11)
        ctx.sync()
        sc = _scalar_bytes("uniform", n, 77 + n)
        torch.cuda.synchronize()
        want = _closed_form(sc, n, 11)
        try:
            ctx.set_sort(1, 1)
            got1 = ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
        finally:
            ctx.set_sort(1, 1 << 22)
        assert _dec(*got1) == want
