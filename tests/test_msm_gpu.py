"""GPU parity: bp_msm (C ABI, csrc/msm.cu) vs the CPU oracle on the same inputs.
Covers SURVEY.md section 7 step 4's cases: N = 1,2,3,31,32,33,2^k+-1, zero scalars, repeated
points, identity bases, all-equal scalars, p-1."""
import random

import pytest

import bp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def _ctx():
    from ark_bulletproofs_b200 import Context
    return Context("secq256k1", 0)


@pytest.fixture(params=["tiny", "buckets"])
def ctx(_ctx, request):
    """Every case runs through both small-MSM paths: the single-launch kernel for <= 768 terms per MSM (default) and
    the bucket pipeline that larger inputs always take (bp_msm_set_tiny(0))."""
    _ctx.set_tiny(768 if request.param == "tiny" else 0)
    yield _ctx
    _ctx.set_tiny(768)


def _points(cv, n, rnd):
    base = O.pt_mul(cv, rnd.randrange(1, cv.r), cv.G)
    pts, P = [], None
    for _ in range(n):
        P = O.pt_add(cv, P, base)
        pts.append(P)
    return pts


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 127, 128, 129, 1000, 4097])
def test_msm_random(ctx, n):
    cv = O.SECQ256K1
    rnd = random.Random(n)
    pts = _points(cv, n, rnd)
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    assert ctx.msm(pts, sc) == O.msm(cv, pts, sc)


@pytest.mark.parametrize("c", [3, 4, 7, 8, 11, 13, 16])
def test_msm_windows(ctx, c):
    cv = O.SECQ256K1
    rnd = random.Random(100 + c)
    n = 257
    pts = _points(cv, n, rnd)
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    ctx.set_window(c)
    try:
        assert ctx.msm(pts, sc) == O.msm(cv, pts, sc)
    finally:
        ctx.set_window(0)


def test_msm_edge_scalars(ctx):
    cv = O.SECQ256K1
    rnd = random.Random(5)
    n = 200
    pts = _points(cv, n, rnd)
    for sc in ([0] * n, [1] * n, [cv.r - 1] * n, [rnd.randrange(cv.r)] * n,
               [0 if i % 2 else rnd.randrange(cv.r) for i in range(n)],
               [(1 << 255)] * n, [(1 << 256) - (1 << 32) - 978] * n, [2**128 - 1] * n):
        got = ctx.msm(pts, sc)
        assert got == O.msm(cv, pts, sc)


def test_msm_edge_points(ctx):
    cv = O.SECQ256K1
    rnd = random.Random(6)
    n = 150
    pts = _points(cv, n, rnd)
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    # identity bases (verifier.rs:582-584: A_I2.. may be the identity), repeated points, P and -P
    pts2 = list(pts)
    for i in range(0, n, 7):
        pts2[i] = None
    pts2[1] = pts2[2] = pts2[3]
    pts2[4] = O.pt_neg(cv, pts2[5])
    assert ctx.msm(pts2, sc) == O.msm(cv, pts2, sc)
    # all the same point and all the same scalar: one bucket, every add is a doubling at first
    same = [pts[0]] * n
    assert ctx.msm(same, [3] * n) == O.pt_mul(cv, 3 * n, pts[0])
    # sum that cancels to the identity
    assert ctx.msm([pts[0], O.pt_neg(cv, pts[0])], [5, 5]) is None
    assert ctx.msm([], []) is None


def test_points_sum(ctx):
    cv = O.SECQ256K1
    pts = _points(cv, 9, random.Random(1)) + [None]
    want = None
    for P in pts:
        want = O.pt_add(cv, want, P)
    assert ctx.points_sum(pts) == want


@pytest.mark.parametrize("n", [300, 100])
def test_zorro_msm(n):
    from ark_bulletproofs_b200 import Context
    cv = O.ZORRO
    ctx = Context("zorro", 0)
    rnd = random.Random(77)
    pts = _points(cv, n, rnd)
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    assert ctx.msm(pts, sc) == O.msm(cv, pts, sc)


@pytest.mark.parametrize("tiny", [768, 0])
def test_curve25519_msm(tiny):
    from ark_bulletproofs_b200 import Context
    cv = O.CURVE25519
    ctx = Context("curve25519", 0)
    ctx.set_tiny(tiny)
    rnd = random.Random(78)
    n = 200
    pts = _points(cv, n, rnd)
    pts[5] = None                       # identity base
    pts[7] = pts[8]                     # repeated point (the unified law has no doubling exception)
    pts[9] = O.pt_neg(cv, pts[10])
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    sc[0] = cv.r - 1
    sc[1] = 0
    assert ctx.msm(pts, sc) == O.msm(cv, pts, sc)
    assert ctx.msm([pts[0], O.pt_neg(cv, pts[0])], [5, 5]) is None
    assert ctx.msm([pts[0]] * 40, [3] * 40) == O.pt_mul(cv, 120, pts[0])


def test_msm_host_chunked_overlap(ctx):
    """bp_msm over host buffers streams large inputs in chunks whose H2D copies overlap the previous chunk's kernels;
    all chunks add into one bucket array (msm_run_streamed). Forced here with a tiny chunk size."""
    cv = O.SECQ256K1
    rnd = random.Random(4097)
    n = 4097
    pts = _points(cv, n, rnd)
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    want = O.msm(cv, pts, sc)
    for chunk in (1000, 512, 4096):
        ctx.set_chunk(chunk)
        try:
            assert ctx.msm(pts, sc) == want
        finally:
            ctx.set_chunk(1 << 22)
    # a chunk whose partial sum is the identity
    ctx.set_chunk(100)
    try:
        z = [0] * 100 + sc[100:300]
        assert ctx.msm(pts[:300], z) == O.msm(cv, pts[:300], z)
    finally:
        ctx.set_chunk(1 << 22)


def test_msm_streamed_edge_cases(_ctx):
    """The streamed MSM keeps bucket sums across chunks (interior runs start from the bucket, runs cut by a thread's
    chunk edge move the bucket into their slot, slot sums are added to the bucket): SURVEY 8(d)'s adversarial scalar
    sets -- all-equal scalars (every thread a single run), half zeros, p-1, one repeated point (P + P inside a
    bucket), P and -P (bucket sums passing through the identity) -- against the oracle and the one-shot path."""
    cv = O.SECQ256K1
    rnd = random.Random(77)
    n = 1500
    pts = _points(cv, n, rnd)
    same = rnd.randrange(cv.r)
    cases = [
        (pts, [same] * n),
        (pts, [0 if i % 2 else rnd.randrange(cv.r) for i in range(n)]),
        (pts, [cv.r - 1] * n),
        ([pts[3]] * n, [rnd.randrange(1 << 20) for _ in range(n)]),
        ([pts[i // 2] if i % 2 == 0 else O.pt_neg(cv, pts[i // 2]) for i in range(n)], [same] * n),
        (pts, [rnd.randrange(cv.r) for _ in range(n)]),
    ]
    for bases, sc in cases:
        want = O.msm(cv, bases, sc)
        assert _ctx.msm(bases, sc) == want
        for chunk in (64, 700, 1024):
            _ctx.set_chunk(chunk)
            try:
                assert _ctx.msm(bases, sc) == want
            finally:
                _ctx.set_chunk(1 << 22)


@pytest.mark.parametrize("curve", ["zorro", "curve25519"])
def test_msm_streamed_other_curves(curve):
    from ark_bulletproofs_b200 import Context
    cv = O.CURVES[curve]
    c = Context(curve, 0)
    rnd = random.Random(78)
    n = 600
    pts = _points(cv, n, rnd)
    sc = [rnd.randrange(cv.r) for _ in range(n)]
    want = O.msm(cv, pts, sc)
    c.set_chunk(128)
    assert c.msm(pts, sc) == want


@pytest.mark.parametrize("rounds", [1, 2, 3, 5])
def test_msm_affine_pair_rounds_edge_cases(_ctx, rounds):
    """Batched-affine pair rounds (msm_pair_affine_kernel) forced on small inputs: every exceptional pair is left to the
    XYZZ path -- a repeated point (P + P), P and -P (in one bucket run, all-equal scalars), identity bases, zero scalars --
    and the result equals the oracle's."""
    cv = O.SECQ256K1
    rnd = random.Random(900 + rounds)
    n = 1500
    pts = _points(cv, n, rnd)
    same = rnd.randrange(cv.r)
    pm = [pts[i // 2] if i % 2 == 0 else O.pt_neg(cv, pts[i // 2]) for i in range(n)]
    withid = list(pts)
    for i in range(0, n, 5):
        withid[i] = None
    cases = [
        (pts, [rnd.randrange(cv.r) for _ in range(n)]),
        (pts, [same] * n),                                   # one bucket per window: long runs, sums collide with later bases
        ([pts[3]] * n, [same] * n),                          # all pairs are P + P
        (pm, [same] * n),                                    # neighbours are P and -P
        (withid, [same] * n),
        (pts, [0 if i % 2 else same for i in range(n)]),
        (pts, [cv.r - 1] * n),
        (pts[:7], [same] * 7),
    ]
    _ctx.set_tiny(0)
    _ctx.set_affine_rounds(rounds, 1)
    try:
        for bases, sc in cases:
            assert _ctx.msm(bases, sc) == O.msm(cv, bases, sc)
            _ctx.set_window(5)
            try:
                assert _ctx.msm(bases, sc) == O.msm(cv, bases, sc)
            finally:
                _ctx.set_window(0)
    finally:
        _ctx.set_affine_rounds(0, 1 << 22)
        _ctx.set_tiny(768)


def test_msm_affine_pair_rounds_zorro():
    from ark_bulletproofs_b200 import Context
    cv = O.ZORRO
    c = Context("zorro", 0)
    rnd = random.Random(31)
    n = 700
    pts = _points(cv, n, rnd)
    c.set_tiny(0)
    c.set_affine_rounds(2, 1)
    for sc in ([rnd.randrange(cv.r) for _ in range(n)], [12345] * n):
        assert c.msm(pts, sc) == O.msm(cv, pts, sc)


@pytest.mark.parametrize("curve", ["zorro", "curve25519"])
def test_msm_bucket_sort_other_curves(curve):
    """csrc/msm_sort.cuh on the other two curves (their scalar fields are shorter than 256 bits, so the last window's
    coarse bins see fewer digit values than the planner allows for): same point from both sorts, one-shot and streamed,
    and a 2^12-point prefix against the oracle."""
    import torch
    from ark_bulletproofs_b200 import Context, codec
    cv = {"zorro": O.ZORRO, "curve25519": O.CURVE25519}[curve]
    c = Context(curve, 0)
    n = 1 << 16
    pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    c.synth_points_device(pts.data_ptr(), n, 3)
    c.sync()
    rnd = random.Random(11)
    scal = [rnd.randrange(cv.r) for _ in range(n)]
    raw = b"".join(codec.enc_fe(s, cv.r) for s in scal)
    sc = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    try:
        c.set_sort(1, 1)
        a = c.msm_device(pts.data_ptr(), sc.data_ptr(), n)
        small = c.msm_device(pts.data_ptr(), sc.data_ptr(), 1 << 15)
        c.set_sort(0, 1)
        assert c.msm_device(pts.data_ptr(), sc.data_ptr(), n) == a
        assert c.msm_device(pts.data_ptr(), sc.data_ptr(), 1 << 15) == small
    finally:
        c.set_sort(1, 1 << 22)
    # closed form on the synthetic bases P_i = (3 + i + 1) * G
    tot = sum(s * (4 + i) for i, s in enumerate(scal)) % cv.r
    want = O.pt_mul(cv, tot, cv.G) if tot else None
    got = None if a[1] else codec.dec_point(a[0], curve)
    assert got == want
