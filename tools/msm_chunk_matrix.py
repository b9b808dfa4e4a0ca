"""bp_msm (host bases + scalars) and bp_msm_bases (resident bases, host scalars) at 2^lg points for a matrix of
(chunk, first chunk) settings; every result is checked against the device-resident MSM. One process per setting of
BP_MSM_FIRST_CHUNK (the library reads it per call, so os.environ suffices). Usage: python tools/msm_chunk_matrix.py [lg_n]"""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ark_bulletproofs_b200 import Context  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = Context("secq256k1", 0)
n = 1 << lg
pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
ctx.synth_points_device(pts.data_ptr(), n, 0)
ctx.sync()
g = torch.Generator(device="cuda").manual_seed(2)
sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
sc.view(-1, 32)[:, 31] &= 0x7F
torch.cuda.synchronize()
h_pts = torch.empty(n * 64, dtype=torch.uint8, pin_memory=True)
h_sc = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
h_pts.copy_(pts)
h_sc.copy_(sc)
torch.cuda.synchronize()
ref = ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
hb = ctx.bases_upload(h_pts.data_ptr(), n)


def e2e():
    out = ctypes.create_string_buffer(64)
    idn = ctypes.c_int(0)
    ctx._check(ctx.lib.bp_msm(ctx.h, h_pts.data_ptr(), h_sc.data_ptr(), n, out, ctypes.byref(idn)))
    return out.raw, bool(idn.value)


def e2e_bases():
    return ctx.msm_bases(hb, h_sc.data_ptr(), n)


def timeit(fn, reps=5):
    assert fn() == ref, "streamed MSM disagrees with the device-resident MSM"
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return round((time.perf_counter() - t0) / reps * 1e3, 3)


# (chunk, first chunk, growth in percent, cap of a chunk): bp_msm's schedule; bp_msm_bases keeps its own growth (300 %)
# unless growth_bases is given
for cl, fl, gp, capl, gb in [(21, 18, 150, 22, 300), (21, 18, 130, 22, 300), (21, 18, 125, 22, 200), (21, 18, 130, 23, 400),
                             (21, 17, 130, 22, 300), (21, 18, 140, 22, 250), (21, 19, 125, 22, 300), (21, 18, 120, 22, 300)]:
    ctx.set_chunk(1 << cl)
    os.environ["BP_MSM_FIRST_CHUNK"] = str(1 << fl)
    os.environ["BP_MSM_CHUNK_CAP"] = str(1 << capl)
    os.environ["BP_MSM_GROWTH_PCT"] = str(gp)
    a = timeit(e2e)
    os.environ["BP_MSM_GROWTH_PCT"] = str(gb)
    os.environ["BP_MSM_CHUNK_CAP"] = str(1 << 23)
    os.environ["BP_MSM_FIRST_CHUNK"] = str(1 << (fl + 1))
    b = timeit(e2e_bases)
    print(json.dumps({"lg_n": lg, "chunk_lg": cl, "first_chunk_lg": fl, "growth_pct": gp, "cap_lg": capl, "bp_msm_ms": a,
                      "bases_first_lg": fl + 1, "bases_growth_pct": gb, "bp_msm_bases_ms": b}), flush=True)
