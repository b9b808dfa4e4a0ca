"""Phase timings of the device-resident MSM (bp_msm_device) with a closed-form value check:
    python tools/msm_phases_quick.py [lg_n ...]   -> one JSON line per size
The bases are P_i = (i + 1) * G, so the expected sum is ((sum s_i (i + 1)) mod r) * G (tests/test_parity_large_gpu.py)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from ark_bulletproofs_b200 import Context, codec  # noqa: E402
import bp_oracle as O  # noqa: E402

MODES = [int(x) for x in os.environ.get("BP_MODES", "0,1").split(",")]
lgs = [int(x) for x in sys.argv[1:]] or [24]
ctx = Context("secq256k1", 0)
ctx.set_timing(True)
stream = torch.cuda.ExternalStream(ctx.stream_ptr)
nmax = 1 << max(lgs)
pts = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
ctx.synth_points_device(pts.data_ptr(), nmax, 0)
g = torch.Generator(device="cuda").manual_seed(2)
sc = torch.randint(0, 256, (nmax * 32,), dtype=torch.uint8, device="cuda", generator=g)
sc.view(-1, 32)[:, 31] &= 0x7F
torch.cuda.synchronize()
cv = O.SECQ256K1
R256 = 1 << 256
for lg in lgs:
  if True:
    n = 1 << lg
    w = torch.arange(1, n + 1, dtype=torch.int64, device="cuda")
    b = sc[: n * 32].view(n, 32)
    total = 0
    for l in range(32):
        total += int((b[:, l].to(torch.int64) * w).sum().item()) << (8 * l)
    s = total % cv.r * pow(R256, -1, cv.r) % cv.r
    want = O.pt_mul(cv, s, cv.G) if s else None
  for mode in MODES:
    ctx.set_sort(mode, 1 << 18)
    for _ in range(3):
        res = ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        res = ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
    e1.record(stream)
    e1.synchronize()
    raw, ident = res
    got = None if ident else codec.dec_point(raw, "secq256k1")
    ph = ctx.last_phases()
    print(json.dumps({"lg_n": lg, "sort(1=bucket sort,0=cub)": mode, "ms": round(e0.elapsed_time(e1) / reps, 3), "closed_form_ok": got == want,
                      "phases_ms": {k: round(v, 3) for k, v in ph["ms"].items()}}), flush=True)
