"""A/B of the batched-affine pair rounds in front of the XYZZ bucket accumulation (bp_msm_set_affine_rounds):
    python tools/msm_affine_quick.py [lg_n ...]      -> one JSON line per (size, rounds) with the MSM phases"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_bulletproofs_b200 import Context  # noqa: E402

lgs = [int(x) for x in sys.argv[1:]] or [22, 24]
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
for lg in lgs:
    n = 1 << lg
    ref = None
    for rounds in (0, 1, 2, 3, 4):
        ctx.set_affine_rounds(rounds, 1 << 16)
        for _ in range(2):
            res = ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            res = ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
        e1.record(stream)
        e1.synchronize()
        if ref is None:
            ref = res
        ph = ctx.last_phases()
        print(json.dumps({"lg_n": lg, "affine_rounds": rounds, "ms": round(e0.elapsed_time(e1) / 3, 3), "same_result": res == ref,
                          "phases_ms": {k: round(v, 3) for k, v in ph["ms"].items()}}), flush=True)
