"""Quick MSM timing sweep on device-resident synthetic inputs (not the bench contract)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_bulletproofs_b200 import Context  # noqa: E402

ctx = Context("secq256k1", 0)
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 12, int(sys.argv[2]) if len(sys.argv) > 2 else 22
nmax = 1 << hi
pts = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
ctx.synth_points_device(pts.data_ptr(), nmax, 0)
ctx.sync()
g = torch.Generator(device="cuda").manual_seed(2)
sc = torch.randint(0, 256, (nmax * 32,), dtype=torch.uint8, device="cuda", generator=g)
sc.view(-1, 32)[:, 31] &= 0x7F     # < 2^255 < r: valid Montgomery residues
torch.cuda.synchronize()
res = []
for lg in range(lo, hi + 1):
    n = 1 << lg
    for cw in ([0] if len(sys.argv) <= 3 else [int(x) for x in sys.argv[3].split(",")]):
        ctx.set_window(cw)
        ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
            ts.append(time.perf_counter() - t0)
        t = min(ts)
        res.append({"lg_n": lg, "c": cw, "ms": round(t * 1e3, 3), "mpts_s": round(n / t / 1e6, 2)})
        print(res[-1], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "msm_quick.json"), "w"))
