"""Per-phase device timings of the MSM (cudaEvents inside the library) at a few sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ark_bulletproofs_b200 import Context  # noqa: E402

ctx = Context("secq256k1", 0)
ctx.set_timing(True)
for lg in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,20,24").split(",")]:
    n = 1 << lg
    pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(pts.data_ptr(), n, 0)
    ctx.sync()
    g = torch.Generator(device="cuda").manual_seed(2)
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
    sc.view(-1, 32)[:, 31] &= 0x7F
    torch.cuda.synchronize()
    import time
    for _ in range(3):
        ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
    wall = (time.perf_counter() - t0) / 5 * 1e3
    ph = ctx.last_phases()
    print(lg, "c=%d W=%d" % (ph["c"], ph["windows"]), {k: round(v, 3) for k, v in ph["ms"].items()}, "gpu_sum=%.3f wall=%.3f ms" % (sum(ph["ms"].values()), wall), flush=True)
    del pts, sc
