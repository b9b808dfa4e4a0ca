"""bp_msm over pinned host buffers (the e2e leg of bench.py) for several chunk settings; checks every result against
the device-resident MSM of the same inputs. Usage: python tools/msm_e2e_sweep.py [lg_n] [chunk_lg ...]"""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ark_bulletproofs_b200 import Context  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
chunk_lgs = [int(a) for a in sys.argv[2:]] or [20, 21, 22, 23]
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
t0 = time.perf_counter()
for _ in range(3):
    ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)
dev_ms = (time.perf_counter() - t0) / 3 * 1e3


def e2e():
    out = ctypes.create_string_buffer(64)
    idn = ctypes.c_int(0)
    ctx._check(ctx.lib.bp_msm(ctx.h, h_pts.data_ptr(), h_sc.data_ptr(), n, out, ctypes.byref(idn)))
    return out.raw, bool(idn.value)


for cl in chunk_lgs:
    ctx.set_chunk(1 << cl)
    assert e2e() == ref, "streamed MSM disagrees with the device-resident MSM"
    e2e()
    t0 = time.perf_counter()
    for _ in range(5):
        e2e()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"lg_n": lg, "chunk_lg": cl, "first_chunk_lg": cl - 2, "e2e_ms": round(ms, 3), "device_ms": round(dev_ms, 3),
                      "e2e_mpoints_s": round(n / ms / 1e3, 1)}), flush=True)

# How fast is the H2D copy of the whole input alone, and next to a device-resident MSM (the streamed path's situation)?
import threading  # noqa: E402

side = torch.cuda.Stream()
dst_p, dst_s = torch.empty_like(pts), torch.empty_like(sc)


def copy_ms():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record()
        dst_p.copy_(h_pts, non_blocking=True)
        dst_s.copy_(h_sc, non_blocking=True)
        e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)


copy_ms()
alone = copy_ms()
ctx.set_chunk(1 << 22)
th = threading.Thread(target=lambda: [ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n) for _ in range(2)])
th.start()
time.sleep(0.01)
under = copy_ms()
th.join()
print(json.dumps({"lg_n": lg, "h2d_ms_alone": round(alone, 2), "h2d_ms_next_to_msm": round(under, 2),
                  "h2d_gbs_alone": round(n * 96 / alone / 1e6, 1), "h2d_gbs_next_to_msm": round(n * 96 / under / 1e6, 1)}), flush=True)
