"""One streamed bp_msm over pinned host buffers (for ncu captures of the streamed kernels: every msm_accumulate_kernel
launch of this process is the ACC variant). Usage: python tools/msm_stream_once.py [lg_n] [calls]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ark_bulletproofs_b200 import Context  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = Context("secq256k1", 0)
n = 1 << lg
pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
ctx.synth_points_device(pts.data_ptr(), n, 0)
ctx.sync()
g = torch.Generator(device="cuda").manual_seed(2)
sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
sc.view(-1, 32)[:, 31] &= 0x7F
h_pts = torch.empty(n * 64, dtype=torch.uint8, pin_memory=True)
h_sc = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
h_pts.copy_(pts)
h_sc.copy_(sc)
torch.cuda.synchronize()
for _ in range(calls):
    out = ctypes.create_string_buffer(64)
    idn = ctypes.c_int(0)
    ctx._check(ctx.lib.bp_msm(ctx.h, h_pts.data_ptr(), h_sc.data_ptr(), n, out, ctypes.byref(idn)))
print(out.raw.hex()[:32], bool(idn.value))
