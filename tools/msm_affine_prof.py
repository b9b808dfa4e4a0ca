"""One MSM at 2^lg with one batched-affine round, for ncu captures of msm_pair_affine_kernel / msm_accumulate_kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_bulletproofs_b200 import Context  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = Context("secq256k1", 0)
n = 1 << lg
pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
ctx.synth_points_device(pts.data_ptr(), n, 0)
g = torch.Generator(device="cuda").manual_seed(2)
sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
sc.view(-1, 32)[:, 31] &= 0x7F
torch.cuda.synchronize()
ctx.set_affine_rounds(rounds, 1 << 16)
for _ in range(reps):
    print(ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)[0][:8].hex())
