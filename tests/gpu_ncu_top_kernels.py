"""One launch each of the round-2 kernels that carry most of the C2 step, for `ncu --set full`:
transposed convolution forward 64->64 at 64x64 (halo stage, eight epilogue warps, residual + column statistics),
its data gradient with the fused GroupNorm-backward first pass, and the nine-tap weight gradient (64->64 at 64x64,
256->512 at 32x32).  B = 256."""
import math
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")
B = 256
x = torch.randn(B, 64, 64, 64, device=dev).bfloat16()
w = (torch.randn(3, 3, 64, 64, device=dev) / math.sqrt(576)).bfloat16()
bias = torch.zeros(64, device=dev)
res = torch.randn(B, 64, 64, 64, device=dev).bfloat16()
out = torch.empty_like(x)
cs = ops.ColStats(B, 64, dev)
gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
st = ops.groupnorm_stats(x, 8)
dg, db = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
dx = torch.empty_like(x)
dw = torch.zeros(3, 3, 64, 64, device=dev)
x2 = torch.randn(B, 32, 32, 256, device=dev).bfloat16()
dy2 = torch.randn(B, 32, 32, 512, device=dev).bfloat16()
dw2 = torch.zeros(3, 3, 256, 512, device=dev)
for _ in range(2):
    ops.conv3x3_fwd(x, w, bias, res=res, out=out, colstats=(cs, 0))
    ops.conv_dgrad_groupnorm_bwd(out, w, x, 8, st, gamma, beta, 1e-4, dg, db, dx, fused=True)
    ops.conv3x3_wgrad(x, out, dw)
    ops.conv3x3_wgrad(x2, dy2, dw2)
torch.cuda.synchronize()
print("ok")
