"""Run a handful of representative kernels once each (for `ncu --set full`); see profiles/."""
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")
B, R = 256, 64


def mk(h, c):
    return torch.randn(B, h, h, c, device=dev).bfloat16()


x64, dy64 = mk(R, 64), mk(R, 64)
w64 = (torch.randn(3, 3, 64, 64, device=dev) / 24).bfloat16()
x320 = mk(R, 320)
w320 = (torch.randn(3, 3, 320, 64, device=dev) / 54).bfloat16()
x256, dy512 = mk(R // 2, 256), mk(R // 2, 512)
w256 = (torch.randn(3, 3, 256, 512, device=dev) / 48).bfloat16()
y64 = torch.empty_like(x64)
y512 = torch.empty_like(dy512)
dw64 = torch.zeros(3, 3, 64, 64, device=dev)
dw256 = torch.zeros(3, 3, 256, 512, device=dev)
dx256 = torch.empty_like(x256)
g, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
for it in range(2):   # first pass warms up, second is what ncu should capture (-s skips the first)
    ops.conv3x3_fwd(x64, w64, out=y64)                 # tc<64,KMN>   N=64 layer
    ops.conv3x3_fwd(x320, w320, out=y64)               # tc<64,KMN>   K=2880
    ops.conv3x3_fwd(x256, w256, out=y512)              # tc<256,KMN>
    ops.conv3x3_dgrad(dy512, w256, dx256)              # tc<256,KK>
    ops.conv3x3_wgrad(x64, dy64, dw64)                 # wgrad9 64->64
    ops.conv3x3_wgrad(x256, dy512, dw256)              # wgrad9 256->512
    st = ops.groupnorm_stats(x64, 8)
    a = ops.groupnorm_apply(x64, 8, st, g, b, 1e-4, True)
    dg, db = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    ops.groupnorm_bwd(x64, dy64, 8, st, g, b, 1e-4, True, dg, db, y64)
    torch.cuda.synchronize()
print("done")
