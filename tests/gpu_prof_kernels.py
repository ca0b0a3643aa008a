"""Run a handful of representative kernels once each (for `ncu --set full`); see profiles/.
    ncu --set full --clock-control none --import-source on -k regex:"fdx_|gn_" -s 11 -c 11 -o gpurun_out/full \
        python tests/gpu_prof_kernels.py
(the first pass of the loop warms up, the second is what -s skips to)."""
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
x128, dy128 = mk(R // 2, 128), mk(R // 2, 128)
w128 = (torch.randn(3, 3, 128, 128, device=dev) / 34).bfloat16()
y64 = torch.empty_like(x64)
y512 = torch.empty_like(dy512)
y128 = torch.empty_like(dy128)
dw64 = torch.zeros(3, 3, 64, 64, device=dev)
dw256 = torch.zeros(3, 3, 256, 512, device=dev)
dx256 = torch.empty_like(x256)
dx320 = torch.empty_like(x320)
g, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
for it in range(2):
    ops.conv3x3_fwd(x64, w64, out=y64)                 # 1  tct<64,1>    64 -> 64 (transposed, interleaved halves)
    ops.conv3x3_fwd(x320, w320, out=y64)               # 2  tct<64,1>    320 -> 64
    ops.conv3x3_fwd(x128, w128, out=y128)              # 3  tct<128,1>   128 -> 128
    ops.conv3x3_fwd(x256, w256, out=y512)              # 4  tc<256,KMN>  256 -> 512 (generic, N = 256 tiles)
    ops.conv3x3_dgrad(dy512, w256, dx256)              # 5  tc<256,KK>
    ops.conv3x3_dgrad(dy64, w320, dx320)               # 6  tct<64,0>    Ncols = 320 (five 64-channel blocks)
    ops.conv3x3_wgrad(x64, dy64, dw64)                 # 7  wgrad9 64 -> 64
    ops.conv3x3_wgrad(x256, dy512, dw256)              # 8  wgrad9 256 -> 512
    st = ops.groupnorm_stats(x64, 8)                   # 9
    a = ops.groupnorm_apply(x64, 8, st, g, b, 1e-4, True)   # 10
    dg, db = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    ops.groupnorm_bwd(x64, dy64, 8, st, g, b, 1e-4, True, dg, db, y64)   # 11-13 (stats, finalize, apply)
    torch.cuda.synchronize()
print("done")
