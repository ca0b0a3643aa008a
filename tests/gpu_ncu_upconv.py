"""Upsample forward (nearest 2x + conv3x3 as four sub-pixel 2x2 convolutions) at the C2 shape 32x32x128 -> 64x64x256,
B = 256, and the plain 3x3 convolution 128 -> 256 at 64x64, for an `ncu --set full -k regex:fdx_tc_kernel` capture."""
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")
B = 256
x = torch.randn(B, 32, 32, 128, device=dev).bfloat16()
w = torch.randn(3, 3, 128, 256, device=dev) * 0.03
bias = torch.zeros(256, device=dev)
weff = ops.upconv3x3_pack(w)
out = torch.empty(B, 64, 64, 256, device=dev, dtype=torch.bfloat16)
cs = ops.ColStats(B, 256, dev)
for plain in (False, True):
    for _ in range(3):
        ops.upconv3x3_fwd(x, weff, bias, out, colstats=None if plain else (cs, 0))
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.upconv3x3_fwd(x, weff, bias, out, colstats=None if plain else (cs, 0))
    e.record()
    torch.cuda.synchronize()
    print("upconv fwd", "plain" if plain else "colstats", s.elapsed_time(e) / 10 * 1e3, "us")
x2 = torch.randn(B, 64, 64, 128, device=dev).bfloat16()
w16 = w.bfloat16()
for _ in range(3):
    ops.conv3x3_fwd(x2, w16, bias, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    ops.conv3x3_fwd(x2, w16, bias, out=out)
e.record()
torch.cuda.synchronize()
print("conv3x3 128->256 64x64", s.elapsed_time(e) / 10 * 1e3, "us")
