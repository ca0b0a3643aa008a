"""One forward + backward of the fused attention at the C3 self-attention shape (B=64, 8 heads, L=1024, d=64),
for an `ncu --set full -k regex:fdx_attn` capture (tests/gpu_probe_batch14.sh)."""
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")
B, h, L, dh = 64, 8, 1024, 64
q, k, v, do = (torch.randn(B, L, h * dh, device=dev).bfloat16() for _ in range(4))
for _ in range(2):
    o, lse = ops.attention_fwd(q, k, v, h, dh, dh ** -0.5)
    g = ops.attention_bwd(q, k, v, o, lse, do, h, dh, dh ** -0.5)
torch.cuda.synchronize()
print("ok")
