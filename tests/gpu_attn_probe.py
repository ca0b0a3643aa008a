"""One attention case in this process (development probe): python tests/gpu_attn_probe.py B h L Lk dh d mode
mode: fwd | bwd.  Prints errors vs the fp32 torch reference."""
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

B, h, L, Lk, dh, d = (int(a) for a in sys.argv[1:7])
mode = sys.argv[7]
dev = torch.device("cuda")
torch.manual_seed(0)
HD = h * dh


def mk(rows):
    t = torch.randn(B, rows, h, dh, device=dev)
    t[..., d:] = 0
    return t.reshape(B, rows, HD).bfloat16()


q, k, v, do = mk(L), mk(Lk), mk(Lk), mk(L)
scale = d ** -0.5
qf, kf, vf = (t.float().view(t.shape[0], t.shape[1], h, dh).requires_grad_(True) for t in (q, k, v))
logits = torch.einsum("blhd,bkhd->bhlk", qf, kf) * scale
o_ref = torch.einsum("bhlk,bkhd->blhd", torch.softmax(logits, -1), vf).reshape(B, L, HD)
lse_ref = torch.logsumexp(logits, -1)
rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()
if mode == "fwd":
    o, lse = ops.attention_fwd(q, k, v, h, dh, scale)
    torch.cuda.synchronize()
    print("fwd ok", rel(o, o_ref), (lse - lse_ref).abs().max().item(), flush=True)
else:
    gq, gk, gv = torch.autograd.grad(o_ref, (qf, kf, vf), do.float())
    o = o_ref.detach().bfloat16()
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse_ref.detach().contiguous(), do, h, dh, scale)
    torch.cuda.synchronize()
    print("bwd ok", rel(dq, gq.reshape(B, L, HD)), rel(dk, gk.reshape(B, Lk, HD)), rel(dv, gv.reshape(B, Lk, HD)), flush=True)
