"""GPU probe: Unet forward/backward vs the CPU oracle (run under gpurun; not a pytest file)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402
from flaxdiff_b200.models.simple_unet import Unet  # noqa: E402
from oracle import unet_ref  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def run(res, B, attn, do_bwd=True):
    dev = torch.device("cuda")
    torch.manual_seed(0)
    acfg = (None, None, None, {"heads": 8}) if attn else (None,) * 4
    model = Unet(attention_configs=acfg, dtype=torch.bfloat16)
    fp = model.init(4, device=dev)
    # make biases / norm params non-trivial so their paths are exercised
    g = torch.Generator(device=dev); g.manual_seed(1)
    for name, t in fp.named.items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "bias":
            t.copy_(0.1 * torch.randn(t.shape, generator=g, device=dev))
        elif leaf == "scale":
            t.copy_(1 + 0.1 * torch.randn(t.shape, generator=g, device=dev))
    x = torch.randn(B, res, res, 3, device=dev).bfloat16()
    t = torch.randn(B, device=dev)
    t0 = time.time()
    Fo, saved = model.forward(fp, x, t, None, save=True)
    torch.cuda.synchronize()
    print(f"[fwd] res={res} B={B} attn={attn} gpu time (incl. first-call setup) {time.time()-t0:.2f}s", flush=True)
    # oracle
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in fp.named.items()}
    freqs = model._fourier_freqs(dev).cpu()
    t1 = time.time()
    Fr = unet_ref.unet_forward(P, x.float().cpu(), t.cpu(), freqs, attention_configs=acfg)
    print(f"[oracle fwd] {time.time()-t1:.2f}s", flush=True)
    e = rel(Fo, Fr)
    print(f"[{'OK ' if e < 3e-2 else 'BAD'}] forward rel err {e:.3e}  |F|={Fr.norm():.3f}", flush=True)
    if not do_bwd:
        return
    dF = torch.randn(B, res, res, 3, device=dev) / (B * res * res * 3)
    grads = fp.zeros_like()
    model.backward(fp, saved, dF, grads)
    torch.cuda.synchronize()
    (Fr * dF.cpu()).sum().backward()
    bad = 0
    worst = []
    num = den = 0.0
    for k in fp.named:
        gr = P[k].grad
        gg = grads.named[k].cpu()
        r = rel(gg, gr)
        num += (gg - gr).pow(2).sum().item(); den += gr.pow(2).sum().item()
        tol = 6e-2
        if not (r < tol):
            bad += 1
        worst.append((r, k, gr.norm().item()))
    worst.sort(reverse=True)
    print(f"[{'OK ' if bad == 0 else 'BAD'}] backward: {len(worst)-bad}/{len(worst)} tensors within 6e-2; global rel {(num/den)**0.5:.3e}", flush=True)
    for r, k, n in worst[:12]:
        print(f"     {r:.3e}  |g|={n:.3e}  {k}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "noattn"):
        run(32, 2, False)
    if which in ("all", "attn"):
        run(64, 2, True)
    if which in ("all", "small"):
        run(16, 3, False)
