"""Parity at the shapes BASELINE.json is quoted on: 256x256 images, the 2.7 GB concat
buffers' layout at full width, batch 256 - each compared with the CPU oracle on the same injected tensors.

  C3  EDM UNet 256^2 with self-attention at the 32^2 level + middle, forward + backward, B = 2
  C2  EDM UNet 64^2, B = 256: four images of the batch against the oracle run on those four (GroupNorm,
      attention and every other op are per sample, so the other 252 images cannot influence them; dF is
      non-zero only on the four, so the parameter gradients are theirs alone)
  C4  text-conditional UNet 256^2, classifier-free guidance g = 3: one denoise evaluation (model batch 2B)
  C5  Heun 3 steps (5 UNet evaluations) at 256^2
Tolerances as in tests/test_unet_gpu.py (bf16 tensor cores vs the fp32 oracle)."""
import numpy as np
import pytest
import torch

from flaxdiff_b200 import utils
from flaxdiff_b200.inputs import ConditionalInputConfig, DiffusionInputConfig, RandomEmbeddingEncoder
from flaxdiff_b200.models.simple_unet import Unet
from flaxdiff_b200.predictors import KarrasPredictionTransform
from flaxdiff_b200.samplers import EulerAncestralSampler, HeunSampler
from flaxdiff_b200.schedulers import KarrasVENoiseScheduler
from oracle import diffusion_ref as R
from oracle import train_ref, unet_ref

pytestmark = pytest.mark.gpu
dev = torch.device("cuda")


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def cpu_params(fp, grad=False):
    return {k: v.detach().cpu().clone().requires_grad_(grad) for k, v in fp.named.items()}


def perturb(fp, seed=1):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for name, t in fp.named.items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "bias":
            t.copy_(0.1 * torch.randn(t.shape, generator=g, device=dev))
        elif leaf == "scale":
            t.copy_(1 + 0.1 * torch.randn(t.shape, generator=g, device=dev))
    fp.touch()


def check_grads(grads, P, tol_each=6e-2, tol_all=3e-2):
    num = den = 0.0
    for k in P:
        g, gr = grads.named[k].cpu(), P[k].grad
        assert rel(g, gr) < tol_each, (k, rel(g, gr))
        num += (g - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
    assert (num / den) ** 0.5 < tol_all, (num / den) ** 0.5


def test_c3_256px_self_attention_forward_backward_vs_oracle():
    torch.manual_seed(0)
    acfg = (None, None, None, {"heads": 8})
    model = Unet(attention_configs=acfg, dtype=torch.bfloat16)
    fp = model.init(4, device=dev)
    perturb(fp)
    B, res = 2, 256
    x = torch.randn(B, res, res, 3, device=dev).bfloat16()
    t = torch.randn(B, device=dev)
    F, saved = model.forward(fp, x, t, None, save=True)
    P = cpu_params(fp, True)
    Fr = unet_ref.unet_forward(P, x.float().cpu(), t.cpu(), model._fourier_freqs(dev).cpu(), attention_configs=acfg)
    assert F.shape == (B, res, res, 3) and rel(F, Fr) < 3e-2, rel(F, Fr)
    dF = torch.randn(B, res, res, 3, device=dev) / (B * res * res * 3)
    grads = fp.zeros_like()
    model.backward(fp, saved, dF, grads)
    (Fr * dF.cpu()).sum().backward()
    check_grads(grads, P)


def test_c2_batch_256_four_images_vs_oracle():
    torch.manual_seed(0)
    acfg = (None,) * 4
    model = Unet(attention_configs=acfg, dtype=torch.bfloat16)
    fp = model.init(4, device=dev)
    perturb(fp)
    B, res = 256, 64
    pick = [0, 77, 128, 255]
    x = torch.randn(B, res, res, 3, device=dev).bfloat16()
    t = torch.randn(B, device=dev)
    F, saved = model.forward(fp, x, t, None, save=True)
    P = cpu_params(fp, True)
    Fr = unet_ref.unet_forward(P, x[pick].float().cpu(), t[pick].cpu(), model._fourier_freqs(dev).cpu(),
                               attention_configs=acfg)
    assert rel(F[pick], Fr) < 3e-2, rel(F[pick], Fr)
    for i, j in enumerate(pick):                       # every picked image on its own, not just in aggregate
        assert rel(F[j], Fr[i]) < 3e-2
    dF = torch.zeros(B, res, res, 3, device=dev)
    dF[pick] = torch.randn(len(pick), res, res, 3, device=dev) / (len(pick) * res * res * 3)
    grads = fp.zeros_like()
    model.backward(fp, saved, dF, grads)
    (Fr * dF[pick].cpu()).sum().backward()
    check_grads(grads, P)


def test_c4_text_cfg_denoise_eval_256px_vs_oracle():
    torch.manual_seed(0)
    levels = (None, {"heads": 8}, {"heads": 8}, {"heads": 8})
    model = Unet(attention_configs=levels, dtype=torch.bfloat16, context_dim=768)
    fp = model.init(4, device=dev)
    enc = RandomEmbeddingEncoder(77, 768, device=dev)
    cfg = DiffusionInputConfig("image", (256, 256, 3), [ConditionalInputConfig(enc)])
    sched = KarrasVENoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5).to(dev)
    g = 3.0
    smp = EulerAncestralSampler(model, sched, KarrasPredictionTransform(0.5), cfg, guidance_scale=g)
    B, res = 1, 256
    x = torch.randn(B, res, res, 3, device=dev) * 5.0
    tcur = torch.full((B,), 0.4, device=dev)
    cond = enc(["a photo of a cat"]).to(dev)
    x0, eps, Fm = smp.sample_model(fp, x, tcur, cond)
    P, freqs = cpu_params(fp), model._fourier_freqs(dev).cpu()
    null = cfg.get_unconditionals()[0].cpu().expand(B, -1, -1)
    sig = R.karras_sigma(np.full(B, 0.4, np.float32))
    c_in, c_out, c_skip = (torch.from_numpy(a).view(-1, 1, 1, 1) for a in R.karras_coeffs(sig))
    tm = torch.from_numpy(R.karras_model_time(sig))
    xc = x.cpu()
    with torch.no_grad():
        Fc = unet_ref.unet_forward(P, xc * c_in, tm, freqs, attention_configs=levels, textcontext=cond.float().cpu())
        Fu = unet_ref.unet_forward(P, xc * c_in, tm, freqs, attention_configs=levels, textcontext=null.float())
    Fr = Fu + g * (Fc - Fu)
    x0r = c_out * Fr + c_skip * xc
    assert rel(Fm, Fr) < 4e-2, rel(Fm, Fr)
    assert rel(x0, x0r) < 3e-2
    assert rel(eps, (xc - x0r) / torch.from_numpy(sig).view(-1, 1, 1, 1)) < 3e-2


def test_c5_heun_three_steps_256px_vs_oracle(monkeypatch):
    torch.manual_seed(0)
    acfg = (None, None, None, {"heads": 8})
    model = Unet(attention_configs=acfg, dtype=torch.bfloat16)
    fp = model.init(4, device=dev)
    B, res, n = 1, 256, 3
    sched = KarrasVENoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5).to(dev)
    smp = HeunSampler(model, sched, KarrasPredictionTransform(0.5), DiffusionInputConfig("image", (res, res, 3), []))
    prior = torch.randn(B, res, res, 3) * 80.0
    out = smp.generate_samples(fp, B, res, diffusion_steps=n, start_step=1000, priors=prior, device=dev)
    steps = [float(s) for s in smp.get_steps(1000, 0, n)]
    P, freqs = cpu_params(fp), model._fourier_freqs(dev).cpu()
    x = prior.clone()
    one = np.ones(B, np.float32)
    for i in range(n):
        cur, nxt = steps[i] / 1000.0, (steps[i + 1] if i + 1 < n else 0) / 1000.0
        x0, _ = train_ref.karras_denoise_eval(P, x, torch.full((B,), cur), freqs, attention_configs=acfg)
        if i == n - 1:
            x = x0.clamp(-1, 1)
            break
        cs, ns = R.karras_sigma(np.full(B, cur, np.float32)), R.karras_sigma(np.full(B, nxt, np.float32))

        def second(xp, _nxt=nxt):
            return train_ref.karras_denoise_eval(P, torch.from_numpy(xp), torch.full((B,), _nxt), freqs,
                                                 attention_configs=acfg)[0].numpy()
        x = torch.from_numpy(R.heun_step(x.numpy(), x0.numpy(), second, one, cs, one, ns).astype(np.float32))
    assert out.shape == (B, res, res, 3) and rel(out, x) < 5e-2, rel(out, x)
