"""Target of the compute-sanitizer run (tests/gpu_probe_batch10.sh): smoke(), then eager training steps of the
BASELINE architectures at reduced batch (C2: 64x64 with 8x8 self-attention, B=16; C3: 256x256 with self-attention,
B=1; C4: text cross-attention + the full transformer block at 64x64, B=2), a lamb + clip + DynamicScale step and
three Euler / Heun denoise steps.  Every kernel family of the default path launches at least once."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
from flaxdiff_b200.inputs import DiffusionInputConfig  # noqa: E402
from flaxdiff_b200.models.simple_unet import Unet  # noqa: E402
from flaxdiff_b200.predictors import KarrasPredictionTransform  # noqa: E402
from flaxdiff_b200.samplers import EulerSampler, HeunSampler  # noqa: E402
from flaxdiff_b200.schedulers import EDMNoiseScheduler, KarrasVENoiseScheduler  # noqa: E402
from flaxdiff_b200.trainer import GeneralDiffusionTrainer, adamw, chain, clip_by_global_norm, lamb  # noqa: E402

g.smoke()
print("smoke ok", flush=True)
dev = torch.device("cuda", 0)


def train(res, B, acfg, opt, steps=2, **kw):
    model = Unet(attention_configs=acfg, dtype=torch.bfloat16, **kw)
    tr = GeneralDiffusionTrainer(
        model, opt, EDMNoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5),
        DiffusionInputConfig("image", (res, res, 3), []), rngs=4, name="san",
        model_output_transform=KarrasPredictionTransform(sigma_data=0.5), ema_decay=0.999, device=dev,
        use_cuda_graph=False)
    step = tr._define_train_step(B)
    batch = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8, device=dev)
    for _ in range(steps):
        tr.state, loss, tr.rngstate = step(tr.state, tr.rngstate, {"image": batch}, 0)
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all(), loss
    return model, tr


m2, t2 = train(64, 16, (None, None, None, {"heads": 8}), adamw(2.7e-4))
print("c2 ok", flush=True)
train(64, 4, (None, None, None, {"heads": 8}), chain(clip_by_global_norm(1.0), lamb(1e-3)))
print("lamb + clip ok", flush=True)
train(256, 1, (None, None, None, {"heads": 8}), adamw(2.7e-4), steps=1)
print("c3 ok", flush=True)
train(64, 2, (None, None, {"heads": 8, "only_pure_attention": False, "use_projection": True}, {"heads": 8}),
      adamw(1e-4), steps=1)
print("full transformer block ok", flush=True)
icfg = DiffusionInputConfig("image", (64, 64, 3), [])
sched = KarrasVENoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5)
for cls in (EulerSampler, HeunSampler):
    smp = cls(m2, sched, KarrasPredictionTransform(sigma_data=0.5), icfg)
    out = smp.generate_samples(t2.state.ema_params, num_samples=4, resolution=64, diffusion_steps=3, start_step=1000)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
print("samplers ok", flush=True)
