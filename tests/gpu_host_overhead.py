"""Host-side cost of one training step (python + launches), measured with per-step synchronisation.
    python tests/gpu_host_overhead.py            (run under gpurun)
"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from flaxdiff_b200.inputs import DiffusionInputConfig  # noqa: E402
from flaxdiff_b200.models.simple_unet import Unet  # noqa: E402
from flaxdiff_b200.predictors import KarrasPredictionTransform  # noqa: E402
from flaxdiff_b200.schedulers import EDMNoiseScheduler  # noqa: E402
from flaxdiff_b200.trainer import GeneralDiffusionTrainer, adamw  # noqa: E402

dev = torch.device("cuda")
B, res = 256, 64
model = Unet(attention_configs=(None, None, None, None), dtype=torch.bfloat16)
tr = GeneralDiffusionTrainer(model, adamw(2.7e-4), EDMNoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5),
                             DiffusionInputConfig("image", (res, res, 3), []), rngs=4, name="t",
                             model_output_transform=KarrasPredictionTransform(sigma_data=0.5), ema_decay=0.999,
                             device=dev)
step = tr._define_train_step(B)
hb = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8).pin_memory()
for _ in range(3):
    tr.state, loss, tr.rngstate = step(tr.state, tr.rngstate, {"image": hb}, 0)
    loss.item()


def run(n):
    for _ in range(n):
        tr.state, loss, tr.rngstate = step(tr.state, tr.rngstate, {"image": hb}, 0)
        loss.item()


torch.cuda.synchronize()
t0 = time.perf_counter()
run(10)
t1 = time.perf_counter()
print(f"synced step: {(t1 - t0) * 100:.2f} ms")
# host time only: launch without waiting
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    tr.state, loss, tr.rngstate = step(tr.state, tr.rngstate, {"image": hb}, 0)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time per step: {(t1 - t0) * 100:.2f} ms ; total {(t2 - t0) * 100:.2f} ms")
pr = cProfile.Profile()
pr.enable()
run(10)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
