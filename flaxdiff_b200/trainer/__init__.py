"""Training: `TrainState`, optimiser descriptions and `GeneralDiffusionTrainer`
(API of flaxdiff/trainer/{diffusion_trainer.py:27-37, general_diffusion_trainer.py:108-349,
simple_trainer.py:500-677}).

One training step (general_diffusion_trainer.py:248-336) is, on each GPU:
    normalise + noise-add + precondition      fdx_diffuse_forward          (1 kernel)
    UNet forward / loss / UNet backward       Unet.forward / fdx_loss_fwd_bwd / Unet.backward
    gradient mean over ranks                  ONE NCCL all-reduce on the flat f32 gradient buffer
    AdamW + EMA + bf16 shadow refresh         fdx_adamw_ema_step           (1 kernel)
The forward/backward part is captured in a CUDA graph after the first (eager) step; timesteps
and noise are drawn outside the graph.  Out of scope (SURVEY.md $2.1 row 4): wandb, orbax,
registry pushes, AutoEncoderTrainer.
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass, field, replace
from typing import Any, Callable, Dict, Optional, Tuple, Type, Union

import torch

from .. import ops, utils
from .._lib import FdxError
from ..models.params import FlatParams
from ..predictors import DiffusionPredictionTransform, EpsilonPredictionTransform, _vec
from ..samplers import DDIMSampler, DiffusionSampler
from ..schedulers import NoiseScheduler
from ..utils import RandomMarkovState


# --------------------------------------------------------------------------- optimisers
@dataclass
class Optimizer:
    """Description of an optax chain (training.py:594-608): [clip_by_global_norm] -> adam/adamw."""
    learning_rate: Union[float, Callable[[int], float]] = 2.7e-4
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 0.0
    clip_norm: float = 0.0

    def lr_at(self, count: int) -> float:
        lr = self.learning_rate
        return float(lr(count)) if callable(lr) else float(lr)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8) -> Optimizer:
    return Optimizer(learning_rate, b1, b2, eps, 0.0)


def adamw(learning_rate, b1=0.9, b2=0.999, eps=1e-8, weight_decay=1e-4) -> Optimizer:
    return Optimizer(learning_rate, b1, b2, eps, weight_decay)


def clip_by_global_norm(max_norm: float) -> Optimizer:
    return Optimizer(clip_norm=max_norm)


def chain(*parts: Optimizer) -> Optimizer:
    out = Optimizer()
    for p in parts:
        if p.clip_norm > 0:
            out = replace(out, clip_norm=p.clip_norm)
        else:
            out = replace(p, clip_norm=out.clip_norm)
    return out


def warmup_cosine_decay_schedule(init_value, peak_value, warmup_steps, decay_steps, end_value=0.0):
    """optax.warmup_cosine_decay_schedule (training.py:263-267)."""
    def sched(count: int) -> float:
        if count < warmup_steps:
            return init_value + (peak_value - init_value) * count / max(warmup_steps, 1)
        t = min(max(count - warmup_steps, 0) / max(decay_steps - warmup_steps, 1), 1.0)
        return end_value + (peak_value - end_value) * 0.5 * (1 + math.cos(math.pi * t))
    return sched


# --------------------------------------------------------------------------- data parallel
def dp_allreduce_sum_(grads_flat: torch.Tensor, loss: torch.Tensor, world_size: int) -> float:
    """The ONE exchange step of data-parallel training (jax.lax.pmean of grads and loss,
    general_diffusion_trainer.py:325,334): sum all-reduce of the flat gradient buffer and of the
    scalar loss over the process group (NCCL over NVLink on GPUs); returns the 1/world factor that
    the fused optimiser kernel applies to turn the sum into the mean."""
    if world_size <= 1:
        return 1.0
    torch.distributed.all_reduce(grads_flat)
    torch.distributed.all_reduce(loss)
    return 1.0 / world_size


def rank_key(key, rank: int):
    """Per-rank RNG stream: fold the rank into the step key (general_diffusion_trainer.py:251-253)."""
    return utils.fold_in(key, int(rank))


# --------------------------------------------------------------------------- state
@dataclass
class TrainState:
    """flax TrainState + ema_params (trainer/diffusion_trainer.py:27-37)."""
    step: int
    params: FlatParams
    ema_params: FlatParams
    opt_state: Dict[str, Any]          # {'mu': flat f32, 'nu': flat f32, 'count': int}
    tx: Optimizer
    apply_fn: Callable = None
    rngs: Tuple[int, int] = (0, 0)
    dynamic_scale: Any = None

    @classmethod
    def create(cls, apply_fn, params: FlatParams, ema_params: FlatParams, tx: Optimizer, rngs=(0, 0),
               dynamic_scale=None, **_):
        z = torch.zeros_like(params.flat)
        return cls(0, params, ema_params, {"mu": z, "nu": z.clone(), "count": 0}, tx, apply_fn, tuple(rngs),
                   dynamic_scale)

    def apply_gradients_and_ema(self, grads: FlatParams, ema_decay: float, grad_scale: float = 1.0,
                                dyn: torch.Tensor = None) -> "TrainState":
        """optimiser update followed by EMA (general_diffusion_trainer.py:327-330), fused."""
        tx = self.tx
        count = self.opt_state["count"] + 1
        gn = ops.sumsq(grads.flat) if tx.clip_norm > 0 else None
        if dyn is not None:
            host = torch.tensor([tx.lr_at(count - 1), 1 - tx.b1 ** count, 1 - tx.b2 ** count], dtype=torch.float32)
            dyn.copy_(host, non_blocking=True)
        ops.adamw_ema_step(self.params.flat, grads.flat, self.opt_state["mu"], self.opt_state["nu"],
                           self.ema_params.flat, self.params.shadow_flat_noupdate(), tx.lr_at(count - 1), tx.b1,
                           tx.b2, tx.eps, tx.weight_decay, count, ema_decay, grad_scale, gn, tx.clip_norm, dyn)
        self.params.mark_shadow_fresh()
        self.opt_state["count"] = count
        self.step += 1
        return self

    def apply_ema(self, decay: float = 0.999):
        return self


# --------------------------------------------------------------------------- trainer
class GeneralDiffusionTrainer:
    def __init__(self, model, optimizer: Optimizer, noise_schedule: NoiseScheduler, input_config,
                 rngs, unconditional_prob: float = 0.12, name: str = "GeneralDiffusion",
                 model_output_transform: DiffusionPredictionTransform = None, autoencoder=None,
                 native_resolution: int = None, frames_per_sample: int = None, wandb_config=None,
                 eval_metrics=None, best_tracker_metric: str = "train/best_loss",
                 distributed_training: bool = None, checkpoint_base_path: str = "./checkpoints",
                 use_dynamic_scale: bool = False, ema_decay: float = 0.999, device=None,
                 use_cuda_graph: bool = True, **kwargs):
        if autoencoder is not None:
            raise FdxError("latent diffusion (autoencoder) is outside the supported hot path")
        if use_dynamic_scale:
            raise FdxError("DynamicScale loss scaling is not implemented (bf16 needs none)")
        self.model = model
        self.optimizer = optimizer
        self.noise_schedule = noise_schedule
        self.input_config = input_config
        self.model_output_transform = model_output_transform or EpsilonPredictionTransform()
        self.unconditional_prob = unconditional_prob
        self.name = name
        self.ema_decay = ema_decay
        self.autoencoder = None
        self.checkpoint_base_path = checkpoint_base_path
        self.use_cuda_graph = use_cuda_graph
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        noise_schedule.to(self.device)
        shape = input_config.sample_data_shape
        self.native_resolution = native_resolution or shape[-2]
        dist = torch.distributed
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world_size > 1 else 0
        self.distributed_training = (self.world_size > 1) if distributed_training is None else distributed_training
        if isinstance(rngs, int):
            rngs = utils.PRNGKey(rngs)
        self.rngstate = RandomMarkovState(tuple(rngs))
        self.state, self.best_state = self.generate_states(optimizer, tuple(rngs), model=model)
        self.best_loss = 1e9
        self._graph = None
        self._static = None
        self._dyn = torch.zeros(3, dtype=torch.float32, device=self.device)
        self._grads = self.state.params.zeros_like()

    # ------------------------------------------------------------------ state
    def generate_states(self, optimizer, rngs, existing_state=None, existing_best_state=None, model=None,
                        param_transforms=None, use_dynamic_scale=False):
        rngs, subkey = utils.split(rngs)
        if existing_state is None:
            ctx_ex = None
            if self.input_config.conditions:
                ctx_ex = self.input_config.conditions[0].get_unconditional()
            params = model.init(subkey, textcontext=ctx_ex, device=self.device)
            ema = params.clone()
        else:
            params, ema = existing_state['params'], existing_state['ema_params']
        state = TrainState.create(apply_fn=model.apply, params=params, ema_params=ema, tx=optimizer, rngs=rngs)
        return state, state

    # ------------------------------------------------------------------ the step
    def _fwd_bwd(self, images, noise, noise_level, ctx=None):
        """noise-add -> UNet fwd -> loss -> UNet bwd; returns the loss tensor (f32[1]); gradients land in
        self._grads.  FDX_MICROBATCH=2 processes the batch as two micro-batches on two streams (the sum of
        the two mean-loss gradients, halved, is the full-batch gradient: GroupNorm is per sample) so that
        the HBM-bound kernels of one half overlap the tensor-core kernels of the other; measured slower
        than one batch with side-stream weight gradients (24.2 vs 22.3 ms/step at C2), hence opt-in."""
        B = images.shape[0]
        nmb = int(os.environ.get("FDX_MICROBATCH", "1"))
        if nmb < 2 or B % nmb != 0 or B // nmb < 2:
            return self._fwd_bwd_one(images, noise, noise_level, ctx, self._grads)
        if getattr(self, "_mb_grads", None) is None or len(self._mb_grads) != nmb:
            self._mb_grads = [self.state.params.zeros_like() for _ in range(nmb)]
            self._mb_streams = [torch.cuda.Stream() for _ in range(nmb)]
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        losses = []
        step = B // nmb
        for h in range(nmb):
            sl = slice(h * step, (h + 1) * step)
            self._mb_streams[h].wait_event(ev)
            with torch.cuda.stream(self._mb_streams[h]):
                losses.append(self._fwd_bwd_one(images[sl], noise[sl], noise_level[sl],
                                                None if ctx is None else ctx[sl], self._mb_grads[h]))
        for h in range(nmb):
            main.wait_stream(self._mb_streams[h])
        g = self._grads.flat
        if nmb == 2:
            torch.lerp(self._mb_grads[0].flat, self._mb_grads[1].flat, 0.5, out=g)
        else:
            torch.add(self._mb_grads[0].flat, self._mb_grads[1].flat, out=g)
            for h in range(2, nmb):
                g.add_(self._mb_grads[h].flat)
            g.mul_(1.0 / nmb)
        loss = losses[0]
        for h in range(1, nmb):
            loss = loss + losses[h]
        return loss / nmb

    def _fwd_bwd_one(self, images, noise, noise_level, ctx, grads):
        sched, tr, st = self.noise_schedule, self.model_output_transform, self.state
        B = images.shape[0]
        rates = sched.get_rates(noise_level, shape=(-1,))
        alpha, sigma = rates
        c_in = tr.get_input_scale(rates)
        c_in = _vec(c_in, B, images.device)
        c_out, c_skip = tr.loss_coeffs(rates)
        weight = sched.get_weights(noise_level, shape=(-1,)).to(torch.float32)
        _, t_model = sched.transform_inputs(None, noise_level)
        x_t, target, model_in = ops.diffuse_forward(images, noise, alpha, sigma, c_in, True, tr.target_kind)
        F, saved = self.model.forward(st.params, model_in, t_model, ctx, save=True)
        loss, dF = ops.loss_fwd_bwd(F, x_t, target, c_out, c_skip, weight, want_grad=True)
        grads.flat.zero_()
        self.model.backward(st.params, saved, dF, grads)
        return loss

    def _define_train_step(self, batch_size=None):
        key = self.input_config.sample_data_key
        dev = self.device

        def train_step(train_state: TrainState, rng_state: RandomMarkovState, batch, local_device_index=None):
            rng_state, key_fold = rng_state.get_random_key()
            idx = self.rank if local_device_index is None else int(local_device_index)
            local = RandomMarkovState(rank_key(key_fold, idx))
            data = batch[key]
            if not isinstance(data, torch.Tensor):
                data = torch.as_tensor(data)
            if data.dtype not in (torch.uint8, torch.float32):
                data = data.to(torch.float32)
            images = data.to(dev, non_blocking=True).contiguous()      # HOST -> DEVICE boundary
            B = images.shape[0]
            local, uncond_key = local.get_random_key()
            ctx = None
            if self.input_config.conditions:
                # per-sample unconditional mask (bernoulli p = unconditional_prob) and null-embedding mixing
                # (general_diffusion_trainer.py:266-275; inputs/__init__.py:123-146)
                mask = utils.device_uniform(uncond_key, (B,), dev) < self.unconditional_prob
                ctx = self.input_config.process_conditioning(batch, uncond_mask=mask)[0]
                ctx = ctx.to(device=dev, dtype=torch.float32).contiguous()
            noise_level, local = self.noise_schedule.generate_timesteps(B, local)
            local, noise_key = local.get_random_key()
            noise = utils.device_normal(noise_key, tuple(images.shape), dev)
            if self.use_cuda_graph:
                loss = self._graphed_fwd_bwd(images, noise, noise_level, ctx)
            else:
                loss = self._fwd_bwd(images, noise, noise_level, ctx)
            gscale = 1.0
            if self.distributed_training and self.world_size > 1:
                gscale = dp_allreduce_sum_(self._grads.flat, loss, self.world_size)
                loss = loss * gscale
            train_state.apply_gradients_and_ema(self._grads, self.ema_decay, gscale, self._dyn)
            return train_state, loss, rng_state

        return train_step

    def _graphed_fwd_bwd(self, images, noise, noise_level, ctx=None):
        if self._graph is None or self._static[0].shape != images.shape or self._static[0].dtype != images.dtype:
            self._static = [images.clone(), noise.clone(), noise_level.clone()]
            if ctx is not None:
                self._static.append(ctx.clone())
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._fwd_bwd(*self._static)        # eager warm-up on a side stream
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._static_loss = self._fwd_bwd(*self._static)
        self._static[0].copy_(images, non_blocking=True)
        self._static[1].copy_(noise)
        self._static[2].copy_(noise_level)
        if ctx is not None:
            self._static[3].copy_(ctx)
        self._graph.replay()
        return self._static_loss.clone()

    # ------------------------------------------------------------------ loops
    def fit(self, data, training_steps_per_epoch, epochs, val_steps_per_epoch=8,
            sampler_class: Type[DiffusionSampler] = DDIMSampler, sampling_noise_schedule=None, verbose=True):
        """Host loop of simple_trainer.py:601-677 without wandb / orbax / validation sampling."""
        train_ds = iter(data['train']())
        step_fn = self._define_train_step(data.get('local_batch_size'))
        for epoch in range(epochs):
            t0 = time.time()
            tot = 0.0
            for i in range(training_steps_per_epoch):
                batch = next(train_ds)
                self.state, loss, self.rngstate = step_fn(self.state, self.rngstate, batch, self.rank)
                if i % 100 == 0 or i == training_steps_per_epoch - 1:
                    lv = float(loss.item())
                    if not math.isfinite(lv) or lv <= 1e-8:
                        raise FdxError(f"loss became invalid ({lv}) at step {self.state.step}")
                    tot = lv
            if verbose and self.rank == 0:
                dt = time.time() - t0
                print(f"epoch {epoch}: last loss {tot:.5f}, {dt / training_steps_per_epoch * 1e3:.1f} ms/step")
            if tot < self.best_loss:
                self.best_loss = tot
        return self.state

    def make_sampler(self, sampler_class: Type[DiffusionSampler] = DDIMSampler, sampling_noise_schedule=None,
                     guidance_scale: float = 0.0):
        return sampler_class(model=self.model, noise_schedule=sampling_noise_schedule or self.noise_schedule,
                             model_output_transform=self.model_output_transform, input_config=self.input_config,
                             guidance_scale=guidance_scale)

    # ------------------------------------------------------------------ checkpoints
    def save(self, path: str):
        st = self.state
        torch.save({"layout": list(st.params.layout.table.items()), "params": st.params.flat.cpu(),
                    "ema_params": st.ema_params.flat.cpu(), "mu": st.opt_state["mu"].cpu(),
                    "nu": st.opt_state["nu"].cpu(), "count": st.opt_state["count"], "step": st.step,
                    "rngs": self.rngstate.rng, "best_loss": self.best_loss}, path)

    def load(self, path: str):
        ck = torch.load(path, map_location="cpu")
        st = self.state
        st.params.flat.copy_(ck["params"])
        st.ema_params.flat.copy_(ck["ema_params"])
        st.opt_state["mu"].copy_(ck["mu"])
        st.opt_state["nu"].copy_(ck["nu"])
        st.opt_state["count"] = ck["count"]
        st.step = ck["step"]
        self.rngstate = RandomMarkovState(tuple(ck["rngs"]))
        self.best_loss = ck["best_loss"]
        return self


DiffusionTrainer = GeneralDiffusionTrainer
