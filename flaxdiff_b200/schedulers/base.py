"""Noise-schedule base classes (API of flaxdiff/schedulers/common.py:16-102) on torch tensors.

Schedules only ever touch (B,)-sized vectors: that is host-side control logic, so it is plain
torch on whatever device the step tensor lives on (CPU tensors work - the "-m not gpu" tests use
them).  Anything image-sized goes through libfdx kernels (see predictors / samplers / trainer).
"""
from __future__ import annotations

from typing import Tuple

import torch

from .. import utils
from ..utils import RandomMarkovState


def default_device() -> torch.device:
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def get_coeff_shapes_tuple(array) -> tuple:
    """(-1, 1, 1, ...) broadcast shape for per-sample coefficients (schedulers/common.py:6-8)."""
    return (-1,) + (1,) * (array.dim() - 1)


def reshape_rates(rates, shape=(-1, 1, 1, 1)):
    a, s = rates
    return torch.reshape(a, shape), torch.reshape(s, shape)


def as_steps(steps, device=None, dtype=None) -> torch.Tensor:
    if isinstance(steps, torch.Tensor):
        return steps
    return torch.as_tensor(steps, device=device if device is not None else default_device(), dtype=dtype)


class NoiseScheduler:
    """schedulers/common.py:16-64.  `timesteps` int > 1 => integer steps in [0, T); else uniform(0, T)."""

    def __init__(self, timesteps, dtype=torch.float32, clip_min=-1.0, clip_max=1.0, *args, **kwargs):
        self.max_timesteps = timesteps
        self.dtype = dtype
        self.clip_min = clip_min
        self.clip_max = clip_max
        self.device = kwargs.get("device", None)
        self._integer_steps = isinstance(timesteps, int) and timesteps > 1

    def _dev(self):
        return self.device if self.device is not None else default_device()

    def to(self, device):
        self.device = torch.device(device)
        return self

    def generate_timesteps(self, batch_size, state: RandomMarkovState) -> Tuple[torch.Tensor, RandomMarkovState]:
        state, rng = state.get_random_key()
        if self._integer_steps:
            t = utils.device_randint(rng, (batch_size,), 0, int(self.max_timesteps), self._dev())
        else:
            t = utils.device_uniform(rng, (batch_size,), self._dev()) * float(self.max_timesteps)
        return t, state

    def get_weights(self, steps, shape=(-1, 1, 1, 1)):
        raise NotImplementedError

    def get_rates(self, steps, shape=(-1, 1, 1, 1)):
        raise NotImplementedError

    def add_noise(self, images, noise, steps):
        from ..predictors import _affine
        a, s = self.get_rates(steps, shape=(-1,))
        return _affine([images, noise], [a, s])

    def remove_all_noise(self, noisy_images, noise, steps, clip_denoised=True, rates=None):
        from ..predictors import _affine
        a, s = self.get_rates(steps, shape=(-1,))
        return _affine([noisy_images, noise], [1.0 / a, -s / a])

    def transform_inputs(self, x, steps):
        return x, steps

    def get_posterior_mean(self, x_0, x_t, steps):
        raise NotImplementedError

    def get_posterior_variance(self, steps, shape=(-1, 1, 1, 1)):
        raise NotImplementedError

    def get_max_variance(self, shape=(-1, 1, 1, 1)):
        a, s = self.get_rates(as_steps(self.max_timesteps, self._dev()), shape=shape)
        return torch.sqrt(a ** 2 + s ** 2)


class GeneralizedNoiseScheduler(NoiseScheduler):
    """VE / EDM family: signal rate 1, noise rate sigma(t) (schedulers/common.py:66-102)."""

    def __init__(self, timesteps, sigma_min=0.002, sigma_max=80.0, sigma_data=1, *args, **kwargs):
        super().__init__(timesteps, *args, **kwargs)
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.sigma_data = sigma_data

    def get_sigmas(self, steps) -> torch.Tensor:
        raise NotImplementedError("implemented by subclasses")

    def get_weights(self, steps, shape=(-1, 1, 1, 1)):
        sigma = self.get_sigmas(steps)
        return (1 + (1 / (1 + ((1 - sigma ** 2) / (sigma ** 2)))) / (self.sigma_max ** 2)).reshape(shape)

    def get_rates(self, steps, shape=(-1, 1, 1, 1)):
        sigmas = self.get_sigmas(steps)
        return reshape_rates((torch.ones_like(sigmas), sigmas), shape=shape)

    def transform_inputs(self, x, steps, num_discrete_chunks=1000):
        steps = as_steps(steps, self._dev())
        chunks = (steps / self.max_timesteps) * num_discrete_chunks
        return x, chunks.to(torch.int32)

    def get_timesteps(self, sigmas):
        raise NotImplementedError("implemented by subclasses")


class ContinuousNoiseScheduler(NoiseScheduler):
    """schedulers/continuous.py:7-11: t in [0, 1)."""

    def __init__(self, *args, **kwargs):
        super().__init__(timesteps=1, *args, **kwargs)
