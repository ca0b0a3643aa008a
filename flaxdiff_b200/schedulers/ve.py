"""Karras / EDM sigma schedules (flaxdiff/schedulers/karras.py:7-77, cosine.py:19-30)."""
from __future__ import annotations

import math

import torch

from .base import GeneralizedNoiseScheduler, as_steps
from .. import utils


class KarrasVENoiseScheduler(GeneralizedNoiseScheduler):
    """sigma(t) = (smax^(1/rho) + clip(1 - t/T, 0, 1) * (smin^(1/rho) - smax^(1/rho)))^rho."""

    def __init__(self, timesteps=1.0, sigma_min=0.002, sigma_max=80, rho=7., sigma_data=0.5, *args, **kwargs):
        super().__init__(timesteps=timesteps, sigma_min=sigma_min, sigma_max=sigma_max,
                         sigma_data=sigma_data, *args, **kwargs)
        self.min_inv_rho = sigma_min ** (1 / rho)
        self.max_inv_rho = sigma_max ** (1 / rho)
        self.rho = rho

    def get_sigmas(self, steps) -> torch.Tensor:
        steps = as_steps(steps, self._dev(), torch.float32).to(torch.float32)
        ramp = torch.clamp(1 - steps / self.max_timesteps, 0.0, 1.0)
        return (self.max_inv_rho + ramp * (self.min_inv_rho - self.max_inv_rho)) ** self.rho

    def get_weights(self, steps, shape=(-1, 1, 1, 1)):
        sigma = self.get_sigmas(steps)
        w = (sigma ** 2 + self.sigma_data ** 2) / ((sigma * self.sigma_data) ** 2 + 1e-6)
        return w.reshape(shape)

    def transform_inputs(self, x, steps, num_discrete_chunks=1000):
        return x, torch.log(self.get_sigmas(steps) + 1e-12) / 4

    def get_timesteps(self, sigmas):
        sigmas = as_steps(sigmas, self._dev(), torch.float32).reshape(-1)
        inv_rho = (sigmas + 1e-12) ** (1 / self.rho)
        denom = self.min_inv_rho - self.max_inv_rho
        if abs(denom) < 1e-7:
            denom = math.copysign(1e-7, denom)
        ramp = torch.clamp((inv_rho - self.max_inv_rho) / denom, 0.0, 1.0)
        return torch.clamp(1 - ramp, 0.0, 1.0) * self.max_timesteps

    def generate_timesteps(self, batch_size, state):
        t, state = super().generate_timesteps(batch_size, state)
        return t.to(torch.float32), state


class SimpleExpNoiseScheduler(KarrasVENoiseScheduler):
    def __init__(self, timesteps, sigma_min=0.002, sigma_max=80, rho=7., sigma_data=0.5, *args, **kwargs):
        super().__init__(timesteps=timesteps, sigma_min=sigma_min, sigma_max=sigma_max, sigma_data=sigma_data,
                         *args, **kwargs)
        n = timesteps if (isinstance(timesteps, int) and timesteps > 1) else 1000
        self._sig = torch.exp(torch.linspace(math.log(sigma_min), math.log(sigma_max), n))
        self._sig_dev = {}

    def _sig_on(self, device) -> torch.Tensor:
        """The sigma table on `device`, copied once (never inside a CUDA-graph capture: the first call
        happens in the warm-up evaluation that precedes every capture)."""
        key = str(device)
        if key not in self._sig_dev:
            self._sig_dev[key] = self._sig.to(device)
        return self._sig_dev[key]

    @property
    def sigmas(self):
        return self._sig_on(self._dev())

    def get_sigmas(self, steps):
        steps = as_steps(steps, self._dev())
        n = self._sig.shape[0]
        idx = steps.to(torch.int16).to(torch.int64)
        idx = torch.where(idx < 0, idx + n, idx).clamp_(0, n - 1)
        return self._sig_on(idx.device)[idx]


class EDMNoiseScheduler(KarrasVENoiseScheduler):
    """Training schedule: t ~ N(0,1), sigma = exp(1.2 * t/T - 1.2) (karras.py:65-77)."""

    def __init__(self, timesteps, sigma_min=0.002, sigma_max=80, rho=7., sigma_data=0.5, *args, **kwargs):
        super().__init__(timesteps=timesteps, sigma_min=sigma_min, sigma_max=sigma_max, sigma_data=sigma_data,
                         *args, **kwargs)

    def get_sigmas(self, steps, std=1.2, mean=-1.2):
        steps = as_steps(steps, self._dev(), torch.float32).to(torch.float32)
        return torch.exp((steps / self.max_timesteps) * std + mean)

    def generate_timesteps(self, batch_size, state):
        state, rng = state.get_random_key()
        return utils.device_normal(rng, (batch_size,), self._dev()), state


class CosineGeneralNoiseScheduler(GeneralizedNoiseScheduler):
    def __init__(self, sigma_min=0.02, sigma_max=80.0, kappa=1.0, *args, **kwargs):
        super().__init__(timesteps=1, sigma_min=sigma_min, sigma_max=sigma_max, *args, **kwargs)
        self.kappa = kappa
        self.theta_max = math.atan(math.exp(-(math.log(kappa) - math.log(sigma_max))))
        self.theta_min = math.atan(math.exp(-(math.log(kappa) - math.log(sigma_min))))

    def get_sigmas(self, steps):
        steps = as_steps(steps, self._dev(), torch.float32)
        return torch.tan(self.theta_min + steps * (self.theta_max - self.theta_min)) / self.kappa
