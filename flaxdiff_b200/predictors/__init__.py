"""Model-output parameterisations (API of flaxdiff/predictors/__init__.py:9-95).

Every transform is affine in (x_t, model_output) with per-sample coefficients, so each
image-sized operation is ONE pass of the libfdx `fdx_affine_combine` kernel; the (B,)-sized
coefficient algebra below is host-side logic and follows the reference formulas exactly
(including the +1e-8 stabilisers of KarrasPredictionTransform, predictors/__init__.py:84-96).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .. import ops
from ..schedulers import NoiseScheduler, get_coeff_shapes_tuple

TARGET_X0, TARGET_EPS, TARGET_V = ops.TARGET_X0, ops.TARGET_EPS, ops.TARGET_V


def _vec(c, B, device) -> torch.Tensor:
    """coefficient -> (B,) f32 tensor."""
    if not isinstance(c, torch.Tensor):
        return torch.full((B,), float(c), dtype=torch.float32, device=device)
    c = c.to(device=device, dtype=torch.float32).reshape(-1)
    return c.expand(B) if c.numel() == 1 else c


def _affine(inputs: Sequence[torch.Tensor], coefs: Sequence, coefs2: Sequence = None, clip=None):
    """sum_i coefs[i] * inputs[i] (and optionally a second combination) on the GPU."""
    x = inputs[0]
    B = x.shape[0]
    ins = [t.to(torch.float32).contiguous() for t in inputs]
    c1 = torch.stack([_vec(c, B, x.device) for c in coefs])
    c2 = torch.stack([_vec(c, B, x.device) for c in coefs2]) if coefs2 is not None else None
    o1, o2, _ = ops.affine_combine(ins, c1, c2, clip=clip)
    return (o1, o2) if coefs2 is not None else o1


def _flat_rates(rates):
    a, s = rates
    return a.reshape(-1).to(torch.float32), s.reshape(-1).to(torch.float32)


class DiffusionPredictionTransform:
    target_kind = TARGET_X0

    # ---- (B,)-sized coefficient algebra --------------------------------------
    def loss_coeffs(self, rates) -> Tuple[torch.Tensor, torch.Tensor]:
        """pred_transform(x_t, F) = c_out * F + c_skip * x_t."""
        a, s = _flat_rates(rates)
        return torch.ones_like(s), torch.zeros_like(s)

    def backward_coeffs(self, rates):
        """backward_diffusion on pred: x0 = p*x_t + q*pred ; eps = r*x_t + u*pred."""
        raise NotImplementedError

    def x0_eps_coeffs(self, rates):
        """Composition pred_transform -> backward_diffusion in terms of the raw model output F."""
        c_out, c_skip = self.loss_coeffs(rates)
        p, q, r, u = self.backward_coeffs(rates)
        return p + q * c_skip, q * c_out, r + u * c_skip, u * c_out

    def get_input_scale(self, rates):
        return 1

    # ---- image-sized ops (libfdx) ----------------------------------------------
    def pred_transform(self, x_t, preds, rates) -> torch.Tensor:
        c_out, c_skip = self.loss_coeffs(rates)
        return _affine([x_t, preds], [c_skip, c_out])

    def backward_diffusion(self, x_t, preds, rates):
        p, q, r, u = self.backward_coeffs(rates)
        return _affine([x_t, preds], [p, q], [r, u])

    def __call__(self, x_t, preds, current_step, noise_schedule: NoiseScheduler):
        rates = noise_schedule.get_rates(current_step, shape=get_coeff_shapes_tuple(x_t))
        p, q, r, u = self.x0_eps_coeffs(rates)
        return _affine([x_t, preds], [p, q], [r, u])

    def forward_diffusion(self, x_0, epsilon, rates):
        """-> (x_t, c_in, target) (predictors/__init__.py:19-24) in one fused pass."""
        a, s = _flat_rates(rates)
        B = x_0.shape[0]
        c_in = self.get_input_scale(rates)
        cvec = _vec(c_in.reshape(-1) if isinstance(c_in, torch.Tensor) else c_in, B, x_0.device)
        x_t, target, _ = ops.diffuse_forward(x_0.to(torch.float32).contiguous(), epsilon.contiguous(),
                                             _vec(a, B, x_0.device), _vec(s, B, x_0.device), cvec,
                                             False, self.target_kind)
        return x_t, c_in, target

    def get_target(self, x_0, epsilon, rates):
        return x_0


class EpsilonPredictionTransform(DiffusionPredictionTransform):
    target_kind = TARGET_EPS

    def backward_coeffs(self, rates):
        a, s = _flat_rates(rates)
        return 1 / a, -s / a, torch.zeros_like(a), torch.ones_like(a)

    def get_target(self, x_0, epsilon, rates):
        return epsilon


class DirectPredictionTransform(DiffusionPredictionTransform):
    def backward_coeffs(self, rates):
        a, s = _flat_rates(rates)
        return torch.zeros_like(a), torch.ones_like(a), 1 / s, -a / s


class VPredictionTransform(DiffusionPredictionTransform):
    target_kind = TARGET_V

    def backward_coeffs(self, rates):
        a, s = _flat_rates(rates)
        var = a ** 2 + s ** 2
        sd = torch.sqrt(var)
        return a / var, -s * sd / var, s / var, a * sd / var

    def get_target(self, x_0, epsilon, rates):
        a, s = _flat_rates(rates)
        sd = torch.sqrt(a ** 2 + s ** 2)
        return _affine([epsilon, x_0], [a / sd, -s / sd])


class KarrasPredictionTransform(DiffusionPredictionTransform):
    def __init__(self, sigma_data=0.5) -> None:
        super().__init__()
        self.sigma_data = sigma_data

    def loss_coeffs(self, rates, epsilon=1e-8):
        _, s = _flat_rates(rates)
        sd = self.sigma_data
        c_out = s * sd / (torch.sqrt(sd ** 2 + s ** 2) + epsilon)
        c_skip = sd ** 2 / (sd ** 2 + s ** 2 + epsilon)
        return c_out, c_skip

    def backward_coeffs(self, rates):
        a, s = _flat_rates(rates)
        return torch.zeros_like(a), torch.ones_like(a), 1 / s, -a / s

    def get_input_scale(self, rates, epsilon=1e-8):
        _, s = rates
        return 1 / (torch.sqrt(self.sigma_data ** 2 + s ** 2) + epsilon)
