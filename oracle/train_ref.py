"""ORACLE (test infrastructure / CPU baseline, not product code) - one reference training step and
one sampler evaluation on the CPU in torch fp32, following trainer/general_diffusion_trainer.py:248-336
and samplers/common.py:96-109.  PARITY UNPINNED (see oracle/unet_ref.py)."""
from __future__ import annotations


import torch

from . import unet_ref


def edm_train_step(P, opt_state, images_u8, noise, t, freqs, attention_configs=(None,) * 4,
                   sigma_data=0.5, lr=2.7e-4, wd=1e-4, ema=None, ema_decay=0.999, step=1):
    """EDM schedule + Karras transform training step; returns loss.  P: dict name -> leaf tensor
    (requires_grad).  Mutates P / opt_state / ema in place (AdamW, then EMA)."""
    data = (images_u8.to(torch.float32) - 127.5) / 127.5
    sigma = torch.exp(t * 1.2 - 1.2)
    s4 = sigma.view(-1, 1, 1, 1)
    x_t = data + s4 * noise
    sd = sigma_data
    c_in = 1 / (torch.sqrt(sd ** 2 + s4 ** 2) + 1e-8)
    c_out = s4 * sd / (torch.sqrt(sd ** 2 + s4 ** 2) + 1e-8)
    c_skip = sd ** 2 / (sd ** 2 + s4 ** 2 + 1e-8)
    t_model = torch.log(sigma + 1e-12) / 4
    F = unet_ref.unet_forward(P, x_t * c_in, t_model, freqs, attention_configs=attention_configs)
    pred = c_out * F + c_skip * x_t
    w = (s4 ** 2 + sd ** 2) / ((s4 * sd) ** 2 + 1e-6)
    loss = (0.5 * (pred - data) ** 2 * w).mean()
    _adamw_ema_update(P, opt_state, loss, lr, wd, ema, ema_decay, step)
    return loss.detach()


def _adamw_ema_update(P, opt_state, loss, lr, wd, ema, ema_decay, step):
    """optax.adamw (training.py:598-608) then apply_ema (trainer/diffusion_trainer.py:31-37), in place."""
    grads = torch.autograd.grad(loss, list(P.values()))
    b1, b2, eps = 0.9, 0.999, 1e-8
    with torch.no_grad():
        for (k, p), g in zip(P.items(), grads):
            m, v = opt_state.setdefault(k, (torch.zeros_like(p), torch.zeros_like(p)))
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            mh, vh = m / (1 - b1 ** step), v / (1 - b2 ** step)
            p.sub_(lr * (mh / (vh.sqrt() + eps) + wd * p))
            if ema is not None:
                ema[k].mul_(ema_decay).add_(p, alpha=1 - ema_decay)


def ddpm_train_step(P, opt_state, images_u8, noise, t_int, freqs, attention_configs=(None,) * 4,
                    timesteps=1000, lr=2.7e-4, wd=1e-4, ema=None, ema_decay=0.999, step=1):
    """BASELINE configs[0]: LinearNoiseSchedule(1000) + EpsilonPredictionTransform training step
    (general_diffusion_trainer.py:285-302 with schedulers/discrete.py:47-57 and predictors/__init__.py:19-44):
    x_t = sqrt(acp[t]) x0 + sqrt(1-acp[t]) eps, the model sees (x_t, t) unscaled, target = eps,
    loss = mean(0.5 (pred - eps)^2 * p2_weight[t])."""
    from . import diffusion_ref as R
    data = (images_u8.to(torch.float32) - 127.5) / 127.5
    T = R.linear_tables(timesteps)
    idx = R.discrete_index(t_int.numpy(), timesteps)
    a = torch.from_numpy(T["sqrt_alpha_cumprod"][idx]).view(-1, 1, 1, 1)
    sg = torch.from_numpy(T["sqrt_one_minus_alpha_cumprod"][idx]).view(-1, 1, 1, 1)
    w = torch.from_numpy(T["p2_loss_weights"][idx]).view(-1, 1, 1, 1)
    x_t = a * data + sg * noise
    pred = unet_ref.unet_forward(P, x_t, t_int.to(torch.float32), freqs, attention_configs=attention_configs)
    loss = (0.5 * (pred - noise) ** 2 * w).mean()
    _adamw_ema_update(P, opt_state, loss, lr, wd, ema, ema_decay, step)
    return loss.detach()


def karras_denoise_eval(P, x_t, t, freqs, attention_configs=(None,) * 4, sigma_data=0.5,
                        sigma_min=0.002, sigma_max=80.0, rho=7.0):
    """One sampler model evaluation: (x0, eps) from x_t at step t in [0,1] (KarrasVE schedule)."""
    ramp = torch.clamp(1 - t, 0, 1)
    hi, lo = sigma_max ** (1 / rho), sigma_min ** (1 / rho)
    sigma = (hi + ramp * (lo - hi)) ** rho
    s4 = sigma.view(-1, 1, 1, 1)
    sd = sigma_data
    c_in = 1 / (torch.sqrt(sd ** 2 + s4 ** 2) + 1e-8)
    c_out = s4 * sd / (torch.sqrt(sd ** 2 + s4 ** 2) + 1e-8)
    c_skip = sd ** 2 / (sd ** 2 + s4 ** 2 + 1e-8)
    with torch.no_grad():
        F = unet_ref.unet_forward(P, x_t * c_in, torch.log(sigma + 1e-12) / 4, freqs,
                                  attention_configs=attention_configs)
    x0 = c_out * F + c_skip * x_t
    return x0, (x_t - x0) / s4
