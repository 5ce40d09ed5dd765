"""The 1000-step cosine DDPM schedule of ``GaussianDiffusion1D.__init__``
(reference ``srcs/losses/ddpm_loss.py:50-60,110-168``), float64 → 13 fp32 buffers.

The checkpoint stores these buffers (``diffusion.*``); the loader prefers the stored values and
uses this function only to build synthetic checkpoints and to sanity-check loaded ones.
"""
import math

import torch

from .config import NUM_TIMESTEPS


def cosine_betas(timesteps=NUM_TIMESTEPS, s=0.008):
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64) / timesteps
    f = torch.cos((x + s) / (1 + s) * math.pi * 0.5) ** 2
    f = f / f[0]
    return torch.clip(1 - (f[1:] / f[:-1]), 0, 0.999)


def make_buffers(timesteps=NUM_TIMESTEPS):
    beta = cosine_betas(timesteps)
    alpha = 1.0 - beta
    abar = torch.cumprod(alpha, dim=0)
    abar_prev = torch.cat([torch.ones(1, dtype=torch.float64), abar[:-1]])
    post_var = beta * (1.0 - abar_prev) / (1.0 - abar)
    out = {
        "betas": beta,
        "alphas_cumprod": abar,
        "alphas_cumprod_prev": abar_prev,
        "sqrt_alphas_cumprod": abar.sqrt(),
        "sqrt_one_minus_alphas_cumprod": (1.0 - abar).sqrt(),
        "log_one_minus_alphas_cumprod": (1.0 - abar).log(),
        "sqrt_recip_alphas_cumprod": (1.0 / abar).sqrt(),
        "sqrt_recipm1_alphas_cumprod": (1.0 / abar - 1).sqrt(),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": post_var.clamp(min=1e-20).log(),
        "posterior_mean_coef1": beta * abar_prev.sqrt() / (1.0 - abar),
        "posterior_mean_coef2": (1.0 - abar_prev) * alpha.sqrt() / (1.0 - abar),
        "p2_loss_weight": (1 + abar / (1 - abar)) ** -0.0,
    }
    return {k: v.to(torch.float32) for k, v in out.items()}
