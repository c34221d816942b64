"""Diffusion schedule and forward noising (model/gaussian_diffusion.py:9-37, 40-84): the input preparation of the
restore path (``x_T = q_sample(z, t=200, noise)``, main/det/test_edtr.py:125-127).  Training losses are out of scope."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


def make_beta_schedule(schedule: str, n_timestep: int, linear_start: float = 1e-4, linear_end: float = 2e-2,
                       cosine_s: float = 8e-3) -> np.ndarray:
    """model/gaussian_diffusion.py:9-31 — "linear" is a linspace of sqrt(beta), squared (fp64)."""
    if schedule == "linear":
        betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    elif schedule == "sqrt_linear":
        betas = np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
    elif schedule == "sqrt":
        betas = np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64) ** 0.5
    elif schedule == "cosine":
        t = np.arange(n_timestep + 1, dtype=np.float64) / n_timestep + cosine_s
        alphas = np.cos(t / (1 + cosine_s) * np.pi / 2) ** 2
        alphas = alphas / alphas[0]
        betas = np.clip(1 - alphas[1:] / alphas[:-1], a_min=0, a_max=0.999)
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas


def extract_into_tensor(a: torch.Tensor, t: torch.Tensor, x_shape) -> torch.Tensor:
    """model/gaussian_diffusion.py:34-37."""
    b = t.shape[0]
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


class Diffusion(nn.Module):
    """model/gaussian_diffusion.py:40-84 (ctor, buffers, ``q_sample``)."""

    def __init__(self, timesteps=1000, beta_schedule="linear", loss_type="l2", linear_start=1e-4, linear_end=2e-2,
                 cosine_s=8e-3, parameterization="eps"):
        super().__init__()
        assert parameterization in ["eps", "x0", "v"], "currently only supporting 'eps' and 'x0' and 'v'"
        self.num_timesteps = timesteps
        self.beta_schedule = beta_schedule
        self.linear_start = linear_start
        self.linear_end = linear_end
        self.cosine_s = cosine_s
        self.parameterization = parameterization
        self.loss_type = loss_type
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end,
                                   cosine_s=cosine_s)
        ac = np.cumprod(1.0 - betas, axis=0)
        self.betas = betas
        for name, v in (("sqrt_alphas_cumprod", np.sqrt(ac)), ("sqrt_one_minus_alphas_cumprod", np.sqrt(1.0 - ac)),
                        ("sqrt_recip_alphas_cumprod", np.sqrt(1.0 / ac)),
                        ("sqrt_recipm1_alphas_cumprod", np.sqrt(1.0 / ac - 1))):
            self.register_buffer(name, torch.tensor(v, dtype=torch.float32))

    def q_sample(self, x_start: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
                extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def p_losses(self, *a, **k):
        raise NotImplementedError("training losses are outside the accelerated path (SURVEY §2)")
