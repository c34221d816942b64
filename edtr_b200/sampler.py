"""Drop-in ``SpacedSampler`` (utils/sampler.py:67-323): same constructor, buffers, method names and
argument order.  The per-step arithmetic (x0 prediction, posterior mean, noise add) is one fused CUDA
kernel (``edtr_sampler_update``); when the model is an ``edtr_b200.ControlLDM`` the whole loop —
cross-attention K/V projection once, then every step's ControlNet+UNet pass and update — replays as
one CUDA graph.  Noise is drawn on the host side with ``torch.randn_like`` in the reference's order
(one draw per step, also at the last step — utils/sampler.py:199), so seeded runs match.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch
from torch import nn


def space_timesteps(num_timesteps: int, section_counts) -> Set[int]:
    """IDDPM respacing (utils/sampler.py:14-64)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == desired:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    start, picked = 0, []
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            picked.append(start + round(pos))
            pos += frac
        start += size
    return set(picked)


_TABLES = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
           "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")


def extract_into_tensor(a: torch.Tensor, t: torch.Tensor, x_shape) -> torch.Tensor:
    """model/gaussian_diffusion.py:34-37."""
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


class SpacedSampler(nn.Module):

    def __init__(self, betas: np.ndarray) -> "SpacedSampler":
        super().__init__()
        self.num_timesteps = len(betas)
        self.original_betas = betas
        self.original_alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.context = {}
        self._schedule_cache: Dict[Tuple, Dict[str, np.ndarray]] = {}

    def register(self, name: str, value: np.ndarray) -> None:
        self.register_buffer(name, torch.tensor(value, dtype=torch.float32))

    def make_schedule(self, num_steps: int, used_timesteps=None) -> None:
        """utils/sampler.py:85-133 — fp64 numpy tables stored as fp32 buffers (cached per key)."""
        if used_timesteps is None:
            used_timesteps = space_timesteps(self.num_timesteps, str(num_steps))
        used = set(int(u) for u in used_timesteps)
        key = (num_steps, tuple(sorted(used)))
        tabs = self._schedule_cache.get(key)
        if tabs is None:
            betas, last = [], 1.0
            for i, ac in enumerate(self.original_alphas_cumprod):
                if i in used:
                    betas.append(1 - ac / last)
                    last = ac
            assert len(betas) == num_steps
            betas = np.array(betas, dtype=np.float64)
            alphas = 1.0 - betas
            ac = np.cumprod(alphas, axis=0)
            ac_prev = np.append(1.0, ac[:-1])
            var = betas * (1.0 - ac_prev) / (1.0 - ac)
            if num_steps == 1:
                logvar = np.array([-10.0])
            else:
                logvar = np.log(np.append(var[1], var[1:]))
            tabs = dict(
                timesteps=np.array(sorted(used), dtype=np.int32),
                sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
                sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
                posterior_variance=var,
                posterior_log_variance_clipped=logvar,
                posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
                posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
            )
            self._schedule_cache[key] = tabs
        self.timesteps = tabs["timesteps"]
        for name in _TABLES:
            self.register(name, tabs[name])

    # -- kept for API parity (plain tensor arithmetic, not used by the fused path) ---------------
    def q_posterior_mean_variance(self, x_start, x_t, t):
        mean = (extract_into_tensor(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + extract_into_tensor(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        var = extract_into_tensor(self.posterior_variance, t, x_t.shape)
        logvar = extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape)
        return mean, var, logvar

    def _predict_xstart_from_eps(self, x_t, t, eps):
        return (extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def predict_noise(self, model, x, t, cond, uncond, cfg_scale):
        if uncond is None or cfg_scale == 1.:
            return model(x, t, cond)
        model_cond = model(x, t, cond)
        model_uncond = model(x, t, uncond)
        return model_uncond + cfg_scale * (model_cond - model_uncond)

    def _tables_for_kernel(self) -> List[torch.Tensor]:
        return [self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.posterior_mean_coef1,
                self.posterior_mean_coef2, self.posterior_variance]

    @torch.no_grad()
    def p_sample(self, model, x, t, index, cond, uncond, cfg_scale):
        """utils/sampler.py:184-204 -> (x_prev, pred_x0)."""
        from .engine import resolve_ops

        ops = resolve_ops()
        eps = self.predict_noise(model, x, t, cond, uncond, cfg_scale)
        noise = torch.randn_like(x)
        return ops.sampler_update(x.float().contiguous(), eps.float().contiguous(), noise.float().contiguous(),
                                  index.long().contiguous(), self._tables_for_kernel())

    # ------------------------------------------------------------------------------ loops
    def _loop(self, model, device, img, batch_size, cond, uncond, cfg_scale, tiled, tile_size, tile_stride,
              progress, progress_leave, return_intermediates):
        from .cldm import ControlLDM

        timesteps = np.flip(self.timesteps)
        total = len(self.timesteps)
        bf16 = getattr(model, "precision", "bf16") == "bf16"    # the fp32 mode steps through model(x, t, cond) + the fused update
        fused = (isinstance(model, ControlLDM) and bf16 and not tiled and (uncond is None or cfg_scale == 1.)
                 and type(model).forward is ControlLDM.forward and "forward" not in model.__dict__)
        if fused:
            x = img.float().contiguous()
            noise = [torch.randn_like(x) for _ in range(total)]
            tables = {k: getattr(self, k) for k in _TABLES}
            c_txt = cond["c_txt"]
            ck = getattr(c_txt, "_edtr_ctx_key", None)          # set by edtr_b200.clip.encode (constant-prompt cache)
            ck = ck[0] if ck is not None and ck[1] == c_txt._version else None
            out = model.engine().sample(x, [int(s) for s in timesteps], tables, cond["c_img"].float().contiguous(),
                                        c_txt.float().contiguous(), noise,
                                        control_scales=model.control_scales,
                                        return_intermediates=return_intermediates, ctx_key=ck)
            if return_intermediates:
                return out[0], out[1]
            return out
        ours = (isinstance(model, ControlLDM) and bf16 and type(model).forward is ControlLDM.forward
                and "forward" not in model.__dict__)
        if tiled and ours and (uncond is None or cfg_scale == 1.):
            # batched tiles + on-device blend instead of the per-tile Python loop
            intermediates = []
            for i, step in enumerate(timesteps):
                ts = torch.full((batch_size,), int(step), device=device, dtype=torch.long)
                index = torch.full_like(ts, fill_value=total - i - 1)
                eps = model.forward_tiled(img, ts, cond, tile_size, tile_stride)
                noise = torch.randn_like(img)
                from .engine import resolve_ops

                img, pred_x0 = resolve_ops().sampler_update(img.float().contiguous(), eps.float().contiguous(),
                                                  noise.float().contiguous(), index, self._tables_for_kernel())
                if return_intermediates:
                    intermediates.append(pred_x0)
            return (img, intermediates) if return_intermediates else img
        if tiled:
            from .tiling import make_tiled_fn

            forward = model.forward
            had_own_forward = "forward" in getattr(model, "__dict__", {})
            own_forward = model.__dict__.get("forward") if had_own_forward else None
            model.forward = make_tiled_fn(
                lambda x_tile, t, cond, hi, hi_end, wi, wi_end: forward(
                    x_tile, t, {"c_txt": cond["c_txt"], "c_img": cond["c_img"][..., hi:hi_end, wi:wi_end]}),
                tile_size, tile_stride)
        try:
            intermediates = []
            for i, step in enumerate(timesteps):
                ts = torch.full((batch_size,), int(step), device=device, dtype=torch.long)
                index = torch.full_like(ts, fill_value=total - i - 1)
                img, pred_x0 = self.p_sample(model, img, ts, index, cond, uncond, cfg_scale)
                if return_intermediates:
                    intermediates.append(pred_x0)
        finally:
            if tiled:
                # the reference leaves the wrapper installed and nests it on the next call
                # (utils/sampler.py:289-303); restoring is numerically identical.  Restore the
                # instance dict exactly: leaving a bound method behind would switch the fused paths off for good
                if had_own_forward:
                    model.__dict__["forward"] = own_forward
                else:
                    model.__dict__.pop("forward", None)
        if return_intermediates:
            return img, intermediates
        return img

    @torch.no_grad()
    def sample(self, model, device, steps, batch_size, x_size, cond, uncond, cfg_scale, tiled=False, tile_size=-1,
               tile_stride=-1, x_T=None, progress=True, progress_leave=True, return_intermediates=False):
        """utils/sampler.py:206-265."""
        self.make_schedule(steps)
        self.to(device)
        img = torch.randn((batch_size, *x_size), device=device) if x_T is None else x_T
        return self._loop(model, device, img, batch_size, cond, uncond, cfg_scale, tiled, tile_size, tile_stride,
                          progress, progress_leave, return_intermediates)

    @torch.no_grad()
    def manual_sample_with_timesteps(self, model, device, x_T, steps, used_timesteps, batch_size, cond, uncond,
                                     cfg_scale, tiled=False, tile_size=-1, tile_stride=-1, progress=True,
                                     progress_leave=True, return_intermediates=False):
        """utils/sampler.py:267-323."""
        self.make_schedule(steps, used_timesteps)
        self.to(device)
        return self._loop(model, device, x_T, batch_size, cond, uncond, cfg_scale, tiled, tile_size, tile_stride,
                          progress, progress_leave, return_intermediates)
