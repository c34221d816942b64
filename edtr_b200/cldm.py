"""Drop-in ``ControlLDM`` (model/cldm.py:17-194): same constructor, attributes, weight loaders and
call signatures; ``forward`` / ``vae_decode`` execute on the sm_100a kernels through ``engine.py``.

Select it from a reference config by changing ``target: model.cldm.ControlLDM`` to
``target: edtr_b200.cldm.ControlLDM`` (the reference instantiates models by dotted path,
utils/common.py:23-34), or build it from an existing reference instance with
``ControlLDM.from_reference(ref_model)``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Set, Tuple

import torch
from torch import nn

from .nets import AutoencoderKL, ControlledUnetModel, ControlNet, state_version


def disabled_train(self: nn.Module) -> nn.Module:
    return self


class ControlLDM(nn.Module):

    def __init__(self, unet_cfg, vae_cfg, clip_cfg, controlnet_cfg, latent_scale_factor, tail_block=False,
                 clip: Optional[nn.Module] = None):
        super().__init__()
        if tail_block:
            raise NotImplementedError("tail_block / woSD is not used by any EDTR config or script (SURVEY §8 a4)")
        self.unet = ControlledUnetModel(**unet_cfg)
        self.vae = AutoencoderKL(**vae_cfg)
        # The OpenCLIP text tower produces an INPUT of the path (c_txt for the constant "" prompt, model/clip.py:41-63):
        # `self.clip` is built from clip_cfg exactly as the reference does (model/cldm.py:31) — edtr_b200.clip keeps the
        # constructor, state-dict keys and encode() of FrozenOpenCLIPEmbedder and caches the embedding of the constant
        # prompt.  clip_cfg=None (synthetic c_txt, as in the benchmarks) leaves it out; a ready embedder may be passed.
        if clip is None and clip_cfg is not None:
            from .clip import FrozenOpenCLIPEmbedder

            clip = FrozenOpenCLIPEmbedder(**clip_cfg)
        self.clip = clip
        self.clip_cfg = clip_cfg
        self.controlnet = ControlNet(**controlnet_cfg)
        self.scale_factor = latent_scale_factor
        self.control_scales = [1.0] * 13
        # Tile-parallel execution of the tiled paths (cldm-tiled sampling, VAEHook encode / decode; config C4) is an
        # explicit opt-in: None (default) = every rank works alone, exactly like the reference; True / a ProcessGroup =
        # the tiles of ONE image, identical on every rank of the group, are spread over its ranks (parallel.tile_sharding).
        self.tile_group = None
        self.tile_group_check = True   # verify (one tiny collective per call) that the ranks really hold the same input
        self._engine = None
        self._engine_version = None
        # "bf16" (default): the tensor-core engine.  "fp32": every tensor and every accumulation in fp32 through the
        # edtr_f32_* kernels (engine_f32.py) — BASELINE.json's fp32 tolerance (per-step latent max-rel error <= 1e-4)
        self.precision = "bf16"
        self._engine32 = None
        self._engine32_version = None
        self._vae32 = None
        self._vae32_version = None

    def set_precision(self, precision: str) -> "ControlLDM":
        """"bf16": tensor-core kernels, CUDA graphs (throughput mode).  "fp32": fp32 storage and accumulation on the CUDA
        cores (accuracy mode; forward / sampling / vae_decode, untiled)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        return self

    def engine_f32(self):
        from .engine_f32 import CldmEngineF32

        dev = next(self.unet.parameters()).device
        ver = (state_version(self.unet), state_version(self.controlnet), dev)
        if self._engine32 is None or self._engine32_version != ver:
            self._engine32 = CldmEngineF32(self.unet.cfg, self.controlnet.cfg, self.unet.state_dict(),
                                           self.controlnet.state_dict(), dev)
            self._engine32_version = ver
        return self._engine32

    def _vae_decoder_f32(self):
        from .engine_f32 import VaeDecoderF32

        dev = next(self.vae.parameters()).device
        ver = (state_version(self.vae), dev)
        if self._vae32 is None or self._vae32_version != ver:
            self._vae32 = VaeDecoderF32(self.vae.ddconfig, self.vae.embed_dim, self.vae.state_dict(), dev)
            self._vae32_version = ver
        return self._vae32

    def _vae_encoder_f32(self):
        from .engine_f32 import VaeEncoderF32

        dev = next(self.vae.parameters()).device
        ver = (state_version(self.vae), dev)
        if getattr(self, "_vae32e", None) is None or self._vae32e_version != ver:
            self._vae32e = VaeEncoderF32(self.vae.ddconfig, self.vae.embed_dim, self.vae.state_dict(), dev)
            self._vae32e_version = ver
        return self._vae32e

    # ----------------------------------------------------------------- construction helpers
    @classmethod
    def from_reference(cls, ref: nn.Module, unet_cfg: Dict, vae_cfg: Dict, controlnet_cfg: Dict) -> "ControlLDM":
        """Adopt the weights (and CLIP embedder) of a reference ``model.cldm.ControlLDM`` instance."""
        m = cls(unet_cfg, vae_cfg, None, controlnet_cfg, ref.scale_factor, clip=getattr(ref, "clip", None))
        m.unet.load_state_dict(ref.unet.state_dict(), strict=True)
        m.controlnet.load_state_dict(ref.controlnet.state_dict(), strict=True)
        m.vae.load_state_dict(ref.vae.state_dict(), strict=True)
        m.control_scales = list(ref.control_scales)
        return m

    # ----------------------------------------------------------------- reference weight loaders
    @torch.no_grad()
    def load_pretrained_sd(self, sd: Dict[str, torch.Tensor], is_turbo: bool = False) -> Set[str]:
        """model/cldm.py:46-77: SD-2.1 checkpoint keys -> unet / vae (/ clip when an embedder is attached)."""
        module_map = {"unet": "model.diffusion_model", "vae": "first_stage_model",
                      "clip": "conditioner.embedders.0" if is_turbo else "cond_stage_model"}
        modules = [("unet", self.unet), ("vae", self.vae)]
        if self.clip is not None:
            modules.append(("clip", self.clip))
        used = set()
        for name, module in modules:
            init_sd = {}
            for key in module.state_dict():
                target_key = ".".join([module_map[name], key])
                init_sd[key] = sd[target_key].clone()
                used.add(target_key)
            module.load_state_dict(init_sd, strict=True)
        for module in [m for m in (self.clip, self.unet) if m is not None]:
            module.eval()
            module.train = disabled_train
            for p in module.parameters():
                p.requires_grad = False
        return set(sd.keys()) - used

    @torch.no_grad()
    def load_controlnet_from_ckpt(self, sd: Dict[str, torch.Tensor]) -> None:
        self.controlnet.load_state_dict(sd, strict=True)

    @torch.no_grad()
    def load_controlnet_from_unet(self) -> Tuple[Set[str], Set[str]]:
        """model/cldm.py:83-105: copy the UNet encoder into the ControlNet, zero-padding the first conv."""
        unet_sd = self.unet.state_dict()
        scratch_sd = self.controlnet.state_dict()
        init_sd, with_new_zero, with_scratch = {}, set(), set()
        for key, this in scratch_sd.items():
            if key in unet_sd:
                target = unet_sd[key]
                if this.size() == target.size():
                    init_sd[key] = target.clone()
                else:
                    oc, _, h, w = this.size()
                    pad = torch.zeros((oc, this.size(1) - target.size(1), h, w), dtype=target.dtype,
                                      device=target.device)
                    init_sd[key] = torch.cat((target, pad), dim=1)
                    with_new_zero.add(key)
            else:
                init_sd[key] = this.clone()
                with_scratch.add(key)
        self.controlnet.load_state_dict(init_sd, strict=True)
        return with_new_zero, with_scratch

    # ----------------------------------------------------------------- engine plumbing
    def invalidate_engine(self) -> None:
        """Drop the packed bf16 weight copies (and captured graphs): call after writing parameters through `.data`
        (EMA updates, hand-written loaders) — such writes do not bump the version counters `engine()` watches."""
        self._engine = None
        self._engine_version = None
        self._engine32 = self._engine32_version = self._vae32 = self._vae32_version = None
        self._vae32e = self._vae32e_version = None
        self.vae.invalidate_engine()

    refresh_weights = invalidate_engine

    def engine(self):
        from .engine import CldmEngine, resolve_ops

        dev = next(self.unet.parameters()).device
        ver = (state_version(self.unet), state_version(self.controlnet), dev)
        if self._engine is None or self._engine_version != ver:
            if dev.type != "cuda" and getattr(resolve_ops(), "REQUIRES_CUDA", True):
                raise RuntimeError("edtr_b200 has no CPU path: move the model to a CUDA device first")
            self._engine = CldmEngine(self.unet.cfg, self.controlnet.cfg, self.unet.state_dict(),
                                      self.controlnet.state_dict(), dev)
            self._engine_version = ver
        return self._engine

    # ----------------------------------------------------------------- reference API
    @torch.no_grad()
    def forward_tiled(self, x_noisy, t, cond, tile_size: int, tile_stride: int) -> torch.Tensor:
        """Batched equivalent of the sampler's tiled wrapper (utils/sampler.py:288-303): all latent tiles of the step
        in one forward, gaussian-blended on the device.  With `self.tile_group` set (opt-in, same latent on every
        rank) the tiles are spread over the ranks of that group: one all-reduce of the partial blend per call."""
        from .parallel import check_same_across_ranks, tile_sharding

        rank, world, red = tile_sharding(self.tile_group)
        if world > 1 and self.tile_group_check:
            check_same_across_ranks(x_noisy, self.tile_group, "latent")
        eps = self.engine().forward_tiled(x_noisy.float().contiguous(), t.long().contiguous(),
                                          cond["c_img"].float().contiguous(), cond["c_txt"].float().contiguous(),
                                          tile_size, tile_stride, control_scales=self.control_scales, rank=rank,
                                          world=world, reduce_fn=red)
        return eps.to(x_noisy.dtype)

    @torch.no_grad()
    def vae_encode(self, image: torch.Tensor, sample: bool = True, tiled: bool = False, tile_size: int = -1):
        """model/cldm.py:107-134: posterior sample or mode, times the latent scale factor; tiled=True is the
        reference's VAEHook encode (pad 32, pooled GroupNorm statistics)."""
        if self.precision == "fp32":
            if tiled:
                raise NotImplementedError("the tiled VAE encode runs in the bf16 mode only")
            from .nets import DiagonalGaussianDistribution

            posterior = DiagonalGaussianDistribution(self._vae_encoder_f32().encode(image.float().contiguous()))
        else:
            posterior = (self.vae.encode_tiled(image, tile_size, tile_group=self.tile_group,
                                               tile_group_check=self.tile_group_check)
                         if tiled else self.vae.encode(image))
        z = posterior.sample() if sample else posterior.mode()
        return z * self.scale_factor

    @torch.no_grad()
    def vae_decode(self, z: torch.Tensor, tiled: bool = False, tile_size: int = -1) -> torch.Tensor:
        """model/cldm.py:136-156.  Returns fp32 NCHW in [-1, 1].  tiled=True is the reference's VAEHook decode
        (utils/tilevae/tilevae.py:307-579, pooled GroupNorm statistics); with `self.tile_group` set (opt-in, same
        latent on every rank) its tiles are spread over the ranks of that group (config C4)."""
        if self.precision == "fp32":
            if tiled:
                raise NotImplementedError("the tiled VAE decode runs in the bf16 mode only")
            return self._vae_decoder_f32().decode(z.float().contiguous(), float(self.scale_factor))
        eng = self.vae._decoder_engine()
        if not tiled:
            return eng.decode(z.float().contiguous(), float(self.scale_factor))
        from .parallel import check_same_across_ranks, tile_sharding

        rank, world, red = tile_sharding(self.tile_group)
        if world > 1 and self.tile_group_check:
            check_same_across_ranks(z, self.tile_group, "latent")
        return eng.decode_tiled(z.float().contiguous(), float(self.scale_factor), int(tile_size), rank=rank,
                                world=world, reduce_fn=red)

    @torch.no_grad()
    def prepare_condition(self, clean: torch.Tensor, prompt: List[str]) -> Dict[str, torch.Tensor]:
        """model/cldm.py:158-164: c_txt from the text tower (cached for the constant prompt), c_img = the posterior
        mode of the VAE encoder on the [-1, 1] image times the latent scale."""
        if prompt is None:
            prompt = [""] * clean.size(0)
        if self.clip is None:
            raise RuntimeError("this ControlLDM was built with clip_cfg=None: supply cond['c_txt'] yourself or construct "
                               "it with the reference's clip_cfg")
        return dict(c_txt=self.clip.encode(prompt), c_img=self.vae_encode(clean * 2 - 1, sample=False))

    @torch.no_grad()
    def forward(self, x_noisy, t, cond, woSD=False):
        """model/cldm.py:166-194."""
        if woSD:
            raise NotImplementedError("woSD=True (tail_block) is not on the EDTR path")
        c_txt, c_img = cond["c_txt"], cond["c_img"]
        if self.precision == "fp32":
            eps = self.engine_f32().forward(x_noisy.float().contiguous(), t.long().contiguous(), c_img.float().contiguous(),
                                            c_txt.float().contiguous(), control_scales=self.control_scales)
            return eps.to(x_noisy.dtype)
        eps = self.engine().forward(x_noisy.float().contiguous(), t.long().contiguous(), c_img.float().contiguous(),
                                    c_txt.float().contiguous(), control_scales=self.control_scales)
        return eps.to(x_noisy.dtype)
