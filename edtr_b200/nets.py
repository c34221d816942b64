"""Drop-in network classes: same names, constructor arguments, attributes, state-dict keys and call
signatures as the reference's ``ControlledUnetModel`` / ``ControlNet`` (model/controlnet.py:18-277,
model/unet.py:361-719) and ``AutoencoderKL`` (model/vae.py:681-743) — but they own parameters
only; every ``forward`` runs on the CUDA engine (``engine.py``).

Keeping the reference key layout is the drop-in contract: ``load_pretrained_sd``,
``load_controlnet_from_ckpt``, ``load_state_dict(strict=True)`` and checkpoints saved by the
reference's training scripts keep working unchanged (model/cldm.py:46-105).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import topology as T

_ZERO_SUFFIXES = ("out_layers.3.", "proj_out.")          # zero_module(...) call sites:
_ZERO_PREFIXES = ("zero_convs.", "middle_block_out.", "out.2.")  # model/unet.py:177-179,678; attention.py:280; controlnet.py:260


class _Node(nn.Module):
    """Anonymous container: reproduces the nesting of the reference modules in state-dict keys."""


def _attach(root: nn.Module, key: str, shape: Sequence[int]) -> None:
    parts = key.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, _Node())
        node = node._modules[p]
    node.register_parameter(parts[-1], nn.Parameter(torch.empty(tuple(shape), dtype=torch.float32)))


@torch.no_grad()
def _init_like_reference(module: nn.Module, zero_init: bool = True) -> None:
    """torch's default Conv2d/Linear init (U(+-1/sqrt(fan_in))), norm affine = (1, 0), and the
    reference's zero_module tensors (model/util.py:121-127)."""
    params = dict(module.named_parameters())
    for k, p in params.items():
        if k.endswith("weight") and p.dim() >= 2:
            fan_in = int(math.prod(p.shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            p.uniform_(-bound, bound)
            b = params.get(k[:-6] + "bias")
            if b is not None:
                b.uniform_(-bound, bound)
        elif k.endswith("weight"):
            p.fill_(1.0)
            b = params.get(k[:-6] + "bias")
            if b is not None:
                b.zero_()
    if zero_init:
        for k, p in params.items():
            stem = k.rsplit(".", 1)[0] + "."
            if stem.endswith(_ZERO_SUFFIXES) or stem.startswith(_ZERO_PREFIXES):
                # VAE / transformer to_out are not zero modules: only UNet-family keys reach here
                p.zero_()


def state_version(module: nn.Module) -> int:
    """Changes whenever a parameter is modified in place or replaced (engine re-pack trigger)."""
    v = 0
    for p in module.parameters():
        v = (v * 1000003 + p._version + (p.data_ptr() & 0xFFFFFFFF)) & 0xFFFFFFFFFFFF
    return v


class _UNetFamily(nn.Module):
    _controlnet = False

    def __init__(self, image_size, in_channels, model_channels, num_res_blocks, attention_resolutions,
                 out_channels=None, hint_channels=None, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True,
                 dims=2, num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1,
                 context_dim=None, n_embed=None, legacy=True, disable_self_attentions=None,
                 num_attention_blocks=None, disable_middle_self_attn=False, use_linear_in_transformer=False):
        super().__init__()
        if use_spatial_transformer:
            assert context_dim is not None, "context_dim is required with use_spatial_transformer"
        if context_dim is not None and not isinstance(context_dim, int):
            context_dim = list(context_dim)
            assert len(context_dim) == 1, "one context_dim per transformer depth (depth 1 only)"
            context_dim = int(context_dim[0])
        unsupported = []
        if not use_spatial_transformer or not use_linear_in_transformer:
            unsupported.append("use_spatial_transformer / use_linear_in_transformer must be True (SD-2.x layout)")
        if transformer_depth != 1:
            unsupported.append("transformer_depth != 1")
        if use_scale_shift_norm or resblock_updown or not conv_resample or dims != 2:
            unsupported.append("use_scale_shift_norm / resblock_updown / conv_resample=False / dims != 2")
        if num_classes is not None or n_embed is not None or use_fp16:
            unsupported.append("num_classes / n_embed / use_fp16")
        if disable_self_attentions is not None or num_attention_blocks is not None or disable_middle_self_attn:
            unsupported.append("disable_self_attentions / num_attention_blocks / disable_middle_self_attn")
        if num_head_channels == -1 or legacy:
            unsupported.append("num_head_channels must be set and legacy=False")
        if not isinstance(num_res_blocks, int):
            if len(set(num_res_blocks)) != 1 or len(num_res_blocks) != len(channel_mult):
                unsupported.append("per-level num_res_blocks")
            num_res_blocks = int(num_res_blocks[0])
        if dropout:
            unsupported.append("dropout != 0")
        if unsupported:
            raise NotImplementedError("edtr_b200 covers the EDTR/SD-2.1 ControlLDM configuration only: "
                                      + "; ".join(unsupported))
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.hint_channels = hint_channels
        self.num_res_blocks = len(channel_mult) * [num_res_blocks]
        self.attention_resolutions = attention_resolutions
        self.channel_mult = channel_mult
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float32
        self.num_head_channels = num_head_channels
        self.context_dim = context_dim
        self.cfg = dict(in_channels=in_channels, model_channels=model_channels, num_res_blocks=num_res_blocks,
                        attention_resolutions=tuple(attention_resolutions), channel_mult=tuple(channel_mult),
                        num_head_channels=num_head_channels, context_dim=context_dim)
        if self._controlnet:
            self.cfg["hint_channels"] = hint_channels
        else:
            self.cfg["out_channels"] = out_channels
        for key, shape in T.unet_param_shapes(self.cfg, self._controlnet):
            _attach(self, key, shape)
        _init_like_reference(self)


class ControlledUnetModel(_UNetFamily):
    """model/controlnet.py:18-41.  ``forward`` needs the owning ``ControlLDM`` (the UNet and the
    ControlNet are evaluated as one fused pass); standalone evaluation is routed through it."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, **kw):
        super().__init__(image_size, in_channels, model_channels, num_res_blocks, attention_resolutions,
                         out_channels=out_channels, **kw)

    def forward(self, x, timesteps=None, context=None, control=None, only_mid_control=False, **kwargs):
        raise NotImplementedError(
            "edtr_b200.ControlledUnetModel is evaluated together with its ControlNet: call ControlLDM.forward "
            "(model/cldm.py:166-194); a UNet-only evaluation with an external `control` list is not on the EDTR path")


class ControlNet(_UNetFamily):
    """model/controlnet.py:44-277."""
    _controlnet = True

    def __init__(self, image_size, in_channels, model_channels, hint_channels, num_res_blocks,
                 attention_resolutions, **kw):
        super().__init__(image_size, in_channels, model_channels, num_res_blocks, attention_resolutions,
                         hint_channels=hint_channels, **kw)

    def forward(self, x, hint, timesteps, context, **kwargs):
        raise NotImplementedError(
            "edtr_b200.ControlNet is evaluated inside ControlLDM.forward: its 13 outputs are accumulated into the "
            "UNet skip tensors by the zero-conv GEMM epilogues and never materialised")


class DiagonalGaussianDistribution:
    """model/distributions.py:24-65 (the members the EDTR scripts use): mean | logvar split, logvar clamped to
    [-30, 20], ``sample()`` = mean + std * randn (torch.randn on the host RNG, as the reference), ``mode()`` = mean."""

    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self) -> torch.Tensor:
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.parameters.device)

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKL(nn.Module):
    """model/vae.py:681-743 — parameter holder; ``encode`` / ``decode`` run on the CUDA engines."""

    def __init__(self, ddconfig: Dict, embed_dim: int, train_encoder: bool = False, train_decoder: bool = False):
        super().__init__()
        assert ddconfig["double_z"]
        self.ddconfig = dict(ddconfig)
        self.embed_dim = embed_dim
        self.train_encoder = train_encoder
        self.train_decoder = train_decoder
        for key, shape in T.vae_param_shapes(self.ddconfig, embed_dim):
            _attach(self, key, shape)
        _init_like_reference(self, zero_init=False)
        for k, p in self.named_parameters():
            trainable = (k.startswith("encoder.") and train_encoder) or (k.startswith("decoder.") and train_decoder)
            p.requires_grad = trainable
        self._engine = None
        self._engine_version = None

    def invalidate_engine(self) -> None:
        """Drop the packed weights / graphs of both engines (needed after `.data` writes, see ControlLDM)."""
        self._engine = None
        self._engine_version = None
        self._enc_engine = None
        self._enc_engine_version = None

    def _decoder_engine(self):
        from .engine import VaeDecoderEngine, resolve_ops

        ver = state_version(self)
        dev = next(self.parameters()).device
        if self._engine is None or self._engine_version != (ver, dev):
            if dev.type != "cuda" and getattr(resolve_ops(), "REQUIRES_CUDA", True):
                raise RuntimeError("edtr_b200 has no CPU path: move the model to a CUDA device first")
            self._engine = VaeDecoderEngine(self.ddconfig, self.embed_dim, self.state_dict(), dev)
            self._engine_version = (ver, dev)
        return self._engine

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """post_quant_conv + Decoder.forward (model/vae.py:731-734)."""
        return self._decoder_engine().decode(z.float().contiguous(), 1.0)

    def _encoder_engine(self):
        from .engine import VaeEncoderEngine, resolve_ops

        ver = state_version(self)
        dev = next(self.parameters()).device
        if getattr(self, "_enc_engine", None) is None or self._enc_engine_version != (ver, dev):
            if dev.type != "cuda" and getattr(resolve_ops(), "REQUIRES_CUDA", True):
                raise RuntimeError("edtr_b200 has no CPU path: move the model to a CUDA device first")
            self._enc_engine = VaeEncoderEngine(self.ddconfig, self.embed_dim, self.state_dict(), dev)
            self._enc_engine_version = (ver, dev)
        return self._enc_engine

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> "DiagonalGaussianDistribution":
        """Encoder.forward + quant_conv -> posterior (model/vae.py:725-729)."""
        return DiagonalGaussianDistribution(self._encoder_engine().encode(x.float().contiguous()))

    @torch.no_grad()
    def encode_tiled(self, x: torch.Tensor, tile_size: int, tile_group=None,
                     tile_group_check: bool = True) -> "DiagonalGaussianDistribution":
        """The `encoder` closure of ControlLDM.vae_encode(tiled=True) (model/cldm.py:114-126): VAEHook + quant_conv.
        `tile_group` (opt-in, see parallel.tile_sharding) spreads the tiles of one image — the SAME image on every
        rank — over the ranks of a process group; None (default) never communicates."""
        from .parallel import check_same_across_ranks, tile_sharding

        rank, world, red = tile_sharding(tile_group)
        if world > 1 and tile_group_check:
            check_same_across_ranks(x, tile_group, "image")
        return DiagonalGaussianDistribution(self._encoder_engine().encode_tiled(
            x.float().contiguous(), int(tile_size), rank=rank, world=world, reduce_fn=red))

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("AutoencoderKL.forward (encode + decode) is a training-time call, out of scope")
