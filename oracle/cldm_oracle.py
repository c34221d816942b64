"""CPU oracle for the EDTR ControlLDM restore path — TEST INFRASTRUCTURE ONLY.

A plain PyTorch fp32 restatement (functional, no nn.Module) of the reference
algorithm, written from the reference's behaviour and citing the file:line each
function follows (paths relative to the reference tree, JaehaKim97/EDTR).  It is
imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs; the product path (``edtr_b200``) never
touches it.

Pinning: the reference ships no golden vectors.  The oracle is
pinned against the live reference itself: ``tests/golden/make_golden.py`` imports
/root/reference, loads the same synthetic weights into the reference modules and
records reference outputs as fixtures under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against those fixtures (max abs diff ~1e-6, fp32 CPU).

Weights live in flat dicts keyed exactly like the reference state-dicts
(``input_blocks.1.0.in_layers.0.weight`` ...), so the same dict loads into the
reference modules, this oracle and the product modules.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# The reference's default attention mode is SDPCrossAttention / SDPAttnBlock, i.e. F.scaled_dot_product_attention
# (model/config.py:35-37, model/attention.py:193, model/vae.py:298).  The restatement below spells the softmax out
# (bit-comparable on CPU, pinned by the fixtures); bench.py's GPU reference arm flips this switch so that the timed
# graph issues the same library kernels (flash / memory-efficient SDPA) as the reference does on a GPU.
USE_SDPA = False

# --------------------------------------------------------------------------- configs
# configs/det/voc2012/test/007_edtr-s4.yaml:21-87 (cldm.params), :95-100 (diffusion)
S4 = dict(
    unet=dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1),
              num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_head_channels=64, context_dim=1024),
    controlnet=dict(in_channels=4, hint_channels=4, model_channels=320, attention_resolutions=(4, 2, 1),
                    num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_head_channels=64, context_dim=1024),
    vae=dict(z_channels=4, embed_dim=4, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2),
    latent_scale_factor=0.18215,
    diffusion=dict(linear_start=0.00085, linear_end=0.0120, timesteps=1000),
    used_timesteps=(50, 100, 150, 200),
)
# Same topology rules at toy widths: CPU tests in seconds.
TINY = dict(
    unet=dict(in_channels=4, out_channels=4, model_channels=64, attention_resolutions=(2, 1),
              num_res_blocks=1, channel_mult=(1, 2), num_head_channels=64, context_dim=128),
    controlnet=dict(in_channels=4, hint_channels=4, model_channels=64, attention_resolutions=(2, 1),
                    num_res_blocks=1, channel_mult=(1, 2), num_head_channels=64, context_dim=128),
    vae=dict(z_channels=4, embed_dim=4, in_channels=3, out_ch=3, ch=64, ch_mult=(1, 2), num_res_blocks=1),
    latent_scale_factor=0.18215,
    diffusion=dict(linear_start=0.00085, linear_end=0.0120, timesteps=1000),
    used_timesteps=(50, 100, 150, 200),
)
# 4-level toy VAE (the tiled-VAE hook hard-codes the x8 latent->pixel factor, utils/tilevae/tilevae.py:389,226)
TINY_VAE8 = dict(z_channels=4, embed_dim=4, in_channels=3, out_ch=3, ch=64, ch_mult=(1, 1, 2, 2), num_res_blocks=1)


# ------------------------------------------------------------------- UNet topology
def unet_plan(cfg: dict, controlnet: bool = False):
    """Block list implied by the ctor loops (model/unet.py:494-672, model/controlnet.py:135-255).

    Returns (input_blocks, middle, output_blocks); each block is a list of layer
    tuples: ("conv_in", cin, cout) | ("res", cin, cout) | ("st", ch, heads) |
    ("down", ch) | ("up", ch).
    """
    mc = cfg["model_channels"]
    mult = cfg["channel_mult"]
    nrb = cfg["num_res_blocks"]
    attn_res = cfg["attention_resolutions"]
    hc = cfg["num_head_channels"]
    cin = cfg["in_channels"] + (cfg.get("hint_channels", 0) if controlnet else 0)
    inputs = [[("conv_in", cin, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers.append(("st", ch, ch // hc))
            inputs.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inputs.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch), ("st", ch, ch // hc), ("res", ch, ch)]
    outputs = []
    if not controlnet:
        for level, m in list(enumerate(mult))[::-1]:
            for i in range(nrb + 1):
                ich = chans.pop()
                layers = [("res", ch + ich, mc * m)]
                ch = mc * m
                if ds in attn_res:
                    layers.append(("st", ch, ch // hc))
                if level and i == nrb:
                    layers.append(("up", ch))
                    ds //= 2
                outputs.append(layers)
    return inputs, middle, outputs


def _res_shapes(p: str, cin: int, cout: int, emb: int) -> List[Tuple[str, Tuple[int, ...]]]:
    s = [(p + "in_layers.0.weight", (cin,)), (p + "in_layers.0.bias", (cin,)),
         (p + "in_layers.2.weight", (cout, cin, 3, 3)), (p + "in_layers.2.bias", (cout,)),
         (p + "emb_layers.1.weight", (cout, emb)), (p + "emb_layers.1.bias", (cout,)),
         (p + "out_layers.0.weight", (cout,)), (p + "out_layers.0.bias", (cout,)),
         (p + "out_layers.3.weight", (cout, cout, 3, 3)), (p + "out_layers.3.bias", (cout,))]
    if cin != cout:
        s += [(p + "skip_connection.weight", (cout, cin, 1, 1)), (p + "skip_connection.bias", (cout,))]
    return s


def _st_shapes(p: str, ch: int, ctx: int) -> List[Tuple[str, Tuple[int, ...]]]:
    t = p + "transformer_blocks.0."
    s = [(p + "norm.weight", (ch,)), (p + "norm.bias", (ch,)),
         (p + "proj_in.weight", (ch, ch)), (p + "proj_in.bias", (ch,))]
    for a, kd in (("attn1", ch), ("attn2", ctx)):
        s += [(t + a + ".to_q.weight", (ch, ch)), (t + a + ".to_k.weight", (ch, kd)),
              (t + a + ".to_v.weight", (ch, kd)), (t + a + ".to_out.0.weight", (ch, ch)),
              (t + a + ".to_out.0.bias", (ch,))]
    s += [(t + "ff.net.0.proj.weight", (8 * ch, ch)), (t + "ff.net.0.proj.bias", (8 * ch,)),
          (t + "ff.net.2.weight", (ch, 4 * ch)), (t + "ff.net.2.bias", (ch,))]
    for n in ("norm1", "norm2", "norm3"):
        s += [(t + n + ".weight", (ch,)), (t + n + ".bias", (ch,))]
    s += [(p + "proj_out.weight", (ch, ch)), (p + "proj_out.bias", (ch,))]
    return s


def _layer_shapes(p: str, layer, emb: int, ctx: int):
    kind = layer[0]
    if kind == "conv_in":
        return [(p + "weight", (layer[2], layer[1], 3, 3)), (p + "bias", (layer[2],))]
    if kind == "res":
        return _res_shapes(p, layer[1], layer[2], emb)
    if kind == "st":
        return _st_shapes(p, layer[1], ctx)
    if kind == "down":
        return [(p + "op.weight", (layer[1], layer[1], 3, 3)), (p + "op.bias", (layer[1],))]
    if kind == "up":
        return [(p + "conv.weight", (layer[1], layer[1], 3, 3)), (p + "conv.bias", (layer[1],))]
    raise ValueError(kind)


def unet_param_shapes(cfg: dict, controlnet: bool = False) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) list of ControlledUnetModel / ControlNet state-dicts."""
    mc = cfg["model_channels"]
    emb, ctx = 4 * mc, cfg["context_dim"]
    inputs, middle, outputs = unet_plan(cfg, controlnet)
    s = [("time_embed.0.weight", (emb, mc)), ("time_embed.0.bias", (emb,)),
         ("time_embed.2.weight", (emb, emb)), ("time_embed.2.bias", (emb,))]
    for j, layers in enumerate(inputs):
        for k, layer in enumerate(layers):
            s += _layer_shapes(f"input_blocks.{j}.{k}.", layer, emb, ctx)
    for k, layer in enumerate(middle):
        s += _layer_shapes(f"middle_block.{k}.", layer, emb, ctx)
    if controlnet:
        for j, layers in enumerate(inputs):
            ch = layers[0][2] if layers[0][0] in ("conv_in", "res") else layers[0][1]
            s += [(f"zero_convs.{j}.0.weight", (ch, ch, 1, 1)), (f"zero_convs.{j}.0.bias", (ch,))]
        ch = middle[-1][2]
        s += [("middle_block_out.0.weight", (ch, ch, 1, 1)), ("middle_block_out.0.bias", (ch,))]
    else:
        for j, layers in enumerate(outputs):
            for k, layer in enumerate(layers):
                s += _layer_shapes(f"output_blocks.{j}.{k}.", layer, emb, ctx)
        s += [("out.0.weight", (mc,)), ("out.0.bias", (mc,)),
              ("out.2.weight", (cfg["out_channels"], mc, 3, 3)), ("out.2.bias", (cfg["out_channels"],))]
    return s


def _vae_res_shapes(p: str, cin: int, cout: int):
    s = [(p + "norm1.weight", (cin,)), (p + "norm1.bias", (cin,)),
         (p + "conv1.weight", (cout, cin, 3, 3)), (p + "conv1.bias", (cout,)),
         (p + "norm2.weight", (cout,)), (p + "norm2.bias", (cout,)),
         (p + "conv2.weight", (cout, cout, 3, 3)), (p + "conv2.bias", (cout,))]
    if cin != cout:
        s += [(p + "nin_shortcut.weight", (cout, cin, 1, 1)), (p + "nin_shortcut.bias", (cout,))]
    return s


def vae_decoder_plan(cfg: dict):
    """[(level, [(cin, cout), ...], has_upsample)] from top level down (model/vae.py:493-515)."""
    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    block_in = ch * mult[-1]
    plan = []
    for level in reversed(range(len(mult))):
        block_out = ch * mult[level]
        blocks = []
        for _ in range(nrb + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        plan.append((level, blocks, level != 0))
    return plan, block_in


def vae_decoder_param_shapes(cfg: dict):
    """(key, shape) list of the AutoencoderKL keys the decode path reads (model/vae.py:449-525,689-690)."""
    z, ed = cfg["z_channels"], cfg["embed_dim"]
    top = cfg["ch"] * cfg["ch_mult"][-1]
    s = [("post_quant_conv.weight", (z, ed, 1, 1)), ("post_quant_conv.bias", (z,)),
         ("decoder.conv_in.weight", (top, z, 3, 3)), ("decoder.conv_in.bias", (top,))]
    s += _vae_res_shapes("decoder.mid.block_1.", top, top)
    s += [("decoder.mid.attn_1.norm.weight", (top,)), ("decoder.mid.attn_1.norm.bias", (top,))]
    for n in ("q", "k", "v", "proj_out"):
        s += [(f"decoder.mid.attn_1.{n}.weight", (top, top, 1, 1)), (f"decoder.mid.attn_1.{n}.bias", (top,))]
    s += _vae_res_shapes("decoder.mid.block_2.", top, top)
    plan, last = vae_decoder_plan(cfg)
    for level, blocks, has_up in plan:
        for i, (cin, cout) in enumerate(blocks):
            s += _vae_res_shapes(f"decoder.up.{level}.block.{i}.", cin, cout)
        if has_up:
            c = blocks[-1][1]
            s += [(f"decoder.up.{level}.upsample.conv.weight", (c, c, 3, 3)),
                  (f"decoder.up.{level}.upsample.conv.bias", (c,))]
    s += [("decoder.norm_out.weight", (last,)), ("decoder.norm_out.bias", (last,)),
          ("decoder.conv_out.weight", (cfg["out_ch"], last, 3, 3)), ("decoder.conv_out.bias", (cfg["out_ch"],))]
    return s


# ------------------------------------------------------------- synthetic weights
def synth_tensor(key: str, shape: Sequence[int], seed: int) -> Tensor:
    """Deterministic per-key init.  Every tensor is non-zero: the reference zero-initialises
    ResBlock out-convs, proj_out, zero-convs and unet.out (model/util.py:121-127), which would
    make eps == 0 and parity vacuous."""
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    u = torch.rand(tuple(shape), generator=g, dtype=torch.float32) * 2 - 1
    if key.endswith("weight") and len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        return u / math.sqrt(fan_in)
    if key.endswith("weight"):  # norm gain
        return 1.0 + 0.2 * u
    return 0.1 * u  # biases


def make_weights(shapes, seed: int) -> SD:
    return {k: synth_tensor(k, s, seed) for k, s in shapes}


def make_cldm_weights(cfg: dict, seed: int = 0) -> Dict[str, SD]:
    return dict(unet=make_weights(unet_param_shapes(cfg["unet"]), seed),
                controlnet=make_weights(unet_param_shapes(cfg["controlnet"], True), seed + 1),
                vae=make_weights(vae_decoder_param_shapes(cfg["vae"]), seed + 2))


def make_inputs(cfg: dict, batch: int, latent_hw: int = 64, seed: int = 1):
    """Seeded synthetic (x_T, cond, per-step noise): SURVEY.md §8(d), with the kernel-only
    c_img = 0.8 N(0,1) and x_T = q_sample(c_img, t=max(used), N(0,1)) (model/gaussian_diffusion.py:80-84)."""
    g = torch.Generator().manual_seed(seed)
    zc = cfg["unet"]["in_channels"]
    c_img = 0.8 * torch.randn(batch, zc, latent_hw, latent_hw, generator=g)
    c_txt = torch.randn(batch, 77, cfg["unet"]["context_dim"], generator=g)
    betas = make_betas(**cfg["diffusion"])
    ac = np.cumprod(1.0 - betas)
    t = max(cfg["used_timesteps"])
    n0 = torch.randn(c_img.shape, generator=g)
    x_T = float(np.sqrt(ac[t])) * c_img + float(np.sqrt(1.0 - ac[t])) * n0
    steps = len(cfg["used_timesteps"])
    noise = [torch.randn(c_img.shape, generator=g) for _ in range(steps)]
    return x_T, dict(c_txt=c_txt, c_img=c_img), noise


# ------------------------------------------------------------------------ leaf ops
def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """model/util.py:98-118 (repeat_only=False)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd: SD, p: str, x: Tensor, eps: float) -> Tensor:
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], eps)


def _conv(sd: SD, p: str, x: Tensor, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, sd[p + "weight"], sd[p + "bias"], stride=stride, padding=padding)


def _lin(sd: SD, p: str, x: Tensor, bias: bool = True) -> Tensor:
    return F.linear(x, sd[p + "weight"], sd[p + "bias"] if bias else None)


def resblock(sd: SD, p: str, x: Tensor, emb: Tensor) -> Tensor:
    """ResBlock._forward, use_scale_shift_norm=False, no up/down (model/unet.py:203-223)."""
    h = _conv(sd, p + "in_layers.2.", F.silu(_gn(sd, p + "in_layers.0.", x, 1e-5)))
    e = _lin(sd, p + "emb_layers.1.", F.silu(emb))
    h = h + e[:, :, None, None]
    h = _conv(sd, p + "out_layers.3.", F.silu(_gn(sd, p + "out_layers.0.", h, 1e-5)))
    if (p + "skip_connection.weight") in sd:
        x = _conv(sd, p + "skip_connection.", x, padding=0)
    return x + h


def attention(sd: SD, p: str, x: Tensor, ctx: Optional[Tensor], heads: int) -> Tensor:
    """SDPCrossAttention.forward (model/attention.py:176-203): softmax(q k^T / sqrt(d)) v per head."""
    ctx = x if ctx is None else ctx
    q, k, v = _lin(sd, p + "to_q.", x, False), _lin(sd, p + "to_k.", ctx, False), _lin(sd, p + "to_v.", ctx, False)
    b, n, c = q.shape
    d = c // heads

    def split(t):
        return t.view(b, t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    if USE_SDPA:
        o = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(b, n, c)
        return _lin(sd, p + "to_out.0.", o)
    w = torch.softmax(q @ k.transpose(-1, -2) * (d ** -0.5), dim=-1)
    o = (w @ v).permute(0, 2, 1, 3).reshape(b, n, c)
    return _lin(sd, p + "to_out.0.", o)


def transformer_block(sd: SD, p: str, x: Tensor, ctx: Tensor, heads: int) -> Tensor:
    """BasicTransformerBlock._forward (model/attention.py:230-234) with GEGLU feed-forward (:20-47)."""
    c = x.shape[-1]

    def ln(n, t):
        return F.layer_norm(t, (c,), sd[p + n + ".weight"], sd[p + n + ".bias"], 1e-5)

    x = attention(sd, p + "attn1.", ln("norm1", x), None, heads) + x
    x = attention(sd, p + "attn2.", ln("norm2", x), ctx, heads) + x
    h, gate = _lin(sd, p + "ff.net.0.proj.", ln("norm3", x)).chunk(2, dim=-1)
    x = _lin(sd, p + "ff.net.2.", h * F.gelu(gate)) + x
    return x


def spatial_transformer(sd: SD, p: str, x: Tensor, ctx: Tensor, heads: int) -> Tensor:
    """SpatialTransformer.forward, use_linear=True, depth 1 (model/attention.py:283-302)."""
    b, c, hh, ww = x.shape
    h = _gn(sd, p + "norm.", x, 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    h = _lin(sd, p + "proj_in.", h)
    h = transformer_block(sd, p + "transformer_blocks.0.", h, ctx, heads)
    h = _lin(sd, p + "proj_out.", h)
    return h.reshape(b, hh, ww, c).permute(0, 3, 1, 2) + x


def _run_layers(sd: SD, prefix: str, layers, h: Tensor, emb: Tensor, ctx: Tensor) -> Tensor:
    """TimestepEmbedSequential.forward (model/unet.py:40-48)."""
    for k, layer in enumerate(layers):
        p = f"{prefix}{k}."
        kind = layer[0]
        if kind == "conv_in":
            h = _conv(sd, p, h)
        elif kind == "res":
            h = resblock(sd, p, h, emb)
        elif kind == "st":
            h = spatial_transformer(sd, p, h, ctx, layer[2])
        elif kind == "down":  # Downsample, conv stride 2 pad 1 (model/unet.py:99-108)
            h = _conv(sd, p + "op.", h, stride=2)
        elif kind == "up":    # Upsample: nearest x2 then conv (model/unet.py:69-79)
            h = _conv(sd, p + "conv.", F.interpolate(h, scale_factor=2, mode="nearest"))
    return h


def _time_embed(sd: SD, t: Tensor, mc: int) -> Tensor:
    e = _lin(sd, "time_embed.0.", timestep_embedding(t, mc))
    return _lin(sd, "time_embed.2.", F.silu(e))


def controlnet_forward(sd: SD, cfg: dict, x: Tensor, hint: Tensor, t: Tensor, ctx: Tensor) -> List[Tensor]:
    """ControlNet.forward (model/controlnet.py:263-277): 13 zero-conv outputs for the s4 config."""
    inputs, middle, _ = unet_plan(cfg, controlnet=True)
    emb = _time_embed(sd, t, cfg["model_channels"])
    h = torch.cat((x, hint), dim=1)
    outs = []
    for j, layers in enumerate(inputs):
        h = _run_layers(sd, f"input_blocks.{j}.", layers, h, emb, ctx)
        outs.append(_conv(sd, f"zero_convs.{j}.0.", h, padding=0))
    h = _run_layers(sd, "middle_block.", middle, h, emb, ctx)
    outs.append(_conv(sd, "middle_block_out.0.", h, padding=0))
    return outs


def unet_forward(sd: SD, cfg: dict, x: Tensor, t: Tensor, ctx: Tensor, control: Optional[List[Tensor]]) -> Tensor:
    """ControlledUnetModel.forward, only_mid_control=False (model/controlnet.py:20-41)."""
    inputs, middle, outputs = unet_plan(cfg)
    emb = _time_embed(sd, t, cfg["model_channels"])
    control = list(control) if control is not None else None
    hs = []
    h = x
    for j, layers in enumerate(inputs):
        h = _run_layers(sd, f"input_blocks.{j}.", layers, h, emb, ctx)
        hs.append(h)
    h = _run_layers(sd, "middle_block.", middle, h, emb, ctx)
    if control is not None:
        h = h + control.pop()
    for j, layers in enumerate(outputs):
        skip = hs.pop()
        if control is not None:
            skip = skip + control.pop()
        h = _run_layers(sd, f"output_blocks.{j}.", layers, torch.cat([h, skip], dim=1), emb, ctx)
    return _conv(sd, "out.2.", F.silu(_gn(sd, "out.0.", h, 1e-5)))


def cldm_forward(w: Dict[str, SD], cfg: dict, x_noisy: Tensor, t: Tensor, cond: Dict[str, Tensor]) -> Tensor:
    """ControlLDM.forward, woSD=False, control_scales all 1.0 (model/cldm.py:166-194)."""
    control = controlnet_forward(w["controlnet"], cfg["controlnet"], x_noisy, cond["c_img"], t, cond["c_txt"])
    return unet_forward(w["unet"], cfg["unet"], x_noisy, t, cond["c_txt"], control)


# -------------------------------------------------------------------------- sampler
def make_betas(linear_start: float, linear_end: float, timesteps: int) -> np.ndarray:
    """make_beta_schedule("linear") (model/gaussian_diffusion.py:9-13)."""
    return np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2


def space_timesteps(num_timesteps: int, section_counts) -> set:
    """IDDPM respacing (utils/sampler.py:14-64)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == want:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


def make_schedule(betas: np.ndarray, num_steps: int, used_timesteps=None) -> Dict[str, np.ndarray]:
    """SpacedSampler.make_schedule (utils/sampler.py:85-133): fp64 tables, stored fp32."""
    ac_full = np.cumprod(1.0 - betas, axis=0)
    if used_timesteps is None:
        used_timesteps = space_timesteps(len(betas), str(num_steps))
    used = set(int(u) for u in used_timesteps)
    new_betas, last = [], 1.0
    for i, a in enumerate(ac_full):
        if i in used:
            new_betas.append(1 - a / last)
            last = a
    assert len(new_betas) == num_steps
    b = np.array(new_betas, dtype=np.float64)
    alphas = 1.0 - b
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    var = b * (1.0 - ac_prev) / (1.0 - ac)
    tables = dict(
        sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
        sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
        posterior_variance=var,
        posterior_mean_coef1=b * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    )
    out = {k: v.astype(np.float32) for k, v in tables.items()}
    out["timesteps"] = np.array(sorted(used), dtype=np.int32)
    return out


def p_sample_update(sched, x: Tensor, eps: Tensor, index: int, noise: Tensor) -> Tuple[Tensor, Tensor]:
    """Arithmetic of SpacedSampler.p_sample after the model call (utils/sampler.py:150-164,196-203)."""
    f = lambda k: torch.tensor(sched[k][index], dtype=torch.float32)
    pred_x0 = f("sqrt_recip_alphas_cumprod") * x - f("sqrt_recipm1_alphas_cumprod") * eps
    mean = f("posterior_mean_coef1") * pred_x0 + f("posterior_mean_coef2") * x
    nz = 1.0 if index != 0 else 0.0
    return mean + nz * torch.sqrt(f("posterior_variance")) * noise, pred_x0


def sample(w, cfg: dict, x_T: Tensor, cond, noise: List[Tensor], used_timesteps=None):
    """SpacedSampler.manual_sample_with_timesteps, cfg_scale=1, untiled (utils/sampler.py:267-323).
    `noise[i]` stands in for the i-th torch.randn_like draw.  Returns (x_0, [x_prev per step], [pred_x0])."""
    used = cfg["used_timesteps"] if used_timesteps is None else used_timesteps
    betas = make_betas(**cfg["diffusion"])
    sched = make_schedule(betas, len(used), used)
    ts = sched["timesteps"][::-1]
    total = len(ts)
    x = x_T
    xs, x0s = [], []
    for i, step in enumerate(ts):
        t = torch.full((x.shape[0],), int(step), dtype=torch.long, device=x.device)
        eps = cldm_forward(w, cfg, x, t, cond)
        x, pred_x0 = p_sample_update(sched, x, eps, total - i - 1, noise[i])
        xs.append(x)
        x0s.append(pred_x0)
    return x, xs, x0s


# ---------------------------------------------------------------------- VAE decoder
def _vae_resblock(sd: SD, p: str, x: Tensor) -> Tensor:
    """ResnetBlock.forward with temb=None (model/vae.py:103-124)."""
    h = _conv(sd, p + "conv1.", F.silu(_gn(sd, p + "norm1.", x, 1e-6)))
    h = _conv(sd, p + "conv2.", F.silu(_gn(sd, p + "norm2.", h, 1e-6)))
    if (p + "nin_shortcut.weight") in sd:
        x = _conv(sd, p + "nin_shortcut.", x, padding=0)
    return x + h


def _vae_attn(sd: SD, p: str, x: Tensor) -> Tensor:
    """SDPAttnBlock.forward: single head, d = C (model/vae.py:279-308)."""
    b, c, hh, ww = x.shape
    h = _gn(sd, p + "norm.", x, 1e-6)
    q, k, v = (_conv(sd, p + n + ".", h, padding=0).reshape(b, c, hh * ww).permute(0, 2, 1) for n in "qkv")
    if USE_SDPA:
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = o.permute(0, 2, 1).reshape(b, c, hh, ww)
        return x + _conv(sd, p + "proj_out.", o, padding=0)
    w = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1)
    o = (w @ v).permute(0, 2, 1).reshape(b, c, hh, ww)
    return x + _conv(sd, p + "proj_out.", o, padding=0)


def vae_decode(sd: SD, cfg: dict, z: Tensor, scale_factor: float) -> Tensor:
    """ControlLDM.vae_decode untiled -> AutoencoderKL.decode -> Decoder.forward
    (model/cldm.py:136-156, model/vae.py:731-734, :527-560)."""
    h = _conv(sd, "post_quant_conv.", z / scale_factor, padding=0)
    h = _conv(sd, "decoder.conv_in.", h)
    h = _vae_resblock(sd, "decoder.mid.block_1.", h)
    h = _vae_attn(sd, "decoder.mid.attn_1.", h)
    h = _vae_resblock(sd, "decoder.mid.block_2.", h)
    plan, _ = vae_decoder_plan(cfg)
    for level, blocks, has_up in plan:
        for i in range(len(blocks)):
            h = _vae_resblock(sd, f"decoder.up.{level}.block.{i}.", h)
        if has_up:
            h = _conv(sd, f"decoder.up.{level}.upsample.conv.", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "decoder.conv_out.", F.silu(_gn(sd, "decoder.norm_out.", h, 1e-6)))


# ---------------------------------------------------------------------- VAE encoder
def vae_encoder_plan(cfg: dict):
    """[(level, [(cin, cout), ...], has_downsample)] from the image side (model/vae.py:376-399)."""
    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    in_mult = (1,) + tuple(mult)
    plan = []
    for level in range(len(mult)):
        cin, cout = ch * in_mult[level], ch * mult[level]
        blocks = []
        for _ in range(nrb):
            blocks.append((cin, cout))
            cin = cout
        plan.append((level, blocks, level != len(mult) - 1))
    return plan, ch * mult[-1]


def vae_encoder_param_shapes(cfg: dict):
    """(key, shape) list of the AutoencoderKL keys the encode path reads (model/vae.py:326-420, :687-688)."""
    z, ed = cfg["z_channels"], cfg["embed_dim"]
    plan, top = vae_encoder_plan(cfg)
    s = [("encoder.conv_in.weight", (cfg["ch"], cfg["in_channels"], 3, 3)), ("encoder.conv_in.bias", (cfg["ch"],))]
    for level, blocks, has_down in plan:
        for i, (cin, cout) in enumerate(blocks):
            s += _vae_res_shapes(f"encoder.down.{level}.block.{i}.", cin, cout)
        if has_down:
            c = blocks[-1][1]
            s += [(f"encoder.down.{level}.downsample.conv.weight", (c, c, 3, 3)),
                  (f"encoder.down.{level}.downsample.conv.bias", (c,))]
    s += _vae_res_shapes("encoder.mid.block_1.", top, top)
    s += [("encoder.mid.attn_1.norm.weight", (top,)), ("encoder.mid.attn_1.norm.bias", (top,))]
    for n in ("q", "k", "v", "proj_out"):
        s += [(f"encoder.mid.attn_1.{n}.weight", (top, top, 1, 1)), (f"encoder.mid.attn_1.{n}.bias", (top,))]
    s += _vae_res_shapes("encoder.mid.block_2.", top, top)
    s += [("encoder.norm_out.weight", (top,)), ("encoder.norm_out.bias", (top,)),
          ("encoder.conv_out.weight", (2 * z, top, 3, 3)), ("encoder.conv_out.bias", (2 * z,)),
          ("quant_conv.weight", (2 * ed, 2 * z, 1, 1)), ("quant_conv.bias", (2 * ed,))]
    return s


def vae_encode_moments(sd: SD, cfg: dict, image: Tensor) -> Tensor:
    """AutoencoderKL.encode up to the posterior parameters: Encoder.forward + quant_conv
    (model/vae.py:421-446, :725-729).  Downsample = F.pad(x, (0,1,0,1)) + conv3x3 stride 2 pad 0 (:54-58)."""
    h = _conv(sd, "encoder.conv_in.", image)
    plan, _ = vae_encoder_plan(cfg)
    for level, blocks, has_down in plan:
        for i in range(len(blocks)):
            h = _vae_resblock(sd, f"encoder.down.{level}.block.{i}.", h)
        if has_down:
            h = _conv(sd, f"encoder.down.{level}.downsample.conv.", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = _vae_resblock(sd, "encoder.mid.block_1.", h)
    h = _vae_attn(sd, "encoder.mid.attn_1.", h)
    h = _vae_resblock(sd, "encoder.mid.block_2.", h)
    h = _conv(sd, "encoder.conv_out.", F.silu(_gn(sd, "encoder.norm_out.", h, 1e-6)))
    return _conv(sd, "quant_conv.", h, padding=0)


def vae_encode(sd: SD, cfg: dict, image: Tensor, scale_factor: float, noise: Optional[Tensor] = None) -> Tensor:
    """ControlLDM.vae_encode untiled (model/cldm.py:107-134): posterior.mode() (noise=None, `sample=False`) or
    posterior.sample() with the given N(0,1) draw, times the latent scale (model/distributions.py:24-65)."""
    mean, logvar = torch.chunk(vae_encode_moments(sd, cfg, image), 2, dim=1)
    if noise is None:
        return mean * scale_factor
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    return (mean + std * noise) * scale_factor


def q_sample(betas: np.ndarray, x_start: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """Diffusion.q_sample (model/gaussian_diffusion.py:80-84): sqrt(ac[t]) x0 + sqrt(1 - ac[t]) noise."""
    ac = np.cumprod(1.0 - betas, axis=0)
    a = torch.from_numpy(np.sqrt(ac)).float()[t].view(-1, 1, 1, 1)
    b = torch.from_numpy(np.sqrt(1.0 - ac)).float()[t].view(-1, 1, 1, 1)
    return a * x_start + b * noise


# ------------------------------------------------------------- tiled VAE decoder
VAE_TILE_PAD_DECODER = 11   # utils/tilevae/tilevae.py:315 (latent pixels)


def _best_tile_size(lowerbound: int, upperbound: int) -> int:
    """VAEHook.get_best_tile_size (utils/tilevae/tilevae.py:325-338)."""
    divider = 32
    while divider >= 2:
        rem = lowerbound % divider
        if rem == 0:
            return lowerbound
        cand = lowerbound - rem + divider
        if cand <= upperbound:
            return cand
        divider //= 2
    return lowerbound


def vae_split_tiles(h: int, w: int, tile_size: int, pad: int = VAE_TILE_PAD_DECODER, is_decoder: bool = True):
    """VAEHook.split_tiles (utils/tilevae/tilevae.py:340-399): ([x1,x2,y1,y2] input boxes grown by `pad`,
    output boxes in output pixels)."""
    nh = max(math.ceil((h - 2 * pad) / tile_size), 1)
    nw = max(math.ceil((w - 2 * pad) / tile_size), 1)
    th = _best_tile_size(math.ceil((h - 2 * pad) / nh), tile_size)
    tw = _best_tile_size(math.ceil((w - 2 * pad) / nw), tile_size)
    in_boxes, out_boxes = [], []
    for i in range(nh):
        for j in range(nw):
            ib = [pad + j * tw, min(pad + (j + 1) * tw, w), pad + i * th, min(pad + (i + 1) * th, h)]
            ob = [ib[0] if ib[0] > pad else 0, ib[1] if ib[1] < w - pad else w,
                  ib[2] if ib[2] > pad else 0, ib[3] if ib[3] < h - pad else h]
            out_boxes.append([v * 8 if is_decoder else v // 8 for v in ob])
            in_boxes.append([max(0, ib[0] - pad), min(w, ib[1] + pad), max(0, ib[2] - pad), min(h, ib[3] + pad)])
    return in_boxes, out_boxes


def _tile_group_stats(x: Tensor, groups: int = 32):
    """get_var_mean (utils/tilevae/tilevae.py:177-185): biased var / mean per (sample, group) of one tile."""
    b, c = x.shape[:2]
    xr = x.reshape(b * groups, -1)
    var, mean = torch.var_mean(xr, dim=1, unbiased=False)
    return var, mean


def _apply_group_stats(sd: SD, p: str, x: Tensor, mean: Tensor, var: Tensor, groups: int = 32) -> Tensor:
    """custom_group_norm (utils/tilevae/tilevae.py:188-215): fixed statistics, eps 1e-6, then the affine."""
    b, c = x.shape[:2]
    xr = x.reshape(b * groups, -1)
    y = (xr - mean[:, None]) / torch.sqrt(var[:, None] + 1e-6)
    y = y.reshape(x.shape)
    return y * sd[p + "weight"].view(1, -1, 1, 1) + sd[p + "bias"].view(1, -1, 1, 1)


def _vae_decoder_program(sd: SD, cfg: dict, x: Tensor):
    """Decoder.forward as the task queue of build_task_queue (utils/tilevae/tilevae.py:72-165) for ONE tile:
    a generator that yields (tensor, norm-prefix) at every `pre_norm` task and is resumed with the normalised
    tensor; returns the decoded tile."""

    def res(p, x):
        h = yield (x, p + "norm1.")
        h = _conv(sd, p + "conv1.", F.silu(h))
        h = yield (h, p + "norm2.")
        h = _conv(sd, p + "conv2.", F.silu(h))
        if (p + "nin_shortcut.weight") in sd:
            x = _conv(sd, p + "nin_shortcut.", x, padding=0)
        return x + h

    def attn(p, x):
        b, c, hh, ww = x.shape
        h = yield (x, p + "norm.")
        q, k, v = (_conv(sd, p + n + ".", h, padding=0).reshape(b, c, hh * ww).permute(0, 2, 1) for n in "qkv")
        w = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1)
        o = (w @ v).permute(0, 2, 1).reshape(b, c, hh, ww)
        return x + _conv(sd, p + "proj_out.", o, padding=0)

    h = _conv(sd, "decoder.conv_in.", x)
    h = yield from res("decoder.mid.block_1.", h)
    h = yield from attn("decoder.mid.attn_1.", h)
    h = yield from res("decoder.mid.block_2.", h)
    plan, _ = vae_decoder_plan(cfg)
    for level, blocks, has_up in plan:
        for i in range(len(blocks)):
            h = yield from res(f"decoder.up.{level}.block.{i}.", h)
        if has_up:
            h = _conv(sd, f"decoder.up.{level}.upsample.conv.", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    h = yield (h, "decoder.norm_out.")
    return _conv(sd, "decoder.conv_out.", F.silu(h))


def _run_tile_programs(sd: SD, progs):
    """The tile loop of VAEHook.vae_tile_forward (utils/tilevae/tilevae.py:496-571) without its CPU<->GPU
    shuffling: all tiles advance to their next `pre_norm`, the statistics are pooled (GroupNormParam.summary,
    :263-278: pixel-weighted average of per-tile means and per-tile variances), every tile is normalised with them."""
    pending = [next(g) for g in progs]
    done: List[Optional[Tensor]] = [None] * len(progs)
    while any(d is None for d in done):
        stats = [_tile_group_stats(x) for x, _ in pending]
        pix = torch.tensor([float(x.shape[2] * x.shape[3]) for x, _ in pending])
        wgt = pix / pix.max()
        wgt = wgt / wgt.sum()
        var = sum(wt * st[0] for wt, st in zip(wgt, stats))
        mean = sum(wt * st[1] for wt, st in zip(wgt, stats))
        nxt = []
        for i, (g, (x, p)) in enumerate(zip(progs, pending)):
            try:
                nxt.append(g.send(_apply_group_stats(sd, p, x, mean, var)))
            except StopIteration as fin:     # every tile runs the same queue, so all finish together
                done[i] = fin.value
                nxt.append(None)
        pending = nxt
    return done


def vae_decode_tiled(sd: SD, cfg: dict, z: Tensor, scale_factor: float, tile_size: int) -> Tensor:
    """ControlLDM.vae_decode(tiled=True) (model/cldm.py:142-156) -> VAEHook.__call__ / vae_tile_forward in the
    default (non-fast) mode (utils/tilevae/tilevae.py:317-323, :442-579): post_quant_conv on the whole latent,
    overlapping latent tiles, every GroupNorm uses the pixel-weighted average of the per-tile mean AND of the
    per-tile variance (GroupNormParam.summary, :263-278), attention is tile-local, valid regions are pasted."""
    pad = VAE_TILE_PAD_DECODER
    zq = _conv(sd, "post_quant_conv.", z / scale_factor, padding=0)
    n, _, hh, ww = zq.shape
    if max(hh, ww) <= pad * 2 + tile_size:          # "tiny and unnecessary to tile" (:319-321)
        return vae_decode(sd, cfg, z, scale_factor)
    in_boxes, out_boxes = vae_split_tiles(hh, ww, tile_size, pad, True)
    done = _run_tile_programs(sd, [_vae_decoder_program(sd, cfg, zq[:, :, b[2]:b[3], b[0]:b[1]]) for b in in_boxes])
    out = torch.zeros((n, done[0].shape[1], hh * 8, ww * 8), dtype=done[0].dtype)
    for tile, ib, ob in zip(done, in_boxes, out_boxes):
        # crop_valid_region (:218-229)
        m = [ob[i] - ib[i] * 8 for i in range(4)]
        out[:, :, ob[2]:ob[3], ob[0]:ob[1]] = tile[:, :, m[2]:tile.shape[2] + m[3], m[0]:tile.shape[3] + m[1]]
    return out


VAE_TILE_PAD_ENCODER = 32   # utils/tilevae/tilevae.py:315 (image pixels)


def _vae_encoder_program(sd: SD, cfg: dict, x: Tensor):
    """Encoder.forward as the task queue of build_task_queue (is_decoder=False) for ONE tile; yields at `pre_norm`."""

    def res(p, x):
        h = yield (x, p + "norm1.")
        h = _conv(sd, p + "conv1.", F.silu(h))
        h = yield (h, p + "norm2.")
        h = _conv(sd, p + "conv2.", F.silu(h))
        if (p + "nin_shortcut.weight") in sd:
            x = _conv(sd, p + "nin_shortcut.", x, padding=0)
        return x + h

    def attn(p, x):
        b, c, hh, ww = x.shape
        h = yield (x, p + "norm.")
        q, k, v = (_conv(sd, p + n + ".", h, padding=0).reshape(b, c, hh * ww).permute(0, 2, 1) for n in "qkv")
        w = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1)
        o = (w @ v).permute(0, 2, 1).reshape(b, c, hh, ww)
        return x + _conv(sd, p + "proj_out.", o, padding=0)

    h = _conv(sd, "encoder.conv_in.", x)
    plan, _ = vae_encoder_plan(cfg)
    for level, blocks, has_down in plan:
        for i in range(len(blocks)):
            h = yield from res(f"encoder.down.{level}.block.{i}.", h)
        if has_down:
            h = _conv(sd, f"encoder.down.{level}.downsample.conv.", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = yield from res("encoder.mid.block_1.", h)
    h = yield from attn("encoder.mid.attn_1.", h)
    h = yield from res("encoder.mid.block_2.", h)
    h = yield (h, "encoder.norm_out.")
    return _conv(sd, "encoder.conv_out.", F.silu(h))


def vae_encode_tiled(sd: SD, cfg: dict, image: Tensor, scale_factor: float, tile_size: int,
                     noise: Optional[Tensor] = None) -> Tensor:
    """ControlLDM.vae_encode(tiled=True) (model/cldm.py:114-126): VAEHook on the encoder (pad 32 image pixels,
    output boxes // 8), quant_conv on the assembled moments, then posterior mode / sample times the scale."""
    pad = VAE_TILE_PAD_ENCODER
    n, _, hh, ww = image.shape
    if max(hh, ww) <= pad * 2 + tile_size:
        return vae_encode(sd, cfg, image, scale_factor, noise)
    in_boxes, out_boxes = vae_split_tiles(hh, ww, tile_size, pad, False)
    done = _run_tile_programs(sd, [_vae_encoder_program(sd, cfg, image[:, :, b[2]:b[3], b[0]:b[1]]) for b in in_boxes])
    h = torch.zeros((n, done[0].shape[1], hh // 8, ww // 8), dtype=done[0].dtype)
    for tile, ib, ob in zip(done, in_boxes, out_boxes):
        m = [ob[i] - ib[i] // 8 for i in range(4)]        # crop_valid_region with is_decoder=False
        h[:, :, ob[2]:ob[3], ob[0]:ob[1]] = tile[:, :, m[2]:tile.shape[2] + m[3], m[0]:tile.shape[3] + m[1]]
    mean, logvar = torch.chunk(_conv(sd, "quant_conv.", h, padding=0), 2, dim=1)
    if noise is None:
        return mean * scale_factor
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    return (mean + std * noise) * scale_factor

def restore(w, cfg: dict, x_T: Tensor, cond, noise):
    """The unit of work of the headline metric: 4-step sample + VAE decode."""
    z, xs, _ = sample(w, cfg, x_T, cond, noise)
    return vae_decode(w["vae"], cfg["vae"], z, cfg["latent_scale_factor"]), xs


# ----------------------------------------------------------------- wavelet colour fix
def wavelet_blur(image: Tensor, radius: int) -> Tensor:
    """utils/common.py:99-118: depthwise 3x3 binomial kernel, dilation = radius, replicate padding."""
    c = image.shape[1]
    k = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]], dtype=image.dtype)
    k = k[None, None].repeat(c, 1, 1, 1)
    image = F.pad(image, (radius, radius, radius, radius), mode="replicate")
    return F.conv2d(image, k, groups=c, dilation=radius)


def wavelet_decomposition(image: Tensor, levels: int = 5):
    """utils/common.py:121-133."""
    high = torch.zeros_like(image)
    low = image
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i)
        high = high + (image - low)
        image = low
    return high, low


def wavelet_reconstruction(content: Tensor, style: Tensor) -> Tensor:
    """utils/common.py:136-147."""
    return wavelet_decomposition(content)[0] + wavelet_decomposition(style)[1]


# -------------------------------------------------------------------------- metrics
def max_rel_err(new: Tensor, ref: Tensor) -> float:
    """max|new - ref| / max|ref|."""
    return float((new.double() - ref.double()).abs().max() / ref.double().abs().max())


def psnr(a: Tensor, b: Tensor) -> float:
    """calculate_psnr_pt on [0,1] images, fp64 (utils/common.py:245-249)."""
    mse = torch.mean((a.double() - b.double()) ** 2)
    return float(10.0 * torch.log10(1.0 / (mse + 1e-8)))
