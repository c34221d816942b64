"""Block topology and state-dict layout of the ControlLDM networks.

The reference builds its UNet / ControlNet / VAE decoder with constructor loops
(model/unet.py:494-672, model/controlnet.py:135-255, model/vae.py:449-525); the
state-dict keys those loops produce are the contract the SD-2.1 checkpoints and the
reference's weight loaders rely on (model/cldm.py:46-105).  This module derives the
same block lists and (key, shape) enumeration from the constructor arguments, so
that the parameter holders in ``nets.py`` stay loadable by the reference loaders
and the engine knows what to launch for each block.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

Layer = Tuple  # ("conv_in", cin, cout) | ("res", cin, cout) | ("st", ch, heads) | ("down", ch) | ("up", ch)
Shapes = List[Tuple[str, Tuple[int, ...]]]


def unet_blocks(cfg: Dict, controlnet: bool = False):
    """(input_blocks, middle_block, output_blocks) as lists of layer tuples."""
    mc = cfg["model_channels"]
    mult = tuple(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    attn = set(cfg["attention_resolutions"])
    hc = cfg["num_head_channels"]
    cin = cfg["in_channels"] + (cfg.get("hint_channels", 0) if controlnet else 0)
    inputs: List[List[Layer]] = [[("conv_in", cin, mc)]]
    widths = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            block: List[Layer] = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn:
                block.append(("st", ch, ch // hc))
            inputs.append(block)
            widths.append(ch)
        if level + 1 < len(mult):
            inputs.append([("down", ch)])
            widths.append(ch)
            ds *= 2
    middle: List[Layer] = [("res", ch, ch), ("st", ch, ch // hc), ("res", ch, ch)]
    outputs: List[List[Layer]] = []
    if not controlnet:
        for level in reversed(range(len(mult))):
            for i in range(nrb + 1):
                skip = widths.pop()
                block = [("res", ch + skip, mc * mult[level])]
                ch = mc * mult[level]
                if ds in attn:
                    block.append(("st", ch, ch // hc))
                if level > 0 and i == nrb:
                    block.append(("up", ch))
                    ds //= 2
                outputs.append(block)
    return inputs, middle, outputs


def block_out_channels(block: Sequence[Layer]) -> int:
    first = block[0]
    return first[2] if first[0] in ("conv_in", "res") else first[1]


def _res(p: str, cin: int, cout: int, emb: int) -> Shapes:
    s = [(p + "in_layers.0.weight", (cin,)), (p + "in_layers.0.bias", (cin,)),
         (p + "in_layers.2.weight", (cout, cin, 3, 3)), (p + "in_layers.2.bias", (cout,)),
         (p + "emb_layers.1.weight", (cout, emb)), (p + "emb_layers.1.bias", (cout,)),
         (p + "out_layers.0.weight", (cout,)), (p + "out_layers.0.bias", (cout,)),
         (p + "out_layers.3.weight", (cout, cout, 3, 3)), (p + "out_layers.3.bias", (cout,))]
    if cin != cout:
        s += [(p + "skip_connection.weight", (cout, cin, 1, 1)), (p + "skip_connection.bias", (cout,))]
    return s


def _st(p: str, ch: int, ctx: int) -> Shapes:
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


def _layer(p: str, layer: Layer, emb: int, ctx: int) -> Shapes:
    kind = layer[0]
    if kind == "conv_in":
        return [(p + "weight", (layer[2], layer[1], 3, 3)), (p + "bias", (layer[2],))]
    if kind == "res":
        return _res(p, layer[1], layer[2], emb)
    if kind == "st":
        return _st(p, layer[1], ctx)
    if kind == "down":
        return [(p + "op.weight", (layer[1], layer[1], 3, 3)), (p + "op.bias", (layer[1],))]
    if kind == "up":
        return [(p + "conv.weight", (layer[1], layer[1], 3, 3)), (p + "conv.bias", (layer[1],))]
    raise ValueError(kind)


def unet_param_shapes(cfg: Dict, controlnet: bool = False) -> Shapes:
    """State-dict (key, shape) list of ControlledUnetModel (or ControlNet)."""
    mc = cfg["model_channels"]
    emb, ctx = 4 * mc, cfg["context_dim"]
    inputs, middle, outputs = unet_blocks(cfg, controlnet)
    s: Shapes = [("time_embed.0.weight", (emb, mc)), ("time_embed.0.bias", (emb,)),
                 ("time_embed.2.weight", (emb, emb)), ("time_embed.2.bias", (emb,))]
    for j, block in enumerate(inputs):
        for k, layer in enumerate(block):
            s += _layer(f"input_blocks.{j}.{k}.", layer, emb, ctx)
    for k, layer in enumerate(middle):
        s += _layer(f"middle_block.{k}.", layer, emb, ctx)
    if controlnet:
        for j, block in enumerate(inputs):
            ch = block_out_channels(block)
            s += [(f"zero_convs.{j}.0.weight", (ch, ch, 1, 1)), (f"zero_convs.{j}.0.bias", (ch,))]
        ch = middle[-1][2]
        s += [("middle_block_out.0.weight", (ch, ch, 1, 1)), ("middle_block_out.0.bias", (ch,))]
    else:
        for j, block in enumerate(outputs):
            for k, layer in enumerate(block):
                s += _layer(f"output_blocks.{j}.{k}.", layer, emb, ctx)
        s += [("out.0.weight", (mc,)), ("out.0.bias", (mc,)),
              ("out.2.weight", (cfg["out_channels"], mc, 3, 3)), ("out.2.bias", (cfg["out_channels"],))]
    return s


def _vae_res(p: str, cin: int, cout: int) -> Shapes:
    s = [(p + "norm1.weight", (cin,)), (p + "norm1.bias", (cin,)),
         (p + "conv1.weight", (cout, cin, 3, 3)), (p + "conv1.bias", (cout,)),
         (p + "norm2.weight", (cout,)), (p + "norm2.bias", (cout,)),
         (p + "conv2.weight", (cout, cout, 3, 3)), (p + "conv2.bias", (cout,))]
    if cin != cout:
        s += [(p + "nin_shortcut.weight", (cout, cin, 1, 1)), (p + "nin_shortcut.bias", (cout,))]
    return s


def _vae_attn(p: str, ch: int) -> Shapes:
    s = [(p + "norm.weight", (ch,)), (p + "norm.bias", (ch,))]
    for n in ("q", "k", "v", "proj_out"):
        s += [(p + n + ".weight", (ch, ch, 1, 1)), (p + n + ".bias", (ch,))]
    return s


def vae_decoder_levels(dd: Dict):
    """[(level, [(cin, cout), ...], has_upsample)] top level first, and the final width."""
    ch, mult, nrb = dd["ch"], tuple(dd["ch_mult"]), dd["num_res_blocks"]
    width = ch * mult[-1]
    levels = []
    for level in reversed(range(len(mult))):
        out = ch * mult[level]
        blocks = []
        for _ in range(nrb + 1):
            blocks.append((width, out))
            width = out
        levels.append((level, blocks, level != 0))
    return levels, width


def vae_encoder_levels(dd: Dict):
    """[(level, [(cin, cout), ...], has_downsample)] from the input side (model/vae.py:376-399)."""
    ch, mult, nrb = dd["ch"], tuple(dd["ch_mult"]), dd["num_res_blocks"]
    in_mult = (1,) + mult
    levels = []
    for level in range(len(mult)):
        cin, cout = ch * in_mult[level], ch * mult[level]
        blocks = []
        for _ in range(nrb):
            blocks.append((cin, cout))
            cin = cout
        levels.append((level, blocks, level != len(mult) - 1))
    return levels, ch * mult[-1]


def vae_param_shapes(dd: Dict, embed_dim: int) -> Shapes:
    """State-dict (key, shape) list of AutoencoderKL (encoder, decoder, quant convs)."""
    z = dd["z_channels"]
    zz = 2 * z if dd.get("double_z", True) else z
    s: Shapes = []
    # encoder (model/vae.py:326-420)
    enc_levels, top = vae_encoder_levels(dd)
    s += [("encoder.conv_in.weight", (dd["ch"], dd["in_channels"], 3, 3)), ("encoder.conv_in.bias", (dd["ch"],))]
    for level, blocks, has_down in enc_levels:
        for i, (cin, cout) in enumerate(blocks):
            s += _vae_res(f"encoder.down.{level}.block.{i}.", cin, cout)
        if has_down:
            c = blocks[-1][1]
            s += [(f"encoder.down.{level}.downsample.conv.weight", (c, c, 3, 3)),
                  (f"encoder.down.{level}.downsample.conv.bias", (c,))]
    s += _vae_res("encoder.mid.block_1.", top, top)
    s += _vae_attn("encoder.mid.attn_1.", top)
    s += _vae_res("encoder.mid.block_2.", top, top)
    s += [("encoder.norm_out.weight", (top,)), ("encoder.norm_out.bias", (top,)),
          ("encoder.conv_out.weight", (zz, top, 3, 3)), ("encoder.conv_out.bias", (zz,))]
    s += vae_decoder_param_shapes(dd, embed_dim, with_post_quant=False)
    s += [("quant_conv.weight", (2 * embed_dim, zz, 1, 1)), ("quant_conv.bias", (2 * embed_dim,)),
          ("post_quant_conv.weight", (z, embed_dim, 1, 1)), ("post_quant_conv.bias", (z,))]
    return s


def vae_decoder_param_shapes(dd: Dict, embed_dim: int, with_post_quant: bool = True) -> Shapes:
    """Keys the decode path reads (model/vae.py:449-525, :689-690)."""
    z = dd["z_channels"]
    top = dd["ch"] * tuple(dd["ch_mult"])[-1]
    s: Shapes = []
    if with_post_quant:
        s += [("post_quant_conv.weight", (z, embed_dim, 1, 1)), ("post_quant_conv.bias", (z,))]
    s += [("decoder.conv_in.weight", (top, z, 3, 3)), ("decoder.conv_in.bias", (top,))]
    s += _vae_res("decoder.mid.block_1.", top, top)
    s += _vae_attn("decoder.mid.attn_1.", top)
    s += _vae_res("decoder.mid.block_2.", top, top)
    levels, last = vae_decoder_levels(dd)
    for level, blocks, has_up in levels:
        for i, (cin, cout) in enumerate(blocks):
            s += _vae_res(f"decoder.up.{level}.block.{i}.", cin, cout)
        if has_up:
            c = blocks[-1][1]
            s += [(f"decoder.up.{level}.upsample.conv.weight", (c, c, 3, 3)),
                  (f"decoder.up.{level}.upsample.conv.bias", (c,))]
    s += [("decoder.norm_out.weight", (last,)), ("decoder.norm_out.bias", (last,)),
          ("decoder.conv_out.weight", (dd["out_ch"], last, 3, 3)), ("decoder.conv_out.bias", (dd["out_ch"],))]
    return s
