"""CPU oracle for the SwinIR pre-restoration network — TEST INFRASTRUCTURE ONLY.

SURVEY.md §8(f) ranks SwinIR (``model/swinir.py:624-905``) as the next row after the ControlLDM restore path: it
is the dominant cost ahead of the VAE encoder in the real pipeline (``main/det/test_edtr.py:118``:
``val_pre_res_batch = swinir(val_lq_batch)``).  This module is the plain PyTorch fp32 restatement (functional, no
nn.Module) that a CUDA implementation of that row will be checked against; it is written from the reference's
behaviour, cites the file:line each function follows, and is imported only by ``tests/`` (nothing under
``edtr_b200/`` touches it — the product has no SwinIR path yet).

Pinning: the reference ships no golden vectors.  ``tests/golden/make_golden.py --swinir`` imports the
unmodified reference, loads the synthetic weights below into ``model.swinir.SwinIR`` and records its outputs in
``tests/golden/golden_swinir.npz``; ``tests/test_oracle.py`` checks this restatement against that fixture.

Weights live in a flat dict keyed exactly like the reference state-dict (parameters only; the reference's
``relative_position_index`` / ``attn_mask`` buffers are recomputed here from the geometry).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from .cldm_oracle import make_weights

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# configs/det/voc2012/test/007_edtr-s4.yaml:3-19 (swinir.params)
SWINIR_EDTR = dict(img_size=64, in_chans=3, embed_dim=180, depths=(6,) * 8, num_heads=(6,) * 8, window_size=8, mlp_ratio=2,
                   sf=8, img_range=1.0, num_feat=64)
# same topology rules at toy widths (two residual groups of two blocks: one plain and one shifted window block each;
# head dim 30 as in the EDTR configuration)
SWINIR_TINY = dict(img_size=16, in_chans=3, embed_dim=60, depths=(2, 2), num_heads=(2, 2), window_size=8, mlp_ratio=2,
                   sf=8, img_range=1.0, num_feat=64)   # num_feat is hard-coded in the reference (model/swinir.py:686)

RGB_MEAN = (0.4488, 0.4371, 0.4040)   # model/swinir.py:689-691


# ------------------------------------------------------------------------------------------ parameters
def swinir_param_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Parameter names and shapes of ``SwinIR(upsampler='nearest+conv', resi_connection='1conv', unshuffle=True,
    patch_norm=True, ape=False)`` (model/swinir.py:652-814), in state-dict order of the learnable tensors."""
    c, nf, ws = cfg["embed_dim"], cfg["num_feat"], cfg["window_size"]
    cin = cfg["in_chans"] * cfg["sf"] ** 2          # PixelUnshuffle(sf) in front of conv_first (:700-704)
    hid = int(c * cfg["mlp_ratio"])
    out: List[Tuple[str, Tuple[int, ...]]] = [("conv_first.1.weight", (c, cin, 3, 3)), ("conv_first.1.bias", (c,)),
                                              ("patch_embed.norm.weight", (c,)), ("patch_embed.norm.bias", (c,))]
    for i, (depth, heads) in enumerate(zip(cfg["depths"], cfg["num_heads"])):
        for j in range(depth):
            p = f"layers.{i}.residual_group.blocks.{j}."
            out += [(p + "norm1.weight", (c,)), (p + "norm1.bias", (c,)),
                    (p + "attn.relative_position_bias_table", ((2 * ws - 1) ** 2, heads)),
                    (p + "attn.qkv.weight", (3 * c, c)), (p + "attn.qkv.bias", (3 * c,)),
                    (p + "attn.proj.weight", (c, c)), (p + "attn.proj.bias", (c,)),
                    (p + "norm2.weight", (c,)), (p + "norm2.bias", (c,)),
                    (p + "mlp.fc1.weight", (hid, c)), (p + "mlp.fc1.bias", (hid,)),
                    (p + "mlp.fc2.weight", (c, hid)), (p + "mlp.fc2.bias", (c,))]
        out += [(f"layers.{i}.conv.weight", (c, c, 3, 3)), (f"layers.{i}.conv.bias", (c,))]
    out += [("norm.weight", (c,)), ("norm.bias", (c,)),
            ("conv_after_body.weight", (c, c, 3, 3)), ("conv_after_body.bias", (c,)),
            ("conv_before_upsample.0.weight", (nf, c, 3, 3)), ("conv_before_upsample.0.bias", (nf,))]
    n_up = {2: 1, 4: 2, 8: 3}[cfg["sf"]]           # conv_up1 (+ conv_up2 (+ conv_up3)) (:795-802)
    for k in range(1, n_up + 1):
        out += [(f"conv_up{k}.weight", (nf, nf, 3, 3)), (f"conv_up{k}.bias", (nf,))]
    out += [("conv_hr.weight", (nf, nf, 3, 3)), ("conv_hr.bias", (nf,)),
            ("conv_last.weight", (cfg["in_chans"], nf, 3, 3)), ("conv_last.bias", (cfg["in_chans"],))]
    return out


def make_swinir_weights(cfg: dict, seed: int = 7) -> SD:
    return make_weights(swinir_param_shapes(cfg), seed)


# ------------------------------------------------------------------------------------------ window helpers
def window_partition(x: Tensor, ws: int) -> Tensor:
    """[B, H, W, C] -> [B * nW, ws, ws, C] (model/swinir.py:37-49)."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)


def window_reverse(win: Tensor, ws: int, H: int, W: int) -> Tensor:
    """[B * nW, ws, ws, C] -> [B, H, W, C] (model/swinir.py:52-66)."""
    B = win.shape[0] // ((H // ws) * (W // ws))
    x = win.reshape(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def relative_position_index(ws: int) -> Tensor:
    """[ws*ws, ws*ws] index into the (2 ws - 1)^2 bias table (model/swinir.py:97-108)."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)   # 2, N
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()                        # N, N, 2
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shifted_window_mask(H: int, W: int, ws: int, shift: int) -> Tensor:
    """[nW, ws*ws, ws*ws] additive mask (0 / -100) of SW-MSA (model/swinir.py:222-243)."""
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).reshape(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


# ------------------------------------------------------------------------------------------ blocks
def _ln(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def _conv(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.conv2d(x, sd[p + "weight"], sd[p + "bias"], padding=1)


def window_attention(sd: SD, p: str, x: Tensor, heads: int, ws: int, mask) -> Tensor:
    """x [B*nW, N, C]; W-MSA / SW-MSA with the relative position bias (model/swinir.py:120-151)."""
    B_, N, C = x.shape
    d = C // heads
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).reshape(B_, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * d ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + "relative_position_bias_table"][relative_position_index(ws).view(-1)].view(N, N, heads)
    attn = attn + bias.permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = attn.softmax(-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def swin_block(sd: SD, p: str, x: Tensor, hw: Tuple[int, int], heads: int, ws: int, shift: int) -> Tensor:
    """x [B, H*W, C] (model/swinir.py:245-285).  `ws` / `shift` are already resolved against the CONSTRUCTOR
    resolution (block_geometry below), as in the reference, not against the actual feature map."""
    H, W = hw
    B, L, C = x.shape
    h = _ln(sd, p + "norm1.", x).reshape(B, H, W, C)
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    win = window_partition(h, ws).reshape(-1, ws * ws, C)
    mask = shifted_window_mask(H, W, ws, shift) if shift > 0 else None
    win = window_attention(sd, p + "attn.", win, heads, ws, mask).view(-1, ws, ws, C)
    h = window_reverse(win, ws, H, W)
    if shift > 0:
        h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
    x = x + h.reshape(B, L, C)
    m = F.linear(_ln(sd, p + "norm2.", x), sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    m = F.linear(F.gelu(m), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])          # nn.GELU: exact erf form
    return x + m


def block_geometry(cfg: dict, j: int) -> Tuple[int, int]:
    """(window, shift) of block j of a residual group: odd blocks are shifted by ws // 2; when the constructor
    resolution (img_size / patch_size 1) is not larger than the window the block uses one window of that size and no
    shift (model/swinir.py:199-202, :391-393)."""
    ws = cfg["window_size"]
    if cfg["img_size"] <= ws:
        return cfg["img_size"], 0
    return ws, (0 if j % 2 == 0 else ws // 2)


def rstb(sd: SD, cfg: dict, i: int, x: Tensor, hw: Tuple[int, int], depth: int, heads: int) -> Tensor:
    """Residual Swin transformer block: `depth` blocks (odd ones shifted by ws // 2), 3x3 conv, skip
    (model/swinir.py:391-396, :487-488)."""
    H, W = hw
    B, L, C = x.shape
    h = x
    for j in range(depth):
        ws, shift = block_geometry(cfg, j)
        h = swin_block(sd, f"layers.{i}.residual_group.blocks.{j}.", h, hw, heads, ws, shift)
    h = _conv(sd, f"layers.{i}.conv.", h.transpose(1, 2).reshape(B, C, H, W))
    return h.flatten(2).transpose(1, 2) + x


def swinir_forward(sd: SD, cfg: dict, x: Tensor) -> Tensor:
    """LQ image [B, 3, H, W] in [0, 1] -> pre-restored image [B, 3, H, W] (model/swinir.py:856-894, the
    'nearest+conv' branch with PixelUnshuffle in front, so the output has the input's size)."""
    ws, sf = cfg["window_size"], cfg["sf"]
    B, _, H, W = x.shape
    pad_h, pad_w = (ws - H % ws) % ws, (ws - W % ws) % ws
    if pad_h or pad_w:
        x = F.pad(x, (0, pad_w, 0, pad_h), "reflect")                                   # check_image_size (:834-839)
    mean = torch.tensor(RGB_MEAN, dtype=x.dtype).view(1, 3, 1, 1) if cfg["in_chans"] == 3 else torch.zeros(1, 1, 1, 1)
    x = (x - mean) * cfg["img_range"]
    x = _conv(sd, "conv_first.1.", F.pixel_unshuffle(x, sf))
    hw = (x.shape[2], x.shape[3])
    h = _ln(sd, "patch_embed.norm.", x.flatten(2).transpose(1, 2))                      # PatchEmbed (:530-534)
    for i, (depth, heads) in enumerate(zip(cfg["depths"], cfg["num_heads"])):
        h = rstb(sd, cfg, i, h, hw, depth, heads)
    h = _ln(sd, "norm.", h).transpose(1, 2).reshape(B, -1, hw[0], hw[1])                # forward_features (:841-854)
    x = _conv(sd, "conv_after_body.", h) + x
    x = F.leaky_relu(_conv(sd, "conv_before_upsample.0.", x), 0.01)                     # nn.LeakyReLU() default slope
    n_up = {2: 1, 4: 2, 8: 3}[sf]
    for k in range(1, n_up + 1):
        x = F.leaky_relu(_conv(sd, f"conv_up{k}.", F.interpolate(x, scale_factor=2, mode="nearest")), 0.2)
    x = _conv(sd, "conv_last.", F.leaky_relu(_conv(sd, "conv_hr.", x), 0.2))
    x = x / cfg["img_range"] + mean
    return x[:, :, :H * sf, :W * sf]


def swinir_gflops(cfg: dict, H: int, W: int) -> float:
    """2*MAC of one forward on an [H, W] image (for the roofline of the next row)."""
    c, nf, ws, sf = cfg["embed_dim"], cfg["num_feat"], cfg["window_size"], cfg["sf"]
    h, w = H // sf, W // sf
    L = h * w
    hid = int(c * cfg["mlp_ratio"])
    fl = 2.0 * L * c * 9 * cfg["in_chans"] * sf * sf
    for depth in cfg["depths"]:
        per_block = 2.0 * L * c * 3 * c + 2.0 * L * c * c + 4.0 * L * ws * ws * c + 4.0 * L * c * hid
        fl += depth * per_block + 2.0 * L * c * c * 9
    fl += 2.0 * L * c * c * 9 + 2.0 * L * nf * c * 9
    n_up = {2: 1, 4: 2, 8: 3}[sf]
    for k in range(1, n_up + 1):
        fl += 2.0 * L * 4 ** k * nf * nf * 9
    fl += 2.0 * L * 4 ** n_up * (nf * nf * 9 + cfg["in_chans"] * nf * 9)
    return fl / 1e9
