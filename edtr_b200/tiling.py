"""Tiling helpers: latent windows of the cldm-tiled path (utils/common.py:151-165, 351-427) and the tile
split of the tiled VAE (utils/tilevae/tilevae.py:325-399).  Host-side index arithmetic only.

Same numerics as the reference: overlapped windows, gaussian weights, ``out / count``.  The generic wrapper here
evaluates the tiles one by one as the reference does; an ``edtr_b200.ControlLDM`` takes ``CldmEngine.forward_tiled``
instead, which runs all windows of a step (this rank's share of them) as one batched forward and blends them on the
device (``edtr_tile_blend``).
"""
from __future__ import annotations

import math
from typing import Callable, List, Tuple

import numpy as np
import torch

VAE_TILE_PAD_DECODER = 11   # latent pixels (utils/tilevae/tilevae.py:315)
VAE_TILE_PAD_ENCODER = 32   # image pixels


def best_tile_size(lowerbound: int, upperbound: int) -> int:
    """VAEHook.get_best_tile_size (utils/tilevae/tilevae.py:325-338): smallest size >= lowerbound that is a
    multiple of the largest possible power of two (32 ... 2) and still <= upperbound."""
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


def vae_split_tiles(h: int, w: int, tile_size: int, pad: int, is_decoder: bool):
    """VAEHook.split_tiles (utils/tilevae/tilevae.py:340-399).  Boxes are [x1, x2, y1, y2]; returns the input
    boxes (tile + `pad` context, clipped to the tensor) and the output boxes (x8 for the decoder, //8 for the
    encoder) each tile's valid region is pasted into."""
    nh = max(math.ceil((h - 2 * pad) / tile_size), 1)
    nw = max(math.ceil((w - 2 * pad) / tile_size), 1)
    th = best_tile_size(math.ceil((h - 2 * pad) / nh), tile_size)
    tw = best_tile_size(math.ceil((w - 2 * pad) / nw), tile_size)
    in_boxes, out_boxes = [], []
    for i in range(nh):
        for j in range(nw):
            ib = [pad + j * tw, min(pad + (j + 1) * tw, w), pad + i * th, min(pad + (i + 1) * th, h)]
            ob = [ib[0] if ib[0] > pad else 0, ib[1] if ib[1] < w - pad else w,
                  ib[2] if ib[2] > pad else 0, ib[3] if ib[3] < h - pad else h]
            out_boxes.append([v * 8 if is_decoder else v // 8 for v in ob])
            in_boxes.append([max(0, ib[0] - pad), min(w, ib[1] + pad), max(0, ib[2] - pad), min(h, ib[3] + pad)])
    return in_boxes, out_boxes


def gaussian_weights(tile_width: int, tile_height: int) -> np.ndarray:
    """utils/common.py:151-165 (note the reference's asymmetric midpoints: (w-1)/2 for x, h/2 for y)."""
    var = 0.01
    mid_x = (tile_width - 1) / 2
    xs = np.arange(tile_width, dtype=np.float64)
    x_probs = np.exp(-(xs - mid_x) * (xs - mid_x) / (tile_width * tile_width) / (2 * var)) / np.sqrt(2 * np.pi * var)
    mid_y = tile_height / 2
    ys = np.arange(tile_height, dtype=np.float64)
    y_probs = np.exp(-(ys - mid_y) * (ys - mid_y) / (tile_height * tile_height) / (2 * var)) / np.sqrt(2 * np.pi * var)
    return np.outer(y_probs, x_probs)


def sliding_windows(h: int, w: int, tile_size: int, tile_stride: int) -> List[Tuple[int, int, int, int]]:
    """utils/common.py:351-364."""
    his = list(range(0, h - tile_size + 1, tile_stride))
    if (h - tile_size) % tile_stride != 0:
        his.append(h - tile_size)
    wis = list(range(0, w - tile_size + 1, tile_stride))
    if (w - tile_size) % tile_stride != 0:
        wis.append(w - tile_size)
    return [(hi, hi + tile_size, wi, wi + tile_size) for hi in his for wi in wis]


def make_tiled_fn(fn: Callable, size: int, stride: int) -> Callable:
    """utils/common.py:367-427 with scale 1 and gaussian weights (the only mode the sampler uses).

    ``fn(x_tile, t, cond, hi, hi_end, wi, wi_end)`` is the sampler's lambda (utils/sampler.py:290-301);
    it is called once per tile exactly as in the reference."""

    def tiled_fn(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        b, c, h, w = x.shape
        out = torch.zeros((b, c, h, w), dtype=x.dtype, device=x.device)
        count = torch.zeros_like(out, dtype=torch.float32)
        weights = torch.tensor(gaussian_weights(size, size)[None, None], dtype=x.dtype, device=x.device)
        for hi, hi_end, wi, wi_end in sliding_windows(h, w, size, stride):
            x_tile = x[..., hi:hi_end, wi:wi_end]
            kw = dict(kwargs)
            if len(args) or len(kwargs):
                kw.update(dict(hi=hi, hi_end=hi_end, wi=wi, wi_end=wi_end))
            out[..., hi:hi_end, wi:wi_end] += fn(x_tile, *args, **kw) * weights
            count[..., hi:hi_end, wi:wi_end] += weights
        return out / count

    return tiled_fn
