"""Latent tiling helpers of the cldm-tiled path (utils/common.py:151-165, 351-427).

Same numerics as the reference: overlapped windows, gaussian weights, ``out / count``; tiles are
evaluated one by one exactly as the reference does (batching the tiles of a step is planned).
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np
import torch


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
