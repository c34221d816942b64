"""Wavelet colour fix applied to the decoded image (utils/common.py:99-147; ``wavelet_reconstruction((dec + 1) / 2,
pre_res)`` in main/det/test_edtr.py:135, demo.py:123).  Same name and argument order as the reference function."""
from __future__ import annotations

import torch

from .engine import resolve_ops


@torch.no_grad()
def wavelet_reconstruction(content_feat: torch.Tensor, style_feat: torch.Tensor) -> torch.Tensor:
    """content_feat's high frequencies + style_feat's low frequencies (5-level dilated binomial decomposition)."""
    return resolve_ops().wavelet_reconstruction(content_feat.float(), style_feat.float()).to(content_feat.dtype)
