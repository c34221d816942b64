"""fp32 mode: thin wrappers over the ``edtr_f32_*`` entry points of the C-ABI library (include/edtr_b200.h).

Every tensor is fp32 and channels-last (``[B, H, W, C]`` / ``[rows, C]`` views with a uniform row stride, so a channel
slice of a wider buffer is a valid operand or destination).  The contractions accumulate fp32 products on the CUDA
cores: this is the accuracy mode of the path (BASELINE.json: per-step latent max-rel error <= 1e-4), used through
``ControlLDM.set_precision("fp32")``; the bf16 tensor-core kernels of ``ops.py`` are the throughput mode.  No CPU or
PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import lib as _lib
from .lib import EdtrF32Gemm
from .ops import ACT_NONE, ACT_SILU, _require_cuda, _stream, device_guard, rows_view, sampler_update  # noqa: F401

F32 = torch.float32
REQUIRES_CUDA = True
S_CHUNK_ELEMS = 1 << 28      # attention scores are materialised per chunk of samples: at most 1 GiB of fp32 at a time


def _rows(t: torch.Tensor) -> Tuple[int, int, int]:
    return rows_view(t, F32)


def _vecptr(t: Optional[torch.Tensor], n: int, name: str):
    if t is None:
        return None
    if t.dtype != F32 or not t.is_contiguous() or t.numel() < n:
        raise ValueError(f"{name} must be a contiguous fp32 tensor with >= {n} elements")
    return t.data_ptr()


def _epilogue(g: EdtrF32Gemm, M: int, N: int, out: torch.Tensor, *, bias, rowvec, rows_per_group, residual, act, alpha,
              nchw_hw: int) -> None:
    g.alpha = float(alpha)
    g.bias = _vecptr(bias, N, "bias")
    if rowvec is not None:
        if rowvec.dtype != F32 or rowvec.dim() != 2 or rowvec.stride(1) != 1 or rowvec.shape[1] < N:
            raise ValueError("rowvec must be a 2-D fp32 tensor with contiguous rows of >= N columns")
        if rows_per_group <= 0 or M % rows_per_group or rowvec.shape[0] < M // rows_per_group:
            raise ValueError("rowvec / rows_per_group do not match the row count")
        g.rowvec = rowvec.data_ptr()
        g.rowvec_ld = rowvec.stride(0)
        g.rows_per_group = rows_per_group
    if residual is not None:
        r_rows, r_cols, ldr = _rows(residual)
        if (r_rows, r_cols) != (M, N):
            raise ValueError(f"residual shape {tuple(residual.shape)} != ({M}, {N})")
        g.residual = residual.data_ptr()
        g.ldr = ldr
    if act not in (ACT_NONE, ACT_SILU):
        raise ValueError("fp32 epilogue supports ACT_NONE / ACT_SILU")
    g.act = act
    g.C = out.data_ptr()
    if nchw_hw:
        if out.dtype != F32 or not out.is_contiguous() or out.numel() != M * N or M % nchw_hw:
            raise ValueError("NCHW output must be a contiguous fp32 tensor of M*N elements")
        g.out_nchw, g.hw, g.ldc = 1, nchw_hw, 0
    else:
        o_rows, o_cols, ldc = _rows(out)
        if (o_rows, o_cols) != (M, N):
            raise ValueError(f"out shape {tuple(out.shape)} != ({M}, {N})")
        g.ldc = ldc


def _launch(g: EdtrF32Gemm) -> None:
    g.batch1 = g.batch1 or 1
    g.batch2 = g.batch2 or 1
    _lib.check(_lib.device_lib().edtr_f32_gemm(ctypes.byref(g), _stream()), "edtr_f32_gemm")


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, rowvec=None, rows_per_group=0, residual=None, act=ACT_NONE,
         alpha=1.0, out: Optional[torch.Tensor] = None, nchw_hw: int = 0) -> torch.Tensor:
    """``epilogue(a @ w.T)``: a [..., K] rows view, w [N, K], all fp32."""
    _require_cuda(a, w, bias, rowvec, residual, out)
    M, K, lda = _rows(a)
    N, Kw, ldw = _rows(w)
    if Kw != K:
        raise ValueError(f"K mismatch: a has {K}, w has {Kw}")
    if out is None:
        out = torch.empty((M // nchw_hw, N, nchw_hw) if nchw_hw else (M, N), dtype=F32, device=a.device)
    g = EdtrF32Gemm()
    g.A, g.W, g.M, g.N, g.K, g.lda, g.ldw = a.data_ptr(), w.data_ptr(), M, N, K, lda, ldw
    _epilogue(g, M, N, out, bias=bias, rowvec=rowvec, rows_per_group=rows_per_group, residual=residual, act=act,
              alpha=alpha, nchw_hw=nchw_hw)
    _launch(g)
    return out


def conv3x3(x: torch.Tensor, w: torch.Tensor, *, stride: int = 1, pad: Tuple[int, int] = (1, 1),
            out_hw: Optional[Tuple[int, int]] = None, up2x: bool = False, bias=None, rowvec=None, residual=None,
            act=ACT_NONE, alpha=1.0, out: Optional[torch.Tensor] = None, nchw: bool = False) -> torch.Tensor:
    """3x3 convolution as an implicit GEMM.  x [B, H, W, Cin] rows view, w [Cout, 9 * Cin] tap-major / channel-minor;
    `pad` = (top, left) zero padding, `out_hw` the output grid (default: 'same' for stride 1, ceil(H / 2) for stride 2
    with pad (1, 1), floor for the VAE's pad (0, 0) + right / bottom padding); up2x: convolve the nearest-2x up-sampled
    input (model/unet.py:69-79, model/vae.py:36-38) without materialising it."""
    _require_cuda(x, w, bias, rowvec, residual, out)
    if x.dim() != 4:
        raise ValueError("x must be [B, H, W, C]")
    B, H, W, Cin = x.shape
    _, _, ldx = _rows(x)
    Cout, Kw, ldw = _rows(w)
    if Kw != 9 * Cin:
        raise ValueError(f"weight must be [Cout, 9*Cin={9 * Cin}], got {tuple(w.shape)}")
    if up2x:
        Ho, Wo = 2 * H, 2 * W
    elif out_hw is not None:
        Ho, Wo = out_hw
    else:
        Ho, Wo = (H, W) if stride == 1 else ((H + 2 * pad[0] - 3) // stride + 1, (W + 2 * pad[1] - 3) // stride + 1)
    M = B * Ho * Wo
    if out is None:
        out = torch.empty((B, Cout, Ho * Wo) if nchw else (B, Ho, Wo, Cout), dtype=F32, device=x.device)
    g = EdtrF32Gemm()
    g.A, g.W, g.M, g.N, g.K, g.lda, g.ldw = x.data_ptr(), w.data_ptr(), M, Cout, 9 * Cin, ldx, ldw
    g.conv, g.H, g.W_in, g.Cin, g.Ho, g.Wo = 1, H, W, Cin, Ho, Wo
    g.conv_stride, g.pad_top, g.pad_left, g.up2x = stride, pad[0], pad[1], 1 if up2x else 0
    _epilogue(g, M, Cout, out, bias=bias, rowvec=rowvec, rows_per_group=Ho * Wo, residual=residual, act=act, alpha=alpha,
              nchw_hw=Ho * Wo if nchw else 0)
    _launch(g)
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T * scale) v per head, exact fp32.  q [B, Lq, heads*d], k / v [B, Lk, heads*d] rows views.  The scores
    are materialised (fp32, per chunk of samples), soft-maxed in place and multiplied by V with two batched GEMMs."""
    _require_cuda(q, k, v, out)
    B, Lq, C = q.shape
    Lk = k.shape[1]
    if C % heads or k.shape != (B, Lk, C) or v.shape != (B, Lk, C):
        raise ValueError("bad attention shapes")
    d = C // heads
    _, _, ldq = _rows(q)
    _, _, ldk = _rows(k)
    _, _, ldv = _rows(v)
    if out is None:
        out = torch.empty((B, Lq, C), dtype=F32, device=q.device)
    _, _, ldo = _rows(out)
    for t in (q, k, v, out):
        if t.stride(0) != t.shape[1] * t.stride(1):
            raise ValueError("attention operands must have uniformly strided rows across the batch")
    bc = max(1, min(B, S_CHUNK_ELEMS // max(1, heads * Lq * Lk)))
    S = torch.empty((bc, heads, Lq, Lk), dtype=F32, device=q.device)
    L = _lib.device_lib()
    for b0 in range(0, B, bc):
        nb = min(bc, B - b0)
        g = EdtrF32Gemm()                      # S = Q K^T
        g.A, g.W, g.C = q[b0].data_ptr(), k[b0].data_ptr(), S.data_ptr()
        g.M, g.N, g.K, g.lda, g.ldw, g.ldc = Lq, Lk, d, ldq, ldk, Lk
        g.batch1, g.batch2 = nb, heads
        g.a_stride1, g.a_stride2 = Lq * ldq, d
        g.w_stride1, g.w_stride2 = Lk * ldk, d
        g.c_stride1, g.c_stride2 = heads * Lq * Lk, Lq * Lk
        g.alpha = 1.0
        _lib.check(L.edtr_f32_gemm(ctypes.byref(g), _stream()), "edtr_f32_gemm")
        _lib.check(L.edtr_f32_softmax_rows(S.data_ptr(), Lk, nb * heads * Lq, Lk, float(scale), _stream()),
                   "edtr_f32_softmax_rows")
        g = EdtrF32Gemm()                      # O = P V   (V is [Lk, d] row-major: the [K, N] operand form)
        g.A, g.W, g.C = S.data_ptr(), v[b0].data_ptr(), out[b0].data_ptr()
        g.M, g.N, g.K, g.lda, g.ldw, g.ldc = Lq, d, Lk, Lk, ldv, ldo
        g.w_kn = 1
        g.batch1, g.batch2 = nb, heads
        g.a_stride1, g.a_stride2 = heads * Lq * Lk, Lq * Lk
        g.w_stride1, g.w_stride2 = Lk * ldv, d
        g.c_stride1, g.c_stride2 = Lq * ldo, d
        g.alpha = 1.0
        _lib.check(L.edtr_f32_gemm(ctypes.byref(g), _stream()), "edtr_f32_gemm")
    return out


def groupnorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm (+SiLU) over x [B, HW..., C] fp32 rows view; statistics in double precision."""
    _require_cuda(x, gamma, beta, out)
    B = x.shape[0]
    M, C, ldx = _rows(x)
    HW = M // B
    if out is None:
        out = torch.empty(x.shape, dtype=F32, device=x.device)
    Mo, Co, ldy = _rows(out)
    if (Mo, Co) != (M, C):
        raise ValueError("out shape mismatch")
    L = _lib.device_lib()
    scratch = torch.empty((int(L.edtr_f32_groupnorm_scratch_bytes(B, HW, C)) + 7) // 8, dtype=torch.float64, device=x.device)
    _lib.check(L.edtr_f32_groupnorm(x.data_ptr(), ldx, out.data_ptr(), ldy, B, HW, C, groups, _vecptr(gamma, C, "gamma"),
                                    _vecptr(beta, C, "beta"), float(eps), 1 if silu else 0, scratch.data_ptr(), _stream()),
               "edtr_f32_groupnorm")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, gamma, beta, out)
    M, C, ldx = _rows(x)
    if out is None:
        out = torch.empty(x.shape, dtype=F32, device=x.device)
    _, _, ldy = _rows(out)
    _lib.check(_lib.device_lib().edtr_f32_layernorm(x.data_ptr(), ldx, out.data_ptr(), ldy, M, C, _vecptr(gamma, C, "gamma"),
                                                    _vecptr(beta, C, "beta"), float(eps), _stream()), "edtr_f32_layernorm")
    return out


def geglu(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [..., 2N] -> x[..., :N] * gelu_erf(x[..., N:])."""
    _require_cuda(x, out)
    M, C2, ldx = _rows(x)
    N = C2 // 2
    if out is None:
        out = torch.empty(tuple(x.shape[:-1]) + (N,), dtype=F32, device=x.device)
    _, _, ldy = _rows(out)
    _lib.check(_lib.device_lib().edtr_f32_geglu(x.data_ptr(), ldx, out.data_ptr(), ldy, M, N, _stream()), "edtr_f32_geglu")
    return out


def silu(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, out)
    if x.dtype != F32 or not x.is_contiguous():
        raise ValueError("silu needs a contiguous fp32 tensor")
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.device_lib().edtr_f32_silu(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "edtr_f32_silu")
    return out


def nchw_to_nhwc(x: torch.Tensor, out: torch.Tensor, coff: int = 0, scale: float = 1.0) -> torch.Tensor:
    """x [B, C, H, W] fp32 (times `scale`) -> channels [coff, coff + C) of out [B, H, W, Cout] fp32."""
    _require_cuda(x, out)
    if x.dtype != F32 or not x.is_contiguous() or x.dim() != 4:
        raise ValueError("x must be a contiguous fp32 [B, C, H, W] tensor")
    B, C, H, W = x.shape
    _, Co, ldy = _rows(out)
    if out.numel() // Co != B * H * W or coff + C > Co:
        raise ValueError("out does not match x")
    _lib.check(_lib.device_lib().edtr_f32_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), ldy, B, C, H * W, coff, float(scale),
                                                       _stream()), "edtr_f32_nchw_to_nhwc")
    return out


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(t, out)
    if t.dtype != torch.int64 or not t.is_contiguous():
        raise ValueError("t must be a contiguous int64 tensor")
    B = t.numel()
    if out is None:
        out = torch.empty((B, dim), dtype=F32, device=t.device)
    _lib.check(_lib.device_lib().edtr_f32_timestep_embedding(t.data_ptr(), out.data_ptr(), B, dim, float(max_period),
                                                             _stream()), "edtr_f32_timestep_embedding")
    return out
