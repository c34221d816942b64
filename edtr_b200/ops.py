"""Tensor-level wrappers over the C-ABI (``include/edtr_b200.h``).

PyTorch supplies device memory and the current stream only; every function here
validates its arguments (``ValueError`` for shape / dtype / layout problems, as the
reference asserts on shapes — model/unet.py:70,107) and then launches the CUDA
kernels through ctypes.  There is no CPU path: CPU tensors raise ``RuntimeError``.

Activations are bf16 channels-last.  A "rows view" is any tensor whose last dim is
contiguous and whose leading dims collapse to rows with one uniform stride, e.g. a
channel slice ``buf[..., :320]`` of a wider concat buffer.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import lib as _lib
from .lib import EdtrEpilogue

ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_GELU, ACT_LRELU_02, ACT_LRELU_001 = 0, 1, 2, 3, 4, 5
OUT_BF16, OUT_F32, OUT_NCHW_F32, OUT_NCHW_BF16 = 0, 1, 2, 3

BF16 = torch.bfloat16
REQUIRES_CUDA = True  # there is no CPU implementation of any op


def geglu_tile_n() -> int:
    """N-tile width of the GEGLU GEMM (plan-time weight interleaving); no device needed."""
    return int(_lib.load().edtr_gemm_tile_n(128, 1024, 64, ACT_GEGLU))


def use_workspace(index: int) -> None:
    """Split-K scratch slot for launches issued by this host thread from now on (0: main stream, 1: side stream);
    the buffer travels with every call (EdtrEpilogue.workspace) — the library keeps no state."""
    _lib.use_workspace(index)


def set_gemm_max_clusters(n: int) -> None:
    """Limit later GEMM / convolution launches of this host thread to n CTA pairs (74 = whole GPU) so that launches on
    two streams can run side by side (passed per call as EdtrEpilogue.max_clusters)."""
    _lib.set_gemm_max_clusters(n)


def gemm_workspace_size(M: int, N: int, K: int) -> int:
    """Bytes of split-K scratch with which the planner may pick any split factor for an M x N x K problem."""
    return int(_lib.load().edtr_gemm_workspace_size(M, N, K))


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _NullGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL_GUARD = _NullGuard()


def device_guard(device):
    """Context that makes `device` (a tensor's or an engine's) the current CUDA device for the launches inside: kernels
    are issued on that device's current stream with that device's scratch (a model moved to cuda:1 works while the
    current device is cuda:0, as it does in the reference).  Free when the device is already current."""
    if isinstance(device, torch.Tensor):
        device = device.device
    device = torch.device(device)
    if device.type != "cuda":
        return _NULL_GUARD
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx == torch.cuda.current_device():
        return _NULL_GUARD
    return torch.cuda.device(idx)


def row_stats_parts(M: int, N: int, K: int, device=None) -> int:
    """(sum, sum of squares) pairs per row that `gemm(a [M, K], w [N, K], ..., row_stats=...)` writes with the current
    thread's split-K scratch / CTA-pair share: one per column tile and epilogue warp group of the launch's tile plan."""
    dev = torch.cuda.current_device() if device is None or torch.device(device).index is None else torch.device(device).index
    L = _lib.device_lib(dev)
    ep = EdtrEpilogue()
    ep.workspace, ep.workspace_bytes, ep.max_clusters = _lib.gemm_scratch(dev)
    ep.out_mode, ep.act = OUT_BF16, ACT_NONE
    return int(L.edtr_gemm_row_stats_parts(M, N, K, ctypes.byref(ep)))


def _require_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("edtr_b200 ops need CUDA tensors (no CPU fallback)")


def rows_view(t: torch.Tensor, dtype=BF16) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a tensor usable as a row-major matrix with row stride ld."""
    if t.dtype != dtype:
        raise ValueError(f"expected dtype {dtype}, got {t.dtype}")
    if t.dim() < 2:
        raise ValueError("expected at least 2 dims")
    if t.stride(-1) != 1:
        raise ValueError("last dim must be contiguous")
    ld = t.stride(-2)
    cols = t.shape[-1]
    rows = 1
    expect = ld
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1 and t.stride(d) != expect:
            raise ValueError(f"tensor with shape {tuple(t.shape)} strides {t.stride()} is not a uniform rows view")
        expect *= t.shape[d]
        rows *= t.shape[d]
    return rows, cols, ld


def _f32(t: Optional[torch.Tensor], n: int, name: str) -> Optional[int]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() < n:
        raise ValueError(f"{name} must be a contiguous fp32 tensor with >= {n} elements")
    return t.data_ptr()


def _epilogue(M: int, n_out: int, out: torch.Tensor, *, bias, rowvec, rows_per_group, residual, act, out_mode, hw,
              alpha, ln=None, row_stats=None, gn_partial=None, gn_hw=0, gn_phases=1) -> EdtrEpilogue:
    # (the unit width of gn_partial, 2 or 4 channels, is read off its shape: [images, slabs, N / unit, 2])
    ep = EdtrEpilogue()
    dev = out.device.index if out.device.index is not None else torch.cuda.current_device()
    _lib.device_lib(dev)
    ep.workspace, ep.workspace_bytes, ep.max_clusters = _lib.gemm_scratch(dev)
    if ln is not None:
        stats, c, eps, colsum = ln
        n_w = 2 * n_out if act == ACT_GEGLU else n_out
        if stats.dtype != torch.float32 or stats.dim() != 3 or stats.shape[0] != M or stats.shape[2] != 2 \
                or not stats.is_contiguous():
            raise ValueError(f"ln stats must be a contiguous fp32 [M={M}, parts, 2] tensor, got {tuple(stats.shape)}")
        ep.ln_stats = stats.data_ptr()
        ep.ln_parts = stats.shape[1]
        ep.ln_c = int(c)
        ep.ln_eps = float(eps)
        ep.ln_colsum = _f32(colsum, n_w, "ln colsum")
    if row_stats is not None:
        if act == ACT_GEGLU or out_mode != OUT_BF16:
            raise ValueError("row_stats needs a bf16 output and act != GEGLU")
        if row_stats.dtype != torch.float32 or row_stats.dim() != 3 or row_stats.shape[0] != M \
                or row_stats.shape[2] != 2 or not row_stats.is_contiguous():
            raise ValueError(f"row_stats must be a contiguous fp32 [{M}, row_stats_parts(M, N, K), 2] tensor")
        ep.row_stats = row_stats.data_ptr()
        ep.row_stats_cap = row_stats.shape[1]
    if gn_partial is not None:
        # GroupNorm partial sums from the epilogue: fp32 [images, slabs, N/unit, 2], slabs = gn_phases * gn_hw / 32
        if act == ACT_GEGLU or out_mode != OUT_BF16:
            raise ValueError("gn_partial needs a bf16 output and act != GEGLU")
        if gn_hw <= 0 or gn_hw % 32 or (M // gn_phases) % gn_hw or n_out % 4:
            raise ValueError(f"gn_partial needs gn_hw % 32 == 0 and gn_hw | M (gn_hw {gn_hw}, M {M})")
        unit = n_out // gn_partial.shape[2] if gn_partial.dim() == 4 and gn_partial.shape[2] > 0 else 0
        want = (M // gn_phases // gn_hw, gn_phases * gn_hw // 32, n_out // max(unit, 1), 2)
        if unit not in (2, 4) or gn_partial.dtype != torch.float32 or tuple(gn_partial.shape) != want \
                or not gn_partial.is_contiguous():
            raise ValueError(f"gn_partial must be a contiguous fp32 [images, slabs, N/unit (unit 2 or 4), 2] tensor "
                             f"(e.g. {want}), got {tuple(gn_partial.shape)}")
        ep.gn_partial = gn_partial.data_ptr()
        ep.gn_hw = gn_hw
        ep.gn_slabs = want[1]
        ep.gn_slab0 = 0
        ep.gn_unit = unit
    ep.bias = _f32(bias, n_out if act != ACT_GEGLU else 2 * n_out, "bias")
    if rowvec is not None:
        if rowvec.dtype != torch.float32 or rowvec.dim() != 2 or rowvec.stride(1) != 1:
            raise ValueError("rowvec must be a 2-D fp32 tensor with contiguous rows")
        if rows_per_group <= 0 or M % rows_per_group != 0 or rowvec.shape[0] < M // rows_per_group:
            raise ValueError("rowvec / rows_per_group do not match the row count")
        ep.rowvec = rowvec.data_ptr()
        ep.rowvec_ld = rowvec.stride(0)
        ep.rows_per_group = rows_per_group
    if residual is not None:
        r_rows, r_cols, ldr = rows_view(residual)
        if r_rows != M or r_cols != n_out:
            raise ValueError(f"residual shape {tuple(residual.shape)} != ({M}, {n_out})")
        ep.residual = residual.data_ptr()
        ep.ldr = ldr
    ep.out = out.data_ptr()
    if out_mode == OUT_BF16:
        o_rows, o_cols, ldc = rows_view(out)
        if o_rows != M or o_cols != n_out:
            raise ValueError(f"out shape {tuple(out.shape)} != ({M}, {n_out})")
        ep.ldc = ldc
    elif out_mode == OUT_F32:
        o_rows, o_cols, ldc = rows_view(out, torch.float32)
        if o_rows != M or o_cols != n_out:
            raise ValueError(f"out shape {tuple(out.shape)} != ({M}, {n_out})")
        ep.ldc = ldc
    else:
        want = torch.float32 if out_mode == OUT_NCHW_F32 else BF16
        if out.dtype != want or not out.is_contiguous() or out.numel() != M * n_out:
            raise ValueError("NCHW output must be a contiguous tensor of M*N elements")
        ep.ldc = 0
    ep.act = act
    ep.out_mode = out_mode
    ep.hw = hw
    ep.alpha = alpha
    return ep


def _alloc_out(M: int, n_out: int, out_mode: int, hw: int, device) -> torch.Tensor:
    if out_mode == OUT_BF16:
        return torch.empty((M, n_out), dtype=BF16, device=device)
    if out_mode == OUT_F32:
        return torch.empty((M, n_out), dtype=torch.float32, device=device)
    dt = torch.float32 if out_mode == OUT_NCHW_F32 else BF16
    return torch.empty((M // hw, n_out, hw), dtype=dt, device=device)


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, rowvec=None, rows_per_group=0, residual=None,
         act=ACT_NONE, out=None, out_mode=OUT_BF16, hw=0, alpha=1.0, ln=None, row_stats=None, gn_partial=None,
         gn_hw=0) -> torch.Tensor:
    """``epilogue(a @ w.T)`` — a [..., K] rows view, w [N, K] (bf16).

    ln = (stats [M, parts, 2], C, eps, colsum [N]): LayerNorm over the K = C channels of `a` folded into the GEMM — `a`
    holds the un-normalised rows, `w` is pre-scaled by the LayerNorm gain, `bias` already contains W @ beta, `stats`
    are the (sum, sum of squares) partials a producer wrote through `row_stats` (engine.fold_layernorm packs these).
    row_stats = fp32 [M, N/32, 2]: receives per-row partial (sum, sum of squares) of the stored matrix.
    gn_partial = fp32 [M/gn_hw, gn_hw/32, N/4, 2]: receives GroupNorm partial sums of the stored matrix (gn_hw rows per
    image; see gn_partial_supported / groupnorm_from_partial)."""
    _require_cuda(a, w, bias, rowvec, residual, out)
    M, K, lda = rows_view(a)
    N, Kw, ldw = rows_view(w)
    if Kw != K:
        raise ValueError(f"K mismatch: a has {K}, w has {Kw}")
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = _alloc_out(M, n_out, out_mode, hw, a.device)
    if ln is not None and ln[1] != K:
        raise ValueError(f"folded LayerNorm width {ln[1]} != K {K}")
    ep = _epilogue(M, n_out, out, bias=bias, rowvec=rowvec, rows_per_group=rows_per_group, residual=residual,
                   act=act, out_mode=out_mode, hw=hw, alpha=alpha, ln=ln, row_stats=row_stats, gn_partial=gn_partial,
                   gn_hw=gn_hw)
    L = _lib.load()
    if row_stats is not None and row_stats.shape[1] != L.edtr_gemm_row_stats_parts(M, N, K, ctypes.byref(ep)):
        raise ValueError(f"row_stats must hold exactly row_stats_parts(M, N, K) = "
                         f"{L.edtr_gemm_row_stats_parts(M, N, K, ctypes.byref(ep))} pairs per row, got {row_stats.shape[1]}")
    _lib.check(L.edtr_gemm_bf16(a.data_ptr(), lda, w.data_ptr(), ldw, M, N, K, ctypes.byref(ep), _stream()),
               "edtr_gemm_bf16")
    return out


def conv3x3(x: torch.Tensor, w: torch.Tensor, *, bias=None, rowvec=None, residual=None, act=ACT_NONE, out=None,
            out_mode=OUT_BF16, alpha=1.0, gn_partial=None) -> torch.Tensor:
    """3x3/s1/p1 conv. x [B,H,W,Cin] channels-last rows view, w [Cout, 9*Cin] (tap-major)."""
    _require_cuda(x, w, bias, rowvec, residual, out)
    if x.dim() != 4:
        raise ValueError("x must be [B, H, W, C]")
    B, H, W, Cin = x.shape
    M, _, ldx = rows_view(x)
    Cout, Kw, ldw = rows_view(w)
    if Kw != 9 * Cin or ldw != Kw:
        raise ValueError(f"weight must be contiguous [Cout, 9*Cin={9 * Cin}], got {tuple(w.shape)}")
    hw = H * W
    if out is None:
        out = _alloc_out(M, Cout, out_mode, hw, x.device)
    ep = _epilogue(M, Cout, out, bias=bias, rowvec=rowvec, rows_per_group=hw, residual=residual, act=act,
                   out_mode=out_mode, hw=hw, alpha=alpha, gn_partial=gn_partial, gn_hw=hw)
    L = _lib.device_lib()
    _lib.check(L.edtr_conv3x3_bf16(x.data_ptr(), ldx, B, H, W, Cin, w.data_ptr(), Cout, ctypes.byref(ep), _stream()),
               "edtr_conv3x3_bf16")
    return out


def conv3x3_up2x_supported(B: int, H: int, W: int, Cin: int, Cout: int) -> bool:
    """Geometry the phase-decomposed (sub-pixel) up-sampling convolution covers."""
    if Cin % 64 or Cout % 64 or B * H * W < 256 or W < 8 or (W & (W - 1)):
        return False
    rows = max(1, 128 // W)
    return H % rows == 0 if H >= rows else rows % H == 0


def conv3x3_up2x(x: torch.Tensor, w4: torch.Tensor, *, bias=None, act=ACT_NONE, out=None, gn_partial=None) -> torch.Tensor:
    """nearest-2x + 3x3/p1 conv as four 2x2-tap convs. x [B,H,W,Cin]; w4 [4, Cout, 4*Cin] phase filters
    (engine.pack_conv3x3_up2x); out [B,2H,2W,Cout] rows view."""
    _require_cuda(x, w4, bias, out)
    if x.dim() != 4:
        raise ValueError("x must be [B, H, W, C]")
    B, H, W, Cin = x.shape
    _, _, ldx = rows_view(x)
    if w4.dtype != BF16 or w4.dim() != 3 or w4.shape[0] != 4 or w4.shape[2] != 4 * Cin or not w4.is_contiguous():
        raise ValueError(f"w4 must be a contiguous bf16 [4, Cout, 4*Cin={4 * Cin}] tensor, got {tuple(w4.shape)}")
    Cout = w4.shape[1]
    if not conv3x3_up2x_supported(B, H, W, Cin, Cout):
        raise ValueError(f"unsupported up-conv geometry B={B} H={H} W={W} Cin={Cin} Cout={Cout}")
    if act == ACT_GEGLU:
        raise ValueError("GEGLU is not defined for convolutions")
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, Cout), dtype=BF16, device=x.device)
    ep = _epilogue(4 * B * H * W, Cout, out, bias=bias, rowvec=None, rows_per_group=0, residual=None, act=act,
                   out_mode=OUT_BF16, hw=0, alpha=1.0, gn_partial=gn_partial, gn_hw=H * W, gn_phases=4)
    if tuple(out.shape) != (B, 2 * H, 2 * W, Cout):
        raise ValueError(f"out must be [B, 2H, 2W, Cout], got {tuple(out.shape)}")
    L = _lib.device_lib()
    _lib.check(L.edtr_conv3x3_up2x_bf16(x.data_ptr(), ldx, B, H, W, Cin, w4.data_ptr(), Cout, ctypes.byref(ep),
                                        _stream()), "edtr_conv3x3_up2x_bf16")
    return out


def conv3x3_supported(H: int, W: int, Cin: int) -> bool:
    """Geometry the TMA implicit-GEMM path covers (else: im2col + gemm)."""
    if Cin % 64 != 0:
        return False
    if W >= 128:
        return W % 128 == 0
    if W < 8 or 128 % W != 0:
        return False
    rows = 128 // W
    return H % rows == 0 if H >= rows else rows % H == 0


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T * scale) v per head (d=64). q [B,Lq,heads*64], k/v [B,Lk,heads*64] rows views."""
    _require_cuda(q, k, v, out)
    for t in (q, k, v):
        if t.dim() != 3:
            raise ValueError("q/k/v must be [B, L, heads*64]")
    B, Lq, C = q.shape
    Bk, Lk, Ck = k.shape
    if C != heads * 64 or Ck != C or v.shape != k.shape or Bk != B:
        raise ValueError("attention shape mismatch")
    _, _, ldq = rows_view(q)
    _, _, ldk = rows_view(k)
    _, _, ldv = rows_view(v)
    if out is None:
        out = torch.empty((B, Lq, C), dtype=BF16, device=q.device)
    _, _, ldo = rows_view(out)
    L = _lib.device_lib()
    _lib.check(L.edtr_attention_bf16(q.data_ptr(), ldq, k.data_ptr(), ldk, v.data_ptr(), ldv, out.data_ptr(), ldo,
                                     B, heads, Lq, Lk, scale, _stream()), "edtr_attention_bf16")
    return out


GROUPNORM_FUSED = True   # tests flip this to exercise the two-pass kernels on small tensors too


def groupnorm_partial_size(B: int, HW: int, C: int, groups: int) -> int:
    """fp32 elements of the partial-sum scratch edtr_groupnorm_stats writes (no device needed)."""
    return int(_lib.load().edtr_groupnorm_partial_size(B, HW, C, groups))


def groupnorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool,
              stats: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm (+SiLU) over x [B, HW..., C] channels-last rows view. `stats` = optional fp32 scratch of
    at least groupnorm_partial_size(...) elements (contents are overwritten)."""
    _require_cuda(x, gamma, beta, stats, out)
    B = x.shape[0]
    M, C, ldx = rows_view(x)
    HW = M // B
    if C % 8 != 0 or C % groups != 0:
        raise ValueError(f"C ({C}) must be a multiple of 8 and of groups ({groups})")
    need = groupnorm_partial_size(B, HW, C, groups)
    if stats is not None and (stats.dtype != torch.float32 or stats.numel() < need or not stats.is_contiguous()):
        raise ValueError(f"stats must be a contiguous fp32 scratch tensor with >= {need} elements")
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    Mo, Co, ldy = rows_view(out)
    if Mo != M or Co != C:
        raise ValueError("out shape mismatch")
    g = _f32(gamma, C, "gamma")
    b = _f32(beta, C, "beta")
    L = _lib.device_lib()
    st = _stream()
    if GROUPNORM_FUSED and L.edtr_groupnorm_fused_supported(B, HW, C, groups):
        # L2-resident tensor: one launch, a cluster per image (statistics through distributed shared memory)
        _lib.check(L.edtr_groupnorm_fused(x.data_ptr(), ldx, out.data_ptr(), ldy, B, HW, C, groups, g, b, eps,
                                          1 if silu else 0, st), "edtr_groupnorm_fused")
        return out
    if stats is None:
        stats = torch.empty((need,), dtype=torch.float32, device=x.device)
    _lib.check(L.edtr_groupnorm_stats(x.data_ptr(), ldx, B, HW, C, groups, stats.data_ptr(), st),
               "edtr_groupnorm_stats")
    _lib.check(L.edtr_groupnorm_apply(x.data_ptr(), ldx, out.data_ptr(), ldy, B, HW, C, groups, stats.data_ptr(),
                                      g, b, eps, 1 if silu else 0, st), "edtr_groupnorm_apply")
    return out


def groupnorm_pool(x: torch.Tensor, groups: int, weight: float, acc: torch.Tensor,
                   stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Tiled VAE: acc[B, groups, 2] += weight * (mean, biased var) of tile x [B, HW..., C] per (image, group)
    (GroupNormParam.add_tile / summary, utils/tilevae/tilevae.py:241-278)."""
    _require_cuda(x, acc, stats)
    B = x.shape[0]
    M, C, ldx = rows_view(x)
    HW = M // B
    if C % 8 != 0 or C % groups != 0:
        raise ValueError(f"C ({C}) must be a multiple of 8 and of groups ({groups})")
    if acc.dtype != torch.float32 or tuple(acc.shape) != (B, groups, 2) or not acc.is_contiguous():
        raise ValueError(f"acc must be a contiguous fp32 [{B}, {groups}, 2] tensor")
    need = groupnorm_partial_size(B, HW, C, groups)
    if stats is None:
        stats = torch.empty((need,), dtype=torch.float32, device=x.device)
    elif stats.dtype != torch.float32 or stats.numel() < need or not stats.is_contiguous():
        raise ValueError(f"stats must be a contiguous fp32 scratch tensor with >= {need} elements")
    L = _lib.device_lib()
    st = _stream()
    _lib.check(L.edtr_groupnorm_stats(x.data_ptr(), ldx, B, HW, C, groups, stats.data_ptr(), st),
               "edtr_groupnorm_stats")
    _lib.check(L.edtr_groupnorm_pool(stats.data_ptr(), B, HW, C, groups, float(weight), acc.data_ptr(), st),
               "edtr_groupnorm_pool")
    return acc


def gn_partial_unit(C: int, groups: int = 32) -> int:
    """Channels per unit of the epilogue's GroupNorm partial sums: 4 when the group width allows it, else 2 (0: none)."""
    if C % groups:
        return 0
    cpg = C // groups
    return 4 if cpg % 4 == 0 else (2 if cpg % 2 == 0 else 0)


def gn_partial_supported(M: int, HW: int, N: int, K: int, groups: int = 32) -> bool:
    """True when the GEMM / convolution that produces an [M, N] tensor (HW rows per image) can deliver the GroupNorm
    partial sums from its epilogue: CTA-pair kernel, 32-row slabs inside one image, 2- or 4-channel units inside one
    group, and a shape the planner would not split along K anyway."""
    if M < 256 or N % 64 or HW % 32 or M % HW or gn_partial_unit(N, groups) == 0:
        return False
    return gemm_workspace_size(M, N, K) == 0


def gn_partial_shape(B: int, HW: int, C: int, groups: int = 32) -> Tuple[int, int, int, int]:
    return (B, HW // 32, C // gn_partial_unit(C, groups), 2)


def groupnorm_fold(gn_partial: torch.Tensor, C: int, groups: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(mean, biased variance) per (image, group), fp32 [B, groups, 2], from the partial sums an epilogue wrote
    (gn_partial fp32 [B, slabs, C/unit, 2], C channels); feeds groupnorm_apply_stats."""
    _require_cuda(gn_partial, out)
    if gn_partial.dtype != torch.float32 or gn_partial.dim() != 4 or gn_partial.shape[3] != 2 or not gn_partial.is_contiguous():
        raise ValueError("gn_partial must be a contiguous fp32 [B, slabs, C/unit, 2] tensor")
    B, slabs, units, _ = gn_partial.shape
    if C % units or C // units not in (2, 4):
        raise ValueError(f"gn_partial has {units} units per row: not C / 2 or C / 4 of C = {C}")
    if out is None:
        out = torch.empty((B, groups, 2), dtype=torch.float32, device=gn_partial.device)
    elif out.dtype != torch.float32 or tuple(out.shape) != (B, groups, 2) or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous fp32 [{B}, {groups}, 2] tensor")
    _lib.check(_lib.device_lib().edtr_groupnorm_fold(gn_partial.data_ptr(), B, slabs, C, groups, C // units,
                                                     out.data_ptr(), _stream()), "edtr_groupnorm_fold")
    return out


def groupnorm_apply_stats(x: torch.Tensor, mean_var: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                          groups: int, eps: float, silu: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm (+SiLU) with GIVEN statistics mean_var [B, groups, 2] = (mean, biased var)
    (custom_group_norm, utils/tilevae/tilevae.py:188-215)."""
    _require_cuda(x, mean_var, gamma, beta, out)
    B = x.shape[0]
    M, C, ldx = rows_view(x)
    HW = M // B
    if C % 8 != 0 or C % groups != 0:
        raise ValueError(f"C ({C}) must be a multiple of 8 and of groups ({groups})")
    if mean_var.dtype != torch.float32 or tuple(mean_var.shape) != (B, groups, 2) or not mean_var.is_contiguous():
        raise ValueError(f"mean_var must be a contiguous fp32 [{B}, {groups}, 2] tensor")
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    Mo, Co, ldy = rows_view(out)
    if Mo != M or Co != C:
        raise ValueError("out shape mismatch")
    L = _lib.device_lib()
    _lib.check(L.edtr_groupnorm_apply_stats(x.data_ptr(), ldx, out.data_ptr(), ldy, B, HW, C, groups,
                                            mean_var.data_ptr(), _f32(gamma, C, "gamma"), _f32(beta, C, "beta"),
                                            eps, 1 if silu else 0, _stream()), "edtr_groupnorm_apply_stats")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out: Optional[torch.Tensor] = None, c_real: Optional[int] = None) -> torch.Tensor:
    """Row LayerNorm.  c_real < C: rows are padded to C channels, the statistics use the first c_real channels and the
    pads (zeros on input; gamma = beta = 0 there) stay zero."""
    _require_cuda(x, gamma, beta, out)
    M, C, ldx = rows_view(x)
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    _, _, ldy = rows_view(out)
    L = _lib.device_lib()
    _lib.check(L.edtr_layernorm_padded_bf16(x.data_ptr(), ldx, out.data_ptr(), ldy, M, C, C if c_real is None else c_real,
                                            _f32(gamma, C, "gamma"), _f32(beta, C, "beta"), eps, _stream()),
               "edtr_layernorm_padded_bf16")
    return out


def pixel_unshuffle(x: torch.Tensor, out: torch.Tensor, mean: Sequence[float], scale: float, r: int = 8) -> torch.Tensor:
    """fp32 NCHW image -> bf16 channels-last [B, H/r, W/r, >= C r^2] with the per-channel mean removed
    (nn.PixelUnshuffle channel order)."""
    _require_cuda(x, out)
    if x.dtype != torch.float32 or x.dim() != 4 or not x.is_contiguous():
        raise ValueError("x must be a contiguous fp32 [B, C, H, W] tensor")
    B, C, H, W = x.shape
    if out.dtype != BF16 or out.dim() != 4 or tuple(out.shape[:3]) != (B, H // r, W // r) or out.shape[3] < C * r * r:
        raise ValueError(f"out must be bf16 [B, H/{r}, W/{r}, >= {C * r * r}], got {tuple(out.shape)}")
    _, _, ldy = rows_view(out)
    m = (ctypes.c_float * 3)(*([float(v) for v in mean] + [0.0] * 3)[:3])
    L = _lib.device_lib()
    _lib.check(L.edtr_pixel_unshuffle_f32_to_nhwc_bf16(x.data_ptr(), out.data_ptr(), ldy, B, C, H, W, r,
                                                       ctypes.cast(m, ctypes.c_void_p), float(scale), _stream()),
               "edtr_pixel_unshuffle_f32_to_nhwc_bf16")
    return out


def window_attention(qkv: torch.Tensor, heads: int, shift: int, scale: float, bias: torch.Tensor,
                     mask: Optional[torch.Tensor], out: torch.Tensor) -> torch.Tensor:
    """Swin window attention (8x8 windows) on a [B, H, W, 3*heads*32] fused projection; out [B, H, W, heads*32]."""
    _require_cuda(qkv, bias, mask, out)
    if qkv.dim() != 4 or out.dim() != 4 or qkv.shape[:3] != out.shape[:3]:
        raise ValueError("qkv / out must be [B, H, W, C] tensors over the same token grid")
    B, H, W, C3 = qkv.shape
    _, _, ld = rows_view(qkv)
    _, _, ldo = rows_view(out)
    if C3 < 3 * heads * 32 or out.shape[3] < heads * 32:
        raise ValueError("qkv / out are too narrow for heads * 32 columns")
    nb = heads * 64 * 64
    if bias.dtype != torch.float32 or not bias.is_contiguous() or bias.numel() != nb:
        raise ValueError(f"bias must be a contiguous fp32 [heads, 64, 64] tensor ({nb} elements)")
    mp = None
    if mask is not None:
        nm = (H // 8) * (W // 8) * 64 * 64
        if mask.dtype != torch.float32 or not mask.is_contiguous() or mask.numel() != nm:
            raise ValueError(f"mask must be a contiguous fp32 [windows, 64, 64] tensor ({nm} elements)")
        mp = mask.data_ptr()
    L = _lib.device_lib()
    _lib.check(L.edtr_window_attention_bf16(qkv.data_ptr(), ld, out.data_ptr(), ldo, B, H, W, heads, shift, float(scale),
                                            bias.data_ptr(), mp, _stream()), "edtr_window_attention_bf16")
    return out


def softmax_rows(s: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(s, out)
    M, N, lds = rows_view(s, torch.float32)
    if out is None:
        out = torch.empty(s.shape, dtype=BF16, device=s.device)
    Mo, No, ldp = rows_view(out)
    if (Mo, No) != (M, N):
        raise ValueError("softmax out shape mismatch")
    L = _lib.device_lib()
    _lib.check(L.edtr_softmax_rows(s.data_ptr(), lds, out.data_ptr(), ldp, M, N, scale, _stream()),
               "edtr_softmax_rows")
    return out


def upsample2x(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, out)
    B, H, W, C = x.shape
    _, _, ldx = rows_view(x)
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, C), dtype=BF16, device=x.device)
    _, _, ldy = rows_view(out)
    L = _lib.device_lib()
    _lib.check(L.edtr_upsample2x_bf16(x.data_ptr(), ldx, out.data_ptr(), ldy, B, H, W, C, _stream()),
               "edtr_upsample2x_bf16")
    return out


def im2col(x: torch.Tensor, kh: int, kw: int, stride: int, pad_top: int, pad_left: int, Ho: int, Wo: int,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, out)
    B, H, W, C = x.shape
    _, _, ldx = rows_view(x)
    if out is None:
        out = torch.empty((B * Ho * Wo, kh * kw * C), dtype=BF16, device=x.device)
    elif out.dtype != BF16 or not out.is_contiguous() or out.numel() != B * Ho * Wo * kh * kw * C:
        raise ValueError("im2col out must be a contiguous bf16 [B*Ho*Wo, kh*kw*C] tensor")
    L = _lib.device_lib()
    _lib.check(L.edtr_im2col_bf16(x.data_ptr(), ldx, out.data_ptr(), B, H, W, C, kh, kw, stride, pad_top, pad_left,
                                  Ho, Wo, _stream()), "edtr_im2col_bf16")
    return out


def nchw_to_nhwc(x: torch.Tensor, out: torch.Tensor, coff: int = 0) -> torch.Tensor:
    """x [B,C,H,W] fp32 contiguous -> out[..., coff:coff+C] (bf16 channels-last rows view)."""
    _require_cuda(x, out)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 4:
        raise ValueError("x must be a contiguous fp32 NCHW tensor")
    B, C, H, W = x.shape
    M, Co, ldy = rows_view(out)
    if M != B * H * W or coff + C > Co:
        raise ValueError("out does not match x")
    L = _lib.device_lib()
    _lib.check(L.edtr_nchw_f32_to_nhwc_bf16(x.data_ptr(), out.data_ptr(), ldy, coff, B, C, H * W, _stream()),
               "edtr_nchw_f32_to_nhwc_bf16")
    return out


def pointwise_nchw_to_nhwc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], scale: float,
                           out: torch.Tensor, coff: int = 0) -> torch.Tensor:
    """out[..., coff:coff+Cout] = bias + w @ (scale * x) per pixel; x [B,Cin,H,W] fp32, w [Cout,Cin] fp32."""
    _require_cuda(x, w, bias, out)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 4:
        raise ValueError("x must be a contiguous fp32 NCHW tensor")
    B, Cin, H, W = x.shape
    if w.dtype != torch.float32 or not w.is_contiguous() or w.dim() != 2 or w.shape[1] != Cin:
        raise ValueError("w must be a contiguous fp32 [Cout, Cin] matrix")
    Cout = w.shape[0]
    M, Co, ldy = rows_view(out)
    if M != B * H * W or coff + Cout > Co:
        raise ValueError("out does not match x")
    L = _lib.device_lib()
    _lib.check(L.edtr_pointwise_nchw_f32_to_nhwc_bf16(x.data_ptr(), w.data_ptr(), _f32(bias, Cout, "bias"), scale,
                                                      out.data_ptr(), ldy, coff, B, Cin, Cout, H * W, _stream()),
               "edtr_pointwise_nchw_f32_to_nhwc_bf16")
    return out


def nhwc_to_nchw(x: torch.Tensor, B: int, out_f32: bool = True) -> torch.Tensor:
    """x [B*HW, C] rows view (bf16) -> [B, C, HW] contiguous."""
    _require_cuda(x)
    M, C, ldx = rows_view(x)
    HW = M // B
    out = torch.empty((B, C, HW), dtype=torch.float32 if out_f32 else BF16, device=x.device)
    L = _lib.device_lib()
    _lib.check(L.edtr_nhwc_bf16_to_nchw(x.data_ptr(), ldx, out.data_ptr(), B, C, HW, 1 if out_f32 else 0, _stream()),
               "edtr_nhwc_bf16_to_nchw")
    return out


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(x, out)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("x must be contiguous fp32")
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    elif out.dtype != BF16 or not out.is_contiguous() or out.numel() != x.numel():
        raise ValueError("cast out must be a contiguous bf16 tensor of the same size")
    L = _lib.device_lib()
    _lib.check(L.edtr_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "edtr_cast_f32_to_bf16")
    return out


def tile_blend(tiles: torch.Tensor, coords: torch.Tensor, weight: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[b,c,y,x] = sum_t weight * tiles[t,b,c,...] over the tiles covering (y,x). tiles [T,B,C,th,tw] fp32,
    coords int32 [T,2] = (hi, wi), weight [th,tw] fp32, out [B,C,H,W] fp32."""
    _require_cuda(tiles, coords, weight, out)
    if tiles.dim() != 5 or tiles.dtype != torch.float32 or not tiles.is_contiguous():
        raise ValueError("tiles must be a contiguous fp32 [T, B, C, th, tw] tensor")
    T, B, C, th, tw = tiles.shape
    if coords.dtype != torch.int32 or tuple(coords.shape) != (T, 2) or not coords.is_contiguous():
        raise ValueError("coords must be a contiguous int32 [T, 2] tensor")
    if weight.dtype != torch.float32 or tuple(weight.shape) != (th, tw) or not weight.is_contiguous():
        raise ValueError("weight must be a contiguous fp32 [th, tw] tensor")
    if out.dtype != torch.float32 or not out.is_contiguous() or out.dim() != 4 or tuple(out.shape[:2]) != (B, C):
        raise ValueError("out must be a contiguous fp32 [B, C, H, W] tensor")
    H, W = out.shape[2:]
    L = _lib.device_lib()
    _lib.check(L.edtr_tile_blend(tiles.data_ptr(), coords.data_ptr(), T, weight.data_ptr(), out.data_ptr(), B * C, H, W,
                                 th, tw, _stream()), "edtr_tile_blend")
    return out


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(t, out)
    if t.dtype != torch.int64 or t.dim() != 1 or not t.is_contiguous():
        raise ValueError("timesteps must be a contiguous int64 vector")
    if out is None:
        out = torch.empty((t.shape[0], dim), dtype=BF16, device=t.device)
    elif out.dtype != BF16 or not out.is_contiguous() or tuple(out.shape) != (t.shape[0], dim):
        raise ValueError("timestep embedding out must be a contiguous bf16 [B, dim] tensor")
    L = _lib.device_lib()
    _lib.check(L.edtr_timestep_embedding(t.data_ptr(), out.data_ptr(), t.shape[0], dim, max_period, _stream()),
               "edtr_timestep_embedding")
    return out


def sampler_update(x, eps, noise, index, tables, want_pred_x0: bool = True, x_prev=None, pred_x0=None):
    """Fused p_sample arithmetic; tables = (sqrt_recip, sqrt_recipm1, coef1, coef2, var) fp32 device vectors."""
    _require_cuda(x, eps, noise, index, *tables)
    for t in (x, eps, noise, x_prev, pred_x0):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.shape != x.shape):
            raise ValueError("x / eps / noise / outputs must be contiguous fp32 tensors of one shape")
    if index.dtype != torch.int64 or index.numel() != x.shape[0] or not index.is_contiguous():
        raise ValueError("index must be contiguous int64 [B]")
    for t in tables:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("coefficient tables must be contiguous fp32 vectors")
    B = x.shape[0]
    n = x.numel() // B
    if x_prev is None:
        x_prev = torch.empty_like(x)
    pred = pred_x0 if pred_x0 is not None else (torch.empty_like(x) if want_pred_x0 else None)
    with device_guard(x):
        L = _lib.device_lib()
        _lib.check(L.edtr_sampler_update(x.data_ptr(), eps.data_ptr(), noise.data_ptr(), index.data_ptr(),
                                         *[t.data_ptr() for t in tables], x_prev.data_ptr(),
                                         pred.data_ptr() if pred is not None else None, B, n, _stream()),
                   "edtr_sampler_update")
    return x_prev, pred


def wavelet_reconstruction(content: torch.Tensor, style: torch.Tensor, levels: int = 5) -> torch.Tensor:
    """utils/common.py:136-147: high frequencies of `content` + low frequencies of `style` (both [B, C, H, W] fp32),
    five dilated binomial blurs per image, the final sum fused into the last style level."""
    _require_cuda(content, style)
    if content.shape != style.shape or content.dim() != 4:
        raise ValueError(f"content / style must be equal-shaped [B, C, H, W], got {tuple(content.shape)} {tuple(style.shape)}")
    if content.dtype != torch.float32 or style.dtype != torch.float32:
        raise ValueError("the colour fix runs in fp32")
    content, style = content.contiguous(), style.contiguous()
    B, C, H, W = content.shape
    with device_guard(content):
        return _wavelet_reconstruction(content, style, levels, B, C, H, W)


def _wavelet_reconstruction(content, style, levels, B, C, H, W):
    L = _lib.device_lib()
    st = _stream()
    high = torch.empty_like(content)
    ping = [torch.empty_like(content), torch.empty_like(content)]
    cur = content
    for i in range(levels):      # content: accumulate image - low
        dst = ping[i % 2]
        _lib.check(L.edtr_wavelet_level(cur.data_ptr(), dst.data_ptr(), high.data_ptr(), B * C, H, W, 2 ** i, 1,
                                        1 if i == 0 else 0, st), "edtr_wavelet_level")
        cur = dst
    cur = style
    out = torch.empty_like(content)
    for i in range(levels):      # style: keep the low pass; the last level adds the content's high pass
        last = i == levels - 1
        dst = out if last else ping[i % 2]
        _lib.check(L.edtr_wavelet_level(cur.data_ptr(), dst.data_ptr(), high.data_ptr(), B * C, H, W, 2 ** i,
                                        2 if last else 0, 0, st), "edtr_wavelet_level")
        cur = dst
    return out
