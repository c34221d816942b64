"""TEST INFRASTRUCTURE ONLY — a torch/CPU stand-in for ``edtr_b200.ops``.

Mirrors the semantics of every C-ABI kernel (include/edtr_b200.h) with plain fp32 torch
ops so that the host-side engine logic (block walk, buffer aliasing, channel-offset
writes, in-place zero-conv accumulation, weight packing, GEGLU interleave, hoisted
cross-attention K/V) can be checked against the oracle without a GPU.  Outputs are
written into the ``out`` views the engine passes, in the dtype of those views (bf16
storage rounding included).  Never imported by the product.
"""
import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_GELU, ACT_LRELU_02, ACT_LRELU_001 = 0, 1, 2, 3, 4, 5
OUT_BF16, OUT_F32, OUT_NCHW_F32, OUT_NCHW_BF16 = 0, 1, 2, 3
REQUIRES_CUDA = False
GEGLU_TILE = 128
LAUNCHES = [0]


def set_gemm_max_clusters(n):
    pass


def use_workspace(index):
    pass


def geglu_tile_n():
    return GEGLU_TILE


def groupnorm_partial_size(B, HW, C, groups):
    return 64


def _rows(t):
    return t.reshape(-1, t.shape[-1]).float()


def row_stats_parts(M, N, K, device=None):
    return 3     # any split of the columns works for the consumer: it only sums the parts


def device_guard(device):
    import contextlib

    return contextlib.nullcontext()


def _finish(v, M, n_out, *, bias, rowvec, rows_per_group, residual, act, out, out_mode, hw, alpha, geglu_bias=None,
            ln=None, row_stats=None, gn_partial=None, gn_hw=0):
    LAUNCHES[0] += 1
    if ln is not None:
        # folded LayerNorm (EdtrEpilogue.ln_*): acc' = rstd * (acc - mean * colsum[n]) from the producer's partial sums
        stats, c, eps, colsum = ln
        assert stats.dtype == torch.float32 and stats.shape[0] == M and stats.shape[2] == 2
        s1, s2 = stats[:, :, 0].sum(1), stats[:, :, 1].sum(1)
        mu = s1 / c
        rstd = torch.rsqrt(torch.clamp(s2 / c - mu * mu, min=0.0) + eps)
        v = rstd[:, None] * (v - mu[:, None] * colsum.float()[None, :])
    v = v * alpha
    if act == ACT_GEGLU:
        if bias is not None:
            v = v + bias.float()
        half = GEGLU_TILE // 2
        v = v.view(M, -1, 2, half)
        v = (v[:, :, 0] * F.gelu(v[:, :, 1])).reshape(M, n_out)
    elif bias is not None:
        v = v + bias.float()
    if rowvec is not None:
        v = v + rowvec.float().repeat_interleave(rows_per_group, 0)[:M]
    if residual is not None:
        v = v + _rows(residual)
    if act == ACT_SILU:
        v = F.silu(v)
    elif act == ACT_GELU:
        v = F.gelu(v)
    elif act == ACT_LRELU_02:
        v = F.leaky_relu(v, 0.2)
    elif act == ACT_LRELU_001:
        v = F.leaky_relu(v, 0.01)
    if row_stats is not None:
        assert act != ACT_GEGLU and out_mode == OUT_BF16 and row_stats.shape[0] == M and row_stats.shape[2] == 2
        for j, vv in enumerate(torch.tensor_split(v, row_stats.shape[1], dim=1)):
            row_stats[:, j, 0] = vv.sum(-1)
            row_stats[:, j, 1] = (vv * vv).sum(-1)
    if gn_partial is not None:
        # EdtrEpilogue.gn_partial: (sum, sum of squares) per (32-row slab, 4-column unit) of the fp32 values
        assert act != ACT_GEGLU and out_mode == OUT_BF16 and gn_hw % 32 == 0 and M % gn_hw == 0 and n_out % 4 == 0
        unit = n_out // gn_partial.shape[2]
        assert unit in (2, 4) and gn_partial.dtype == torch.float32
        assert tuple(gn_partial.shape) == (M // gn_hw, gn_hw // 32, n_out // unit, 2)
        u = v.view(M // gn_hw, gn_hw // 32, 32, n_out // unit, unit)
        gn_partial[..., 0] = u.sum((2, 4))
        gn_partial[..., 1] = (u * u).sum((2, 4))
    if out_mode in (OUT_BF16, OUT_F32):
        want = torch.bfloat16 if out_mode == OUT_BF16 else torch.float32
        if out is None:
            out = torch.empty((M, n_out), dtype=want)
        assert out.dtype == want and out.numel() == M * n_out, (out.dtype, out.shape, M, n_out)
        out.copy_(v.view(out.shape))
    else:
        want = torch.float32 if out_mode == OUT_NCHW_F32 else torch.bfloat16
        if out is None:
            out = torch.empty((M // hw, n_out, hw), dtype=want)
        assert out.dtype == want and out.is_contiguous() and out.numel() == M * n_out
        out.view(M // hw, n_out, hw).copy_(v.view(M // hw, hw, n_out).permute(0, 2, 1))
    return out


def gemm(a, w, *, bias=None, rowvec=None, rows_per_group=0, residual=None, act=ACT_NONE, out=None,
         out_mode=OUT_BF16, hw=0, alpha=1.0, ln=None, row_stats=None, gn_partial=None, gn_hw=0):
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    A = _rows(a)
    M, K = A.shape
    N, Kw = w.shape
    assert K == Kw and K % 64 == 0, (K, Kw)
    n_out = N // 2 if act == ACT_GEGLU else N
    return _finish(A @ w.float().t(), M, n_out, bias=bias, rowvec=rowvec, rows_per_group=rows_per_group,
                   residual=residual, act=act, out=out, out_mode=out_mode, hw=hw, alpha=alpha, ln=ln,
                   row_stats=row_stats, gn_partial=gn_partial, gn_hw=gn_hw)


def conv3x3(x, w, *, bias=None, rowvec=None, residual=None, act=ACT_NONE, out=None, out_mode=OUT_BF16, alpha=1.0,
            gn_partial=None):
    assert x.dtype == torch.bfloat16 and x.dim() == 4
    B, H, W, Cin = x.shape
    assert Cin % 64 == 0
    Cout = w.shape[0]
    assert w.shape[1] == 9 * Cin
    wt = w.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    return _finish(y, B * H * W, Cout, bias=bias, rowvec=rowvec, rows_per_group=H * W, residual=residual, act=act,
                   out=out, out_mode=out_mode, hw=H * W, alpha=alpha, gn_partial=gn_partial, gn_hw=H * W)


def conv3x3_up2x_supported(B, H, W, Cin, Cout):
    if Cin % 64 or Cout % 64 or B * H * W < 256 or W < 8 or (W & (W - 1)):
        return False
    rows = max(1, 128 // W)
    return H % rows == 0 if H >= rows else rows % H == 0


def conv3x3_up2x(x, w4, *, bias=None, act=ACT_NONE, out=None, gn_partial=None):
    """Four 2x2-tap phase convolutions, evaluated literally from the packed phase filters."""
    LAUNCHES[0] += 4
    B, H, W, Cin = x.shape
    Cout = w4.shape[1]
    assert w4.shape == (4, Cout, 4 * Cin)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))   # [B, C, H+2, W+2], source pixel (y, x) at (y+1, x+1)
    y = torch.zeros((B, 2 * H, 2 * W, Cout), dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            wp = w4[py * 2 + px].float().view(Cout, 2, 2, Cin)
            acc = torch.zeros((B, Cout, H, W), dtype=torch.float32)
            for dy in (0, 1):
                for dx in (0, 1):
                    oy, ox = dy + py - 1, dx + px - 1          # source offset
                    patch = xp[:, :, 1 + oy:1 + oy + H, 1 + ox:1 + ox + W]
                    acc += torch.einsum("oc,bchw->bohw", wp[:, dy, dx, :], patch)
            y[:, py::2, px::2, :] = acc.permute(0, 2, 3, 1)
    if bias is not None:
        y = y + bias.float()
    if act == ACT_SILU:
        y = F.silu(y)
    elif act == ACT_LRELU_02:
        y = F.leaky_relu(y, 0.2)
    if gn_partial is not None:
        # slabs of an image: phase-major (four launches), 32 consecutive low-resolution pixels each
        unit = Cout // gn_partial.shape[2]
        assert (H * W) % 32 == 0 and unit in (2, 4) and tuple(gn_partial.shape) == (B, 4 * H * W // 32, Cout // unit, 2)
        for py in (0, 1):
            for px in (0, 1):
                u = y[:, py::2, px::2, :].reshape(B, H * W // 32, 32, Cout // unit, unit)
                sl = slice((py * 2 + px) * (H * W // 32), (py * 2 + px + 1) * (H * W // 32))
                gn_partial[:, sl, :, 0] = u.sum((2, 4))
                gn_partial[:, sl, :, 1] = (u * u).sum((2, 4))
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, Cout), dtype=torch.bfloat16)
    out.copy_(y)
    return out


def attention(q, k, v, heads, scale, out=None):
    LAUNCHES[0] += 1
    B, Lq, C = q.shape
    split = lambda t: t.float().view(B, t.shape[1], heads, 64).permute(0, 2, 1, 3)
    w = torch.softmax(split(q) @ split(k).transpose(-1, -2) * scale, dim=-1)
    o = (w @ split(v)).permute(0, 2, 1, 3).reshape(B, Lq, C)
    if out is None:
        out = torch.empty((B, Lq, C), dtype=torch.bfloat16)
    out.copy_(o)
    return out


def groupnorm(x, gamma, beta, groups, eps, silu, stats=None, out=None):
    LAUNCHES[0] += 2
    B, C = x.shape[0], x.shape[-1]
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.numel() >= 1
    xf = x.float().reshape(B, -1, C).permute(0, 2, 1)
    y = F.group_norm(xf, groups, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16)
    out.copy_(y.permute(0, 2, 1).reshape(out.shape))
    return out


def groupnorm_pool(x, groups, weight, acc, stats=None):
    LAUNCHES[0] += 2
    B, C = x.shape[0], x.shape[-1]
    v = x.float().reshape(B, -1, groups, C // groups).permute(0, 2, 1, 3).reshape(B, groups, -1)
    var, mean = torch.var_mean(v, dim=2, unbiased=False)
    acc[..., 0] += weight * mean
    acc[..., 1] += weight * var
    return acc


GN_PARTIAL = True   # tests flip this to run the engines without epilogue GroupNorm statistics


def gn_partial_unit(C, groups=32):
    if C % groups:
        return 0
    cpg = C // groups
    return 4 if cpg % 4 == 0 else (2 if cpg % 2 == 0 else 0)


def gn_partial_supported(M, HW, N, K, groups=32):
    return bool(GN_PARTIAL and M >= 256 and N % 64 == 0 and HW % 32 == 0 and M % HW == 0 and gn_partial_unit(N, groups))


def gn_partial_shape(B, HW, C, groups=32):
    return (B, HW // 32, C // gn_partial_unit(C, groups), 2)


def groupnorm_fold(gn_partial, C, groups, out=None):
    LAUNCHES[0] += 1
    B, slabs, units, _ = gn_partial.shape
    unit = C // units
    cpg = C // groups
    assert unit in (2, 4) and cpg % unit == 0
    s = gn_partial.view(B, slabs, groups, cpg // unit, 2).sum((1, 3))
    n = float(cpg * 32 * slabs)
    mean = s[..., 0] / n
    var = torch.clamp(s[..., 1] / n - mean * mean, min=0.0)
    if out is None:
        out = torch.empty((B, groups, 2), dtype=torch.float32)
    out[..., 0] = mean
    out[..., 1] = var
    return out


def groupnorm_apply_stats(x, mean_var, gamma, beta, groups, eps, silu, out=None):
    LAUNCHES[0] += 1
    B, C = x.shape[0], x.shape[-1]
    cpg = C // groups
    v = x.float().reshape(B, -1, groups, cpg)
    mean = mean_var[..., 0].view(B, 1, groups, 1)
    var = mean_var[..., 1].view(B, 1, groups, 1)
    y = (v - mean) * torch.rsqrt(var + eps)
    y = y.reshape(B, -1, C) * gamma.float() + beta.float()
    if silu:
        y = F.silu(y)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16)
    out.copy_(y.view(out.shape))
    return out


def conv3x3_supported(H, W, Cin):
    if Cin % 64 != 0:
        return False
    if W >= 128:
        return W % 128 == 0
    if W < 8 or 128 % W != 0:
        return False
    rows = 128 // W
    return H % rows == 0 if H >= rows else rows % H == 0


def layernorm(x, gamma, beta, eps, out=None, c_real=None):
    LAUNCHES[0] += 1
    C = x.shape[-1]
    if c_real is None or c_real == C:
        y = F.layer_norm(x.float(), (C,), gamma, beta, eps)
    else:
        y = torch.zeros(x.shape, dtype=torch.float32)
        y[..., :c_real] = F.layer_norm(x.float()[..., :c_real], (c_real,), gamma[:c_real], beta[:c_real], eps)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16)
    out.copy_(y.view(out.shape))
    return out


def softmax_rows(s, scale, out=None):
    LAUNCHES[0] += 1
    p = torch.softmax(s.float() * scale, dim=-1)
    if out is None:
        out = torch.empty(s.shape, dtype=torch.bfloat16)
    out.copy_(p)
    return out


def upsample2x(x, out=None):
    LAUNCHES[0] += 1
    B, H, W, C = x.shape
    y = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, C), dtype=torch.bfloat16)
    out.copy_(y)
    return out


def im2col(x, kh, kw, stride, pad_top, pad_left, Ho, Wo, out=None):
    LAUNCHES[0] += 1
    B, H, W, C = x.shape
    pb = max(0, (Ho - 1) * stride + kh - pad_top - H)
    pr = max(0, (Wo - 1) * stride + kw - pad_left - W)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (pad_left, pr, pad_top, pb))
    unf = F.unfold(xp, (kh, kw), stride=stride)  # [B, C*kh*kw, L]
    L = unf.shape[-1]
    assert L == Ho * Wo
    col = unf.view(B, C, kh * kw, L).permute(0, 3, 2, 1).reshape(B * L, kh * kw * C)
    if out is None:
        out = torch.empty(col.shape, dtype=torch.bfloat16)
    out.copy_(col)
    return out


def nchw_to_nhwc(x, out, coff=0):
    LAUNCHES[0] += 1
    C = x.shape[1]
    out[..., coff:coff + C].copy_(x.permute(0, 2, 3, 1))
    return out


def pointwise_nchw_to_nhwc(x, w, bias, scale, out, coff=0):
    LAUNCHES[0] += 1
    y = torch.einsum("oc,bchw->bhwo", w.float(), x.float() * scale)
    if bias is not None:
        y = y + bias.float()
    out[..., coff:coff + w.shape[0]].copy_(y)
    return out


def cast_bf16(x, out=None):
    LAUNCHES[0] += 1
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16)
    out.copy_(x)
    return out


def tile_blend(tiles, coords, weight, out):
    LAUNCHES[0] += 1
    out.zero_()
    th, tw = weight.shape
    for t in range(tiles.shape[0]):
        hi, wi = int(coords[t, 0]), int(coords[t, 1])
        out[..., hi:hi + th, wi:wi + tw] += tiles[t] * weight
    return out


def timestep_embedding(t, dim, max_period=10000.0, out=None):
    LAUNCHES[0] += 1
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    e = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if out is None:
        out = torch.empty((t.shape[0], dim), dtype=torch.bfloat16)
    out.copy_(e)
    return out


def sampler_update(x, eps, noise, index, tables, want_pred_x0=True, x_prev=None, pred_x0=None):
    LAUNCHES[0] += 1
    B = x.shape[0]
    e = lambda t: t[index].view(B, *([1] * (x.dim() - 1)))
    x0 = e(tables[0]) * x - e(tables[1]) * eps
    mean = e(tables[2]) * x0 + e(tables[3]) * x
    xp = mean + (index != 0).float().view(B, *([1] * (x.dim() - 1))) * torch.sqrt(e(tables[4])) * noise
    if x_prev is None:
        x_prev = torch.empty_like(x)
    x_prev.copy_(xp)
    if pred_x0 is None and want_pred_x0:
        pred_x0 = torch.empty_like(x)
    if pred_x0 is not None:
        pred_x0.copy_(x0)
    return x_prev, pred_x0


def pixel_unshuffle(x, out, mean, scale, r=8):
    LAUNCHES[0] += 1
    B, C, H, W = x.shape
    m = torch.tensor(list(mean)[:C], dtype=torch.float32).view(1, C, 1, 1)
    y = F.pixel_unshuffle((x - m) * scale, r).permute(0, 2, 3, 1)
    out[..., :C * r * r].copy_(y)
    return out


def window_attention(qkv, heads, shift, scale, bias, mask, out):
    """Swin W-MSA / SW-MSA on 32-wide (padded) heads, evaluated literally: roll, partition, attention, reverse."""
    LAUNCHES[0] += 1
    B, H, W, _ = qkv.shape
    C = heads * 32
    x = qkv.float()
    if shift:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    win = x.view(B, H // 8, 8, W // 8, 8, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, 64, 3, heads, 32).permute(2, 0, 3, 1, 4)
    q, k, v = win[0], win[1], win[2]                                   # [B*nW, heads, 64, 32]
    attn = (q @ k.transpose(-2, -1)) * scale + bias.view(1, heads, 64, 64)
    if mask is not None:
        nW = (H // 8) * (W // 8)
        attn = (attn.view(B, nW, heads, 64, 64) + mask.view(1, nW, 1, 64, 64)).view(-1, heads, 64, 64)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(B, H // 8, W // 8, 8, 8, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    out[..., :C].copy_(o)
    return out


def wavelet_reconstruction(content, style, levels=5):
    """utils/common.py:99-147 with plain torch ops (dilated 3x3 binomial blur, replicate padding)."""
    def blur(img, r):
        c = img.shape[1]
        k = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]], dtype=img.dtype)
        k = k[None, None].repeat(c, 1, 1, 1)
        return F.conv2d(F.pad(img, (r, r, r, r), mode="replicate"), k, groups=c, dilation=r)

    def decompose(img):
        high = torch.zeros_like(img)
        for i in range(levels):
            low = blur(img, 2 ** i)
            high = high + (img - low)
            img = low
        return high, img

    LAUNCHES[0] += 2 * levels
    return decompose(content)[0] + decompose(style)[1]
