"""Test-only torch stand-in for ``edtr_b200.ops32`` (the fp32-mode kernels): same function signatures and layouts
(fp32 channels-last rows views, tap-major 3x3 filters, in-place `out=` destinations), evaluated with plain PyTorch on
the CPU so that the fp32 engine's dataflow can be checked against the live-reference fixtures without a GPU."""
import contextlib

import torch
import torch.nn.functional as F

ACT_NONE, ACT_SILU = 0, 1
REQUIRES_CUDA = False
F32 = torch.float32


def device_guard(device):
    return contextlib.nullcontext()


def _finish(v, M, N, *, bias, rowvec, rows_per_group, residual, act, alpha, out, nchw_hw):
    v = v * alpha
    if bias is not None:
        v = v + bias
    if rowvec is not None:
        v = v + rowvec[:, :N].repeat_interleave(rows_per_group, 0)[:M]
    if residual is not None:
        v = v + residual.reshape(M, N)
    if act == ACT_SILU:
        v = F.silu(v)
    if nchw_hw:
        res = v.view(M // nchw_hw, nchw_hw, N).permute(0, 2, 1)
        if out is None:
            return res.contiguous()
        out.view(M // nchw_hw, N, nchw_hw).copy_(res)
        return out
    if out is None:
        return v
    assert out.dtype == F32 and out.numel() == M * N
    out.copy_(v.view(out.shape))
    return out


def gemm(a, w, *, bias=None, rowvec=None, rows_per_group=0, residual=None, act=ACT_NONE, alpha=1.0, out=None, nchw_hw=0):
    assert a.dtype == F32 and w.dtype == F32
    A = a.reshape(-1, a.shape[-1])
    return _finish(A @ w.t(), A.shape[0], w.shape[0], bias=bias, rowvec=rowvec, rows_per_group=rows_per_group,
                   residual=residual, act=act, alpha=alpha, out=out, nchw_hw=nchw_hw)


def conv3x3(x, w, *, stride=1, pad=(1, 1), out_hw=None, up2x=False, bias=None, rowvec=None, residual=None, act=ACT_NONE,
            alpha=1.0, out=None, nchw=False):
    assert x.dtype == F32 and x.dim() == 4
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == 9 * Cin
    wt = w.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    xi = x.permute(0, 3, 1, 2)
    if up2x:
        xi = F.interpolate(xi, scale_factor=2.0, mode="nearest")
        y = F.conv2d(xi, wt, None, padding=1)
    else:
        Ho, Wo = out_hw if out_hw is not None else ((H, W) if stride == 1 else
                                                    ((H + 2 * pad[0] - 3) // stride + 1, (W + 2 * pad[1] - 3) // stride + 1))
        # zero padding: `pad` on top / left, whatever the output grid needs on the bottom / right
        pb = max(0, (Ho - 1) * stride + 3 - pad[0] - H)
        pr = max(0, (Wo - 1) * stride + 3 - pad[1] - W)
        y = F.conv2d(F.pad(xi, (pad[1], pr, pad[0], pb)), wt, None, stride=stride)[:, :, :Ho, :Wo]
    Ho, Wo = y.shape[2], y.shape[3]
    y = y.permute(0, 2, 3, 1).reshape(-1, Cout)
    res = _finish(y, B * Ho * Wo, Cout, bias=bias, rowvec=rowvec, rows_per_group=Ho * Wo, residual=residual, act=act,
                  alpha=alpha, out=out, nchw_hw=Ho * Wo if nchw else 0)
    if out is None and not nchw:
        res = res.view(B, Ho, Wo, Cout)
    return res


def attention(q, k, v, heads, scale, out=None):
    B, Lq, C = q.shape
    d = C // heads
    sp = lambda t: t.reshape(B, -1, heads, d).permute(0, 2, 1, 3)
    o = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) * scale, -1) @ sp(v)
    o = o.permute(0, 2, 1, 3).reshape(B, Lq, C)
    if out is None:
        return o
    out.copy_(o)
    return out


def groupnorm(x, gamma, beta, groups, eps, silu, out=None):
    B, C = x.shape[0], x.shape[-1]
    y = F.group_norm(x.reshape(B, -1, C).permute(0, 2, 1), groups, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    y = y.permute(0, 2, 1).reshape(x.shape)
    if out is None:
        return y.contiguous()
    out.copy_(y)
    return out


def layernorm(x, gamma, beta, eps, out=None):
    y = F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)
    if out is None:
        return y
    out.copy_(y)
    return out


def geglu(x, out=None):
    a, g = x.chunk(2, dim=-1)
    y = a * F.gelu(g)
    if out is None:
        return y.contiguous()
    out.copy_(y)
    return out


def silu(x, out=None):
    y = F.silu(x)
    if out is None:
        return y
    out.copy_(y)
    return out


def nchw_to_nhwc(x, out, coff=0, scale=1.0):
    out[..., coff:coff + x.shape[1]] = x.permute(0, 2, 3, 1) * scale
    return out


def timestep_embedding(t, dim, max_period=10000.0, out=None):
    import math

    half = dim // 2
    f = torch.exp(-math.log(max_period) * torch.arange(half, dtype=F32) / half)
    a = t[:, None].float() * f[None]
    e = torch.cat([torch.cos(a), torch.sin(a)], -1)
    if out is None:
        return e
    out.copy_(e)
    return out
