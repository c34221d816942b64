"""GPU unit tests: every C-ABI kernel against a plain PyTorch fp32 evaluation of the same op
(inputs rounded to bf16 first, so only accumulation order and the final bf16 store differ)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def relerr(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


@pytest.fixture(scope="module")
def ops():
    from edtr_b200 import ops as o

    return o


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 256), (300, 320, 320), (8, 1280, 320),
                                    (512, 1280, 1280), (1024, 960, 320), (200, 64, 128), (4096, 320, 1280)])
def test_gemm_plain(ops, M, N, K):
    a = rnd(M, K, seed=1)
    w = rnd(N, K, scale=K ** -0.5, seed=2)
    out = ops.gemm(a, w)
    ref = a.float() @ w.float().t()
    assert relerr(out, ref) < 1e-2


def test_gemm_epilogue_bias_residual_rowvec_silu(ops):
    M, N, K = 512, 320, 640
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    res = rnd(M, N, seed=3)
    rowvec = torch.randn(4, N, device="cuda")
    out = ops.gemm(a, w, bias=bias, residual=res, rowvec=rowvec, rows_per_group=128)
    ref = a.float() @ w.float().t() + bias + res.float() + rowvec.repeat_interleave(128, 0)
    assert relerr(out, ref) < 1e-2
    out = ops.gemm(a, w, bias=bias, act=ops.ACT_SILU)
    assert relerr(out, F.silu(a.float() @ w.float().t() + bias)) < 1e-2


def test_gemm_strided_views_and_inplace_residual(ops):
    M, N, K = 256, 320, 320
    abuf = rnd(M, 960, seed=1)
    a = abuf[:, 320:640]
    w = rnd(N, K, scale=K ** -0.5, seed=2)
    cat = rnd(M, 640, seed=3)
    before = cat.clone()
    ops.gemm(a, w, residual=cat[:, 320:], out=cat[:, 320:])
    ref = a.float() @ w.float().t() + before[:, 320:].float()
    assert relerr(cat[:, 320:], ref) < 1e-2
    assert torch.equal(cat[:, :320], before[:, :320])


def test_gemm_geglu(ops):
    from edtr_b200 import lib

    M, C = 384, 320
    a = rnd(M, C, seed=1)
    w = rnd(8 * C, C, scale=C ** -0.5, seed=2)
    bias = torch.randn(8 * C, device="cuda") * 0.1
    bn = lib.device_lib().edtr_gemm_tile_n(M, 8 * C, C, ops.ACT_GEGLU)
    half = bn // 2
    n_half = 4 * C
    assert n_half % half == 0
    idx = torch.arange(n_half, device="cuda").view(-1, half)
    perm = torch.cat([idx, idx + n_half], dim=1).reshape(-1)
    out = ops.gemm(a, w[perm].contiguous(), bias=bias[perm].contiguous(), act=ops.ACT_GEGLU)
    y = a.float() @ w.float().t() + bias
    x, gate = y.chunk(2, dim=-1)
    assert relerr(out, x * F.gelu(gate)) < 1e-2


@pytest.mark.parametrize("M,N,K", [(20000, 320, 320), (32768, 960, 320), (8192, 640, 2560), (2048, 1280, 1280),
                                    (512, 1280, 5120), (33000, 128, 1152), (4096, 1920, 640)])
def test_gemm_persistent_pair_kernel(ops, M, N, K):
    """Shapes with several tiles per CTA pair (persistence, TMEM double buffering, ragged last column tile)."""
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    res = rnd(M, N, seed=3)
    out = ops.gemm(a, w, bias=bias, residual=res)
    ref = a.float() @ w.float().t() + bias + res.float()
    assert relerr(out, ref) < 1e-2
    out = ops.gemm(a, w)
    assert relerr(out, a.float() @ w.float().t()) < 1e-2


def test_gemm_geglu_large(ops):
    from edtr_b200 import lib

    M, C = 8192, 320
    a = rnd(M, C, seed=1)
    w = rnd(8 * C, C, scale=C ** -0.5, seed=2)
    bias = torch.randn(8 * C, device="cuda") * 0.1
    bn = lib.device_lib().edtr_gemm_tile_n(M, 8 * C, C, ops.ACT_GEGLU)
    half, n_half = bn // 2, 4 * C
    idx = torch.arange(n_half, device="cuda").view(-1, half)
    perm = torch.cat([idx, idx + n_half], dim=1).reshape(-1)
    out = ops.gemm(a, w[perm].contiguous(), bias=bias[perm].contiguous(), act=ops.ACT_GEGLU)
    y = a.float() @ w.float().t() + bias
    x, gate = y.chunk(2, dim=-1)
    assert relerr(out, x * F.gelu(gate)) < 1e-2


def test_conv3x3_large_batch(ops):
    B, H, W, Cin, Cout = 8, 64, 64, 320, 320
    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    bias = torch.randn(Cout, device="cuda")
    emb = torch.randn(B, Cout, device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    cat = rnd(B, H, W, 640, seed=3)
    before = cat.clone()
    ops.conv3x3(x, wp, bias=bias, rowvec=emb, residual=cat[..., 320:], out=cat[..., 320:])
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + emb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1) + before[..., 320:].float()
    assert relerr(cat[..., 320:], ref) < 1e-2
    assert torch.equal(cat[..., :320], before[..., :320])


def test_gemm_out_modes(ops):
    M, N, K, hw = 512, 4, 320, 256
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    ref = a.float() @ w.float().t() + bias
    o32 = ops.gemm(a, w, bias=bias, out_mode=ops.OUT_F32)
    assert relerr(o32, ref) < 1e-3
    on = ops.gemm(a, w, bias=bias, out_mode=ops.OUT_NCHW_F32, hw=hw)
    assert relerr(on, ref.view(M // hw, hw, N).permute(0, 2, 1)) < 1e-3
    w2 = rnd(192, K, scale=K ** -0.5, seed=4)
    ob = ops.gemm(a, w2, out_mode=ops.OUT_NCHW_BF16, hw=hw)
    assert relerr(ob, (a.float() @ w2.float().t()).view(M // hw, hw, 192).permute(0, 2, 1)) < 1e-2


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 64, 128), (1, 8, 8, 128, 64), (3, 8, 8, 64, 64),
                                             (2, 16, 16, 320, 320), (1, 32, 32, 640, 320), (1, 128, 128, 64, 64),
                                             (1, 256, 256, 128, 3), (8, 8, 8, 1280, 1280), (1, 64, 64, 64, 4),
                                             (3, 32, 32, 320, 4), (2, 128, 128, 128, 3), (5, 16, 16, 512, 8)])
def test_conv3x3(ops, B, H, W, Cin, Cout):
    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    bias = torch.randn(Cout, device="cuda")
    emb = torch.randn(B, Cout, device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    if Cout % 64 == 0:
        res = rnd(B, H, W, Cout, seed=3)
        out = ops.conv3x3(x, wp, bias=bias, rowvec=emb, residual=res.view(-1, Cout))
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + emb[:, :, None, None]
        ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout) + res.view(-1, Cout).float()
        assert relerr(out, ref) < 1e-2
    else:
        out = ops.conv3x3(x, wp, bias=bias, out_mode=ops.OUT_NCHW_F32)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
        assert relerr(out.view(B, Cout, H, W), ref) < 1e-2


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(8, 8, 8, 128, 128), (2, 16, 16, 320, 256), (1, 32, 32, 64, 64),
                                             (1, 64, 64, 128, 64), (1, 128, 128, 64, 64), (3, 16, 32, 64, 128)])
def test_conv3x3_up2x(ops, B, H, W, Cin, Cout):
    """Phase-decomposed nearest-2x + conv3x3 vs F.interpolate + F.conv2d."""
    from edtr_b200.engine import pack_conv3x3_up2x

    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    bias = torch.randn(Cout, device="cuda")
    w4 = pack_conv3x3_up2x(w.float(), "cuda")
    cat = rnd(B, 2 * H, 2 * W, Cout + 64, seed=3)
    before = cat.clone()
    ops.conv3x3_up2x(x, w4, bias=bias, out=cat[..., :Cout])
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, w.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert relerr(cat[..., :Cout], ref) < 1.5e-2
    assert torch.equal(cat[..., Cout:], before[..., Cout:])


def test_conv3x3_channel_slice_input(ops):
    B, H, W = 2, 16, 16
    buf = rnd(B, H, W, 192, seed=1)
    x = buf[..., 64:192]
    w = rnd(64, 128, 3, 3, scale=(9 * 128) ** -0.5, seed=2)
    wp = w.permute(0, 2, 3, 1).reshape(64, -1).contiguous()
    out = ops.conv3x3(x, wp)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, 64)
    assert relerr(out, ref) < 1e-2


@pytest.mark.parametrize("B,heads,Lq,Lk", [(2, 5, 256, 256), (1, 2, 4096, 4096), (2, 3, 128, 77), (1, 20, 64, 64),
                                            (1, 1, 200, 130), (2, 2, 1024, 77), (1, 2, 130, 192), (2, 1, 300, 290),
                                            (1, 3, 64, 1), (1, 1, 128, 65),
                                            # short key sequences run cross_attention_small_kernel (Lk <= 80)
                                            (8, 5, 4096, 77), (3, 20, 200, 77), (1, 10, 37, 80), (2, 2, 1000, 33),
                                            (1, 1, 16, 8), (2, 4, 129, 79), (1, 2, 300, 81)])
def test_attention(ops, B, heads, Lq, Lk):
    C = heads * 64
    qkv = rnd(B, Lq, 3 * C, seed=1)
    if Lk == Lq:
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = qkv[..., :C]
        kv = rnd(B, Lk, 2 * C, seed=2)
        k, v = kv[..., :C], kv[..., C:]
    out = ops.attention(q, k, v, heads, 0.125)

    def split(t):
        return t.float().view(B, -1, heads, 64).permute(0, 2, 1, 3)

    ref = F.scaled_dot_product_attention(split(q), split(k), split(v)).permute(0, 2, 1, 3).reshape(B, Lq, C)
    assert relerr(out, ref) < 2e-2


@pytest.mark.parametrize("Lq,Lk", [(256, 1024), (384, 200), (128, 4096), (128, 320)])
def test_attention_growing_scores_exercise_lazy_rescale(ops, Lq, Lk):
    """Scores that keep growing along the key axis force the in-TMEM rescale of O many times per row."""
    B, heads = 2, 2
    C = heads * 64
    q = rnd(B, Lq, C, seed=1, scale=2.0)
    ramp = torch.linspace(0.5, 4.0, Lk, device="cuda").view(1, Lk, 1)
    k = (rnd(B, Lk, C, seed=2).float() * ramp).to(BF)
    v = rnd(B, Lk, C, seed=3)
    out = ops.attention(q, k, v, heads, 0.125)

    def split(t):
        return t.float().view(B, -1, heads, 64).permute(0, 2, 1, 3)

    ref = F.scaled_dot_product_attention(split(q), split(k), split(v)).permute(0, 2, 1, 3).reshape(B, Lq, C)
    assert relerr(out, ref) < 2e-2


@pytest.mark.parametrize("B,HW,C,silu,eps", [(2, 4096, 320, True, 1e-5), (2, 64, 2560, True, 1e-5),
                                              (1, 1024, 1920, False, 1e-6), (3, 256, 960, True, 1e-5),
                                              (1, 65536, 128, True, 1e-6), (2, 4096, 512, True, 1e-6),
                                              (8, 4096, 320, True, 1e-5), (8, 1024, 640, True, 1e-6), (3, 4096, 960, True, 1e-5),
                                              (8, 256, 1280, True, 1e-5), (2, 1000, 64, False, 1e-5), (1, 77, 2560, True, 1e-5),
                                              (8, 1024, 320, True, 1e-5), (5, 256, 1920, False, 1e-6), (2, 262144, 128, True, 1e-6),
                                              (8, 1024, 1280, True, 1e-5), (1, 4096, 320, False, 1e-6), (3, 1024, 640, True, 1e-5)])
@pytest.mark.parametrize("fused", [True, False])
def test_groupnorm(ops, B, HW, C, silu, eps, fused):
    """fused=True: single-launch cluster kernel where eligible (L2-resident tensors); False: two-pass kernels."""
    x = (rnd(B, HW, C, seed=1).float() * 1.5 + 0.7).to(BF)
    gamma = torch.randn(C, device="cuda")
    beta = torch.randn(C, device="cuda")
    old = ops.GROUPNORM_FUSED
    ops.GROUPNORM_FUSED = fused
    try:
        out = ops.groupnorm(x, gamma, beta, 32, eps, silu)
    finally:
        ops.GROUPNORM_FUSED = old
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    assert relerr(out, ref.permute(0, 2, 1)) < 1e-2


@pytest.mark.parametrize("B,H,W,C", [(2, 38, 24, 128), (1, 86, 75, 64), (3, 16, 16, 512)])
def test_groupnorm_pool_and_apply_stats(ops, B, H, W, C):
    """Tiled-VAE GroupNorm: pooled (mean, var) over two tiles with pixel weights, then apply with given statistics
    (GroupNormParam.summary + custom_group_norm, utils/tilevae/tilevae.py:188-215, 263-278)."""
    tiles = [(rnd(B, H, W, C, seed=1).float() * 1.3 + 0.4).to(BF), (rnd(B, H // 2, W, C, seed=2).float() * 0.7 - 0.2).to(BF)]
    pix = [t.shape[1] * t.shape[2] for t in tiles]
    acc = torch.zeros(B, 32, 2, device="cuda")
    ref = torch.zeros(B, 32, 2, device="cuda")
    for t, p in zip(tiles, pix):
        wgt = p / sum(pix)
        ops.groupnorm_pool(t, 32, wgt, acc)
        v = t.float().reshape(B, -1, 32, C // 32).permute(0, 2, 1, 3).reshape(B, 32, -1)
        var, mean = torch.var_mean(v, dim=2, unbiased=False)
        ref[..., 0] += wgt * mean
        ref[..., 1] += wgt * var
    assert relerr(acc, ref) < 1e-4
    gamma, beta = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    for silu in (False, True):
        out = ops.groupnorm_apply_stats(tiles[0], acc, gamma, beta, 32, 1e-6, silu)
        x = tiles[0].float().reshape(B, -1, 32, C // 32)
        y = (x - ref[..., 0].view(B, 1, 32, 1)) * torch.rsqrt(ref[..., 1].view(B, 1, 32, 1) + 1e-6)
        y = y.reshape(B, H, W, C) * gamma + beta
        if silu:
            y = F.silu(y)
        assert relerr(out, y) < 1e-2


def test_groupnorm_on_channel_slice(ops):
    B, HW = 2, 256
    buf = rnd(B, HW, 640, seed=1)
    x = buf[..., :320]
    gamma, beta = torch.randn(320, device="cuda"), torch.randn(320, device="cuda")
    out = ops.groupnorm(x, gamma, beta, 32, 1e-5, False)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, 1e-5).permute(0, 2, 1)
    assert relerr(out, ref) < 1e-2


@pytest.mark.parametrize("M,C", [(1000, 320), (512, 640), (300, 1280)])
def test_layernorm(ops, M, C):
    x = (rnd(M, C, seed=1).float() * 2 + 0.3).to(BF)
    gamma, beta = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    out = ops.layernorm(x, gamma, beta, 1e-5)
    assert relerr(out, F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)) < 1e-2


def test_softmax_rows(ops):
    s = torch.randn(300, 4096, device="cuda") * 20
    out = ops.softmax_rows(s, 0.044)
    assert relerr(out, torch.softmax(s * 0.044, -1)) < 1e-2


def test_upsample_im2col_layout(ops):
    x = rnd(2, 8, 16, 64, seed=1)
    up = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    # stride-2 pad-1 3x3 gather == unfold
    col = ops.im2col(x, 3, 3, 2, 1, 1, 4, 8)
    unf = F.unfold(x.float().permute(0, 3, 1, 2), 3, padding=1, stride=2)  # [B, C*9, L]
    unf = unf.view(2, 64, 9, -1).permute(0, 3, 2, 1).reshape(2 * 32, 9 * 64)
    assert torch.equal(col.float(), unf)
    # asymmetric pad (0,1,0,1), stride 2, no top/left pad (VAE encoder downsample)
    col = ops.im2col(x, 3, 3, 2, 0, 0, 4, 8)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    unf = F.unfold(xp, 3, stride=2).view(2, 64, 9, -1).permute(0, 3, 2, 1).reshape(2 * 32, 9 * 64)
    assert torch.equal(col.float(), unf)
    # NCHW fp32 -> channels-last bf16 at a channel offset, and back
    src = torch.randn(2, 4, 8, 8, device="cuda")
    dst = torch.zeros(2, 8, 8, 64, dtype=BF, device="cuda")
    ops.nchw_to_nhwc(src, dst, coff=4)
    assert torch.equal(dst[..., 4:8].float(), src.to(BF).float().permute(0, 2, 3, 1))
    assert dst[..., :4].abs().max() == 0 and dst[..., 8:].abs().max() == 0
    back = ops.nhwc_to_nchw(dst.view(-1, 64)[:, 4:8], 2)
    assert torch.equal(back.view(2, 4, 8, 8), src.to(BF).float())
    t16 = ops.nhwc_to_nchw(x.view(2, 128, 64), 2, out_f32=False)
    assert torch.equal(t16, x.view(2, 128, 64).permute(0, 2, 1))
    assert torch.equal(ops.cast_bf16(src), src.to(BF))


def test_timestep_embedding(ops):
    t = torch.tensor([200, 150, 100, 50, 0, 999], device="cuda")
    out = ops.timestep_embedding(t, 320)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert (out.float() - ref).abs().max() < 1e-2


def test_sampler_update(ops):
    B = 3
    x, eps, noise = (torch.randn(B, 4, 64, 64, device="cuda") for _ in range(3))
    idx = torch.tensor([3, 0, 1], device="cuda")
    tabs = [torch.rand(4, device="cuda") + 0.1 for _ in range(5)]
    xp, x0 = ops.sampler_update(x, eps, noise, idx, tabs)
    e = lambda t: t[idx].view(B, 1, 1, 1)
    rx0 = e(tabs[0]) * x - e(tabs[1]) * eps
    mean = e(tabs[2]) * rx0 + e(tabs[3]) * x
    ref = mean + (idx != 0).float().view(B, 1, 1, 1) * torch.sqrt(e(tabs[4])) * noise
    assert torch.allclose(x0, rx0, atol=1e-6) and torch.allclose(xp, ref, atol=1e-6)


def test_validation_errors(ops):
    a = rnd(128, 100, seed=1)
    w = rnd(64, 100, seed=2)
    with pytest.raises(ValueError):
        ops.gemm(a, w)  # K not a multiple of 64
    with pytest.raises(ValueError):
        ops.gemm(rnd(128, 64).float(), rnd(64, 64))
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(128, 64, dtype=BF), torch.zeros(64, 64, dtype=BF))


def test_wavelet_reconstruction_against_reference_fixture(ops):
    """The colour fix after the decode (utils/common.py:136-147) vs the fixture recorded from the live reference,
    and vs the same arithmetic in torch at 512x512."""
    import os

    import numpy as np

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_wavelet.npz"))
    out = ops.wavelet_reconstruction(torch.from_numpy(g["content"]).cuda(), torch.from_numpy(g["style"]).cuda())
    assert relerr(out.cpu(), torch.from_numpy(g["out"])) < 1e-5
    from oracle import cldm_oracle as O

    gen = torch.Generator().manual_seed(3)
    c, s = torch.rand(2, 3, 512, 512, generator=gen), torch.rand(2, 3, 512, 512, generator=gen)
    assert relerr(ops.wavelet_reconstruction(c.cuda(), s.cuda()).cpu(), O.wavelet_reconstruction(c, s)) < 1e-5


# ------------------------------------------------------------------------------------------ SwinIR kernels (swin.cu)
def _fake():
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fake_ops

    return fake_ops


@pytest.mark.parametrize("B,H,W,heads,shift", [(2, 16, 16, 6, 0), (2, 16, 16, 6, 4), (1, 8, 24, 2, 4), (3, 64, 64, 6, 4),
                                               (1, 8, 8, 2, 0)])
def test_window_attention(ops, B, H, W, heads, shift):
    """W-MSA / SW-MSA on 32-wide padded heads (two zero columns per head) against the literal roll / partition /
    softmax / reverse evaluation."""
    from edtr_b200.swinir import shifted_window_mask

    C = heads * 32
    qkv = rnd(B, H, W, 3 * C, seed=1)
    qkv.view(B, H, W, 3, heads, 32)[..., 30:] = 0
    bias = torch.randn(heads, 64, 64, device="cuda") * 0.5
    mask = shifted_window_mask(H, W, 8, shift).contiguous().cuda() if shift else None
    out = torch.zeros(B, H, W, C, dtype=BF, device="cuda")
    ops.window_attention(qkv, heads, shift, 30 ** -0.5, bias, mask, out)
    ref = torch.zeros(B, H, W, C, dtype=BF)
    _fake().window_attention(qkv.cpu(), heads, shift, 30 ** -0.5, bias.cpu(), None if mask is None else mask.cpu(), ref)
    assert relerr(out.cpu(), ref) < 2e-2
    assert float(out.view(B, H, W, heads, 32)[..., 30:].abs().max()) == 0.0


@pytest.mark.parametrize("M,C,c_real", [(1000, 192, 180), (512, 64, 60), (300, 320, 320)])
def test_layernorm_padded(ops, M, C, c_real):
    x = rnd(M, C, seed=1)
    x[:, c_real:] = 0
    gamma, beta = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    gamma[c_real:] = 0
    beta[c_real:] = 0
    out = ops.layernorm(x, gamma, beta, 1e-5, c_real=c_real)
    ref = F.layer_norm(x.float()[:, :c_real], (c_real,), gamma[:c_real], beta[:c_real], 1e-5)
    assert relerr(out[:, :c_real], ref) < 1e-2
    assert float(out[:, c_real:].abs().max()) == 0.0 if c_real < C else True


def test_pixel_unshuffle(ops):
    x = torch.rand(2, 3, 128, 192, device="cuda")
    mean = (0.4488, 0.4371, 0.4040)
    out = torch.empty(2, 16, 24, 192, dtype=BF, device="cuda")
    ops.pixel_unshuffle(x, out, mean, 1.0, 8)
    m = torch.tensor(mean, device="cuda").view(1, 3, 1, 1)
    ref = F.pixel_unshuffle(x - m, 8).permute(0, 2, 3, 1)
    assert relerr(out, ref) < 1e-2


@pytest.mark.parametrize("act", [3, 4, 5])
@pytest.mark.parametrize("M,N,K", [(2048, 384, 192), (64, 128, 64)])
def test_gemm_pointwise_activations(ops, act, M, N, K):
    """GELU (erf) / LeakyReLU(0.2) / LeakyReLU(0.01) epilogues on both tensor-core kernels."""
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    v = a.float() @ w.float().t() + bias
    ref = F.gelu(v) if act == 3 else F.leaky_relu(v, 0.2 if act == 4 else 0.01)
    assert relerr(ops.gemm(a, w, bias=bias, act=act), ref) < 1e-2


def test_conv3x3_up2x_leaky_relu(ops):
    B, H, W, Cin, Cout = 2, 16, 16, 64, 64
    from edtr_b200.engine import pack_conv3x3_up2x

    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    bias = torch.randn(Cout, device="cuda")
    out = ops.conv3x3_up2x(x, pack_conv3x3_up2x(w.float(), "cuda"), bias=bias, act=ops.ACT_LRELU_02)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.leaky_relu(F.conv2d(up, w.float(), bias, padding=1), 0.2).permute(0, 2, 3, 1)
    assert relerr(out, ref) < 1e-2


# ------------------------------------------------------------------ round 2: epilogue statistics / folded LayerNorm / in-kernel split-K
@pytest.mark.parametrize("M,N,K,res", [(32768, 320, 320, True), (2048, 1280, 1280, True), (512, 1280, 1280, False),
                                        (128, 640, 640, True), (300, 320, 320, False)])
def test_gemm_row_stats(ops, M, N, K, res):
    """EdtrEpilogue.row_stats: per-row partial (sum, sum of squares) of the stored matrix, one pair per column tile and
    epilogue warp group, on the CTA-pair kernel (M >= 256) and on the single-CTA kernel (M < 256); the parts of a row
    add up to the row's totals."""
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    r = rnd(M, N, seed=3) if res else None
    parts = ops.row_stats_parts(M, N, K)
    assert 1 <= parts <= 2 * ((N + 63) // 64) or M < 256
    rs = torch.full((M, parts, 2), float("nan"), device="cuda")
    out = ops.gemm(a, w, bias=bias, residual=r, row_stats=rs)
    ref = a.float() @ w.float().t() + bias + (r.float() if res else 0)
    assert relerr(out, ref) < 1e-2
    assert not torch.isnan(rs).any()
    assert relerr(rs[..., 0].sum(1), ref.sum(-1)) < 2e-3
    assert relerr(rs[..., 1].sum(1), (ref * ref).sum(-1)) < 2e-3
    with pytest.raises(ValueError):
        ops.gemm(a, w, row_stats=torch.empty((M, parts + 1, 2), device="cuda"))


@pytest.mark.parametrize("M,C,N", [(32768, 320, 960), (2048, 1280, 1280), (512, 1280, 3840), (128, 640, 640), (64, 1280, 1280)])
def test_gemm_folded_layernorm(ops, M, C, N):
    """LayerNorm folded into the consuming GEMM (engine.fold_layernorm + EdtrEpilogue.ln_*) vs F.layer_norm + matmul;
    rows get a large common offset so that the mean-cancellation term matters."""
    from edtr_b200.engine import fold_layernorm

    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn(M, C, generator=g, device="cuda") * 1.5 + 3.0 * torch.randn(M, 1, generator=g, device="cuda")).to(BF)
    w = torch.randn(N, C, generator=g, device="cuda") * C ** -0.5
    b = torch.randn(N, generator=g, device="cuda") * 0.1
    gamma = 1 + 0.3 * torch.randn(C, generator=g, device="cuda")
    beta = 0.2 * torch.randn(C, generator=g, device="cuda")
    wg, cs, bf = fold_layernorm(w, b, gamma, beta, "cuda")
    xf = x.float().view(M, C // 32, 32)
    stats = torch.stack([xf.sum(-1), (xf * xf).sum(-1)], -1).contiguous()
    out = ops.gemm(x, wg, bias=bf, ln=(stats, C, 1e-5, cs))
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5) @ w.t() + b
    assert relerr(out, ref) < 1.5e-2


def test_gemm_folded_layernorm_geglu_and_producer_chain(ops):
    """Producer GEMM (+residual) emits the row statistics, the GEGLU projection consumes them: the dataflow of
    BasicTransformerBlock's `ff(norm3(x)) + x` (model/attention.py:233) without a LayerNorm kernel."""
    from edtr_b200.engine import fold_layernorm, geglu_permutation

    M, C = 4096, 320
    g = torch.Generator(device="cuda").manual_seed(6)
    a = rnd(M, C, seed=1)
    w0 = rnd(C, C, scale=C ** -0.5, seed=2)
    res = rnd(M, C, seed=3)
    rs = torch.empty((M, ops.row_stats_parts(M, C, C), 2), device="cuda")
    t = ops.gemm(a, w0, residual=res, row_stats=rs)          # stored bf16 residual stream + its statistics
    w = torch.randn(8 * C, C, generator=g, device="cuda") * C ** -0.5
    b = torch.randn(8 * C, generator=g, device="cuda") * 0.1
    gamma = 1 + 0.3 * torch.randn(C, generator=g, device="cuda")
    beta = 0.2 * torch.randn(C, generator=g, device="cuda")
    perm = geglu_permutation(4 * C, ops.geglu_tile_n()).cuda()
    wg, cs, bf = fold_layernorm(w[perm], b[perm], gamma, beta, "cuda")
    out = ops.gemm(t, wg, bias=bf, ln=(rs, C, 1e-5, cs), act=ops.ACT_GEGLU)
    y = F.layer_norm(t.float(), (C,), gamma, beta, 1e-5) @ w.t() + b
    xh, gate = y.chunk(2, dim=-1)
    assert relerr(out, xh * F.gelu(gate)) < 1.5e-2


@pytest.mark.parametrize("M,N,K", [(512, 1280, 11520), (512, 1280, 5120), (2048, 1280, 11520), (256, 640, 5760)])
def test_gemm_split_k(ops, M, N, K):
    """Under-filled problems split K into the per-call workspace (EdtrEpilogue.workspace); the partial tiles are summed
    in split order (bit-reproducible) with bias / rowvec / residual / SiLU applied by the reduce pass."""
    if M <= 512:
        assert ops.gemm_workspace_size(M, N, K) > 0      # the planner does split these
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    res = rnd(M, N, seed=3)
    rowvec = torch.randn(M // 64, N, device="cuda")
    ref = a.float() @ w.float().t() + bias + res.float() + rowvec.repeat_interleave(64, 0)
    outs = [ops.gemm(a, w, bias=bias, residual=res, rowvec=rowvec, rows_per_group=64) for _ in range(3)]
    assert relerr(outs[0], ref) < 1e-2
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    out = ops.gemm(a, w, bias=bias, act=ops.ACT_SILU)
    assert relerr(out, F.silu(a.float() @ w.float().t() + bias)) < 1e-2
    # in-place accumulation (zero-conv into the skip tensor) through the split path
    acc = res.clone()
    ops.gemm(a, w, bias=bias, residual=acc, out=acc, alpha=0.5)
    assert relerr(acc, 0.5 * (a.float() @ w.float().t()) + bias + res.float()) < 1e-2


def test_conv3x3_split_k(ops):
    B, H, W, Cin, Cout = 8, 8, 8, 1280, 1280
    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    bias = torch.randn(Cout, device="cuda")
    emb = torch.randn(B, Cout, device="cuda")
    out = ops.conv3x3(x, wp, bias=bias, rowvec=emb)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + emb[:, :, None, None]
    assert relerr(out.view(B, H, W, Cout), ref.permute(0, 2, 3, 1)) < 1e-2


# ------------------------------------------------- GroupNorm statistics from the GEMM / convolution epilogue
def _gn_from_partial(ops, y_bf16, part, gamma, beta, eps, silu):
    mv = ops.groupnorm_fold(part, y_bf16.shape[-1], 32)
    return ops.groupnorm_apply_stats(y_bf16, mv, gamma, beta, 32, eps, silu), mv


def _check_partial(part, ref_rows, B, HW, C):
    """part [B, HW/32, C/unit, 2] vs the fp32 reference rows [B*HW, C] in launch-row order."""
    unit = C // part.shape[2]
    assert unit in (2, 4)
    u = ref_rows.view(B, HW // 32, 32, C // unit, unit)
    s1, s2 = u.sum((2, 4)), (u * u).sum((2, 4))
    assert (part[..., 0] - s1).abs().max().item() < 2e-2 * s1.abs().max().item() + 1e-3
    assert (part[..., 1] - s2).abs().max().item() < 2e-2 * s2.abs().max().item() + 1e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,res", [(2, 64, 64, 128, 128, True), (1, 128, 128, 256, 128, False),
                                                 (4, 32, 32, 512, 512, True), (3, 64, 64, 64, 256, False),
                                                 (8, 16, 16, 128, 192, True), (2, 64, 64, 320, 320, True),
                                                 (2, 32, 32, 320, 640, False), (8, 8, 8, 1280, 1280, True),
                                                 (2, 16, 16, 64, 960, False)])
def test_conv3x3_groupnorm_partials_and_fold(ops, B, H, W, Cin, Cout, res):
    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    bias = torch.randn(Cout, device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    r = rnd(B, H, W, Cout, seed=3) if res else None
    HW = H * W
    part = torch.full(ops.gn_partial_shape(B, HW, Cout), float("nan"), device="cuda")
    out = ops.conv3x3(x, wp, bias=bias, residual=None if r is None else r.view(-1, Cout), gn_partial=part)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    if r is not None:
        ref = ref + r.view(-1, Cout).float()
    assert relerr(out, ref) < 1e-2
    assert not torch.isnan(part).any()          # every (slab, unit) entry is written exactly once
    _check_partial(part, ref, B, HW, Cout)
    if (Cout // 32) % 2 == 0:
        gamma, beta = torch.randn(Cout, device="cuda"), torch.randn(Cout, device="cuda")
        y, mv = _gn_from_partial(ops, out.view(B, H, W, Cout), part, gamma, beta, 1e-6, True)
        g = ref.view(B, HW, 32, Cout // 32)
        assert (mv[..., 0] - g.mean((1, 3))).abs().max().item() < 1e-3
        assert relerr(mv[..., 1], g.var((1, 3), unbiased=False)) < 1e-3
        yref = F.silu(F.group_norm(ref.view(B, HW, Cout).permute(0, 2, 1), 32, gamma, beta, 1e-6)).permute(0, 2, 1)
        assert relerr(y.view(B, HW, Cout), yref) < 2e-2
        # and the same numbers as the statistics-pass GroupNorm of the stored tensor, up to bf16 rounding of its input
        y2 = ops.groupnorm(out.view(B, H, W, Cout), gamma, beta, 32, 1e-6, True)
        assert relerr(y, y2) < 2e-2


def test_gemm_groupnorm_partials(ops):
    B, HW, K, N = 4, 1024, 512, 512
    a, w = rnd(B * HW, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    r = rnd(B * HW, N, seed=3)
    part = torch.full(ops.gn_partial_shape(B, HW, N), float("nan"), device="cuda")
    out = ops.gemm(a, w, bias=bias, residual=r, gn_partial=part, gn_hw=HW)
    ref = a.float() @ w.float().t() + bias + r.float()
    assert relerr(out, ref) < 1e-2
    assert not torch.isnan(part).any()
    _check_partial(part, ref, B, HW, N)
    assert ops.gn_partial_supported(B * HW, HW, N, K) in (True, False)
    with pytest.raises(Exception):      # slabs must not straddle images
        ops.gemm(a, w, gn_partial=part, gn_hw=1000)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 32, 128, 128), (1, 64, 64, 256, 256), (8, 16, 16, 512, 512)])
def test_conv3x3_up2x_groupnorm_partials(ops, B, H, W, Cin, Cout):
    from edtr_b200.engine import pack_conv3x3_up2x

    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    bias = torch.randn(Cout, device="cuda")
    w4 = pack_conv3x3_up2x(w.float(), "cuda")
    HW = 4 * H * W
    part = torch.full(ops.gn_partial_shape(B, HW, Cout), float("nan"), device="cuda")
    out = ops.conv3x3_up2x(x, w4, bias=bias, gn_partial=part)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, w.float(), bias, padding=1).permute(0, 2, 3, 1)      # [B, 2H, 2W, C]
    assert relerr(out, ref) < 1.5e-2
    assert not torch.isnan(part).any()
    # slabs are phase-major, 32 consecutive low-resolution pixels each: the per-(image, group) totals are what matters
    gamma, beta = torch.randn(Cout, device="cuda"), torch.randn(Cout, device="cuda")
    y, mv = _gn_from_partial(ops, out, part, gamma, beta, 1e-6, False)
    g = ref.reshape(B, HW, 32, Cout // 32)
    assert (mv[..., 0] - g.mean((1, 3))).abs().max().item() < 1e-3
    assert relerr(mv[..., 1], g.var((1, 3), unbiased=False)) < 1e-3
    yref = F.group_norm(ref.reshape(B, HW, Cout).permute(0, 2, 1), 32, gamma, beta, 1e-6).permute(0, 2, 1)
    assert relerr(y.view(B, HW, Cout), yref) < 2e-2
    ph = ref[:, 1::2, 0::2, :].reshape(B, H * W // 32, 32, Cout // 4, 4)      # phase (py, px) = (1, 0) -> slab block 2
    blk = part[:, 2 * (H * W // 32):3 * (H * W // 32)]
    assert (blk[..., 0] - ph.sum((2, 4))).abs().max().item() < 2e-2 * ph.sum((2, 4)).abs().max().item() + 1e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,res", [(1, 128, 128, 128, 128, True), (2, 16, 256, 64, 128, False),
                                                 (1, 8, 512, 128, 64, True), (3, 5, 128, 256, 128, False),
                                                 (1, 64, 512, 128, 3, False), (1, 3, 384, 64, 64, True)])
def test_conv3x3_halo_mode(ops, B, H, W, Cin, Cout, res):
    """Narrow convolutions at W % 128 == 0 run the kernel's halo mode (one 130-pixel row segment per stage, the three
    dx taps through shifted A descriptors): borders (x = -1, x = W, y = -1, y = H), several channel blocks, odd H."""
    x = rnd(B, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    bias = torch.randn(Cout, device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
    if Cout % 64 == 0:
        r = rnd(B, H, W, Cout, seed=3) if res else None
        out = ops.conv3x3(x, wp, bias=bias, residual=None if r is None else r.view(-1, Cout))
        ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
        if r is not None:
            ref = ref + r.view(-1, Cout).float()
        assert relerr(out, ref) < 1e-2
    else:
        out = ops.conv3x3(x, wp, bias=bias, out_mode=ops.OUT_NCHW_F32)
        assert relerr(out.view(B, Cout, H, W), ref) < 1e-2


# ------------------------------------------------- fp32 mode kernels (edtr_f32_*), against fp32 / fp64 PyTorch
@pytest.fixture(scope="module")
def ops32():
    from edtr_b200 import ops32 as o

    return o


def frnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda") * scale


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (300, 320, 320), (8, 1280, 320), (4096, 77, 64), (1000, 4, 36), (130, 70, 50)])
def test_f32_gemm_epilogue(ops32, M, N, K):
    a, w = frnd(M, K, seed=1), frnd(N, K, scale=K ** -0.5, seed=2)
    bias, res = frnd(N, seed=3), frnd(M, N, seed=4)
    ref = (a.double() @ w.double().t()).float()
    assert relerr(ops32.gemm(a, w), ref) < 2e-6
    rpg = M // 2 if M % 2 == 0 else M
    rowvec = frnd(M // rpg, N, seed=5)
    out = ops32.gemm(a, w, bias=bias, residual=res, rowvec=rowvec, rows_per_group=rpg, act=ops32.ACT_SILU, alpha=0.5)
    ref2 = F.silu(0.5 * ref + bias + res + rowvec.repeat_interleave(rpg, 0))
    assert relerr(out, ref2) < 5e-6
    # strided operand and destination (channel slices of wider buffers), in-place residual
    wide = frnd(M, N + 40, seed=6)
    before = wide.clone()
    ops32.gemm(a, w, residual=wide[:, 8:8 + N], out=wide[:, 8:8 + N])
    assert relerr(wide[:, 8:8 + N], ref + before[:, 8:8 + N]) < 5e-6
    assert torch.equal(wide[:, :8], before[:, :8]) and torch.equal(wide[:, 8 + N:], before[:, 8 + N:])


@pytest.mark.parametrize("B,H,W,Cin,Cout,mode", [(2, 16, 16, 4, 64, "same"), (1, 17, 9, 32, 48, "same"), (2, 16, 16, 64, 64, "s2"),
                                                  (1, 15, 15, 32, 32, "s2"), (2, 16, 16, 32, 32, "vae_s2"), (2, 8, 8, 64, 32, "up"),
                                                  (1, 32, 32, 128, 3, "nchw")])
def test_f32_conv3x3_modes(ops32, B, H, W, Cin, Cout, mode):
    x = frnd(B, H, W, Cin, seed=1)
    w = frnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2)
    bias = frnd(Cout, seed=3)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    xi = x.double().permute(0, 3, 1, 2)
    wd, bd = w.double(), bias.double()
    if mode == "same":
        emb = frnd(B, Cout, seed=4)
        out = ops32.conv3x3(x, wp, bias=bias, rowvec=emb)
        ref = (F.conv2d(xi, wd, bd, padding=1) + emb.double()[:, :, None, None]).permute(0, 2, 3, 1)
    elif mode == "s2":        # UNet Downsample: stride 2, pad 1 (model/unet.py:99-101)
        out = ops32.conv3x3(x, wp, bias=bias, stride=2, pad=(1, 1))
        ref = F.conv2d(xi, wd, bd, stride=2, padding=1).permute(0, 2, 3, 1)
    elif mode == "vae_s2":    # VAE Downsample: pad right / bottom only, stride 2 (model/vae.py:54-58)
        out = ops32.conv3x3(x, wp, bias=bias, stride=2, pad=(0, 0), out_hw=(H // 2, W // 2))
        ref = F.conv2d(F.pad(xi, (0, 1, 0, 1)), wd, bd, stride=2).permute(0, 2, 3, 1)
    elif mode == "up":
        out = ops32.conv3x3(x, wp, bias=bias, up2x=True)
        ref = F.conv2d(F.interpolate(xi, scale_factor=2.0, mode="nearest"), wd, bd, padding=1).permute(0, 2, 3, 1)
    else:
        out = ops32.conv3x3(x, wp, bias=bias, nchw=True).view(B, Cout, H, W)
        ref = F.conv2d(xi, wd, bd, padding=1)
    assert out.shape == ref.shape
    assert relerr(out, ref.float()) < 5e-6


@pytest.mark.parametrize("B,heads,Lq,Lk,d", [(2, 5, 256, 256, 64), (1, 3, 130, 77, 64), (2, 1, 64, 64, 512), (3, 2, 33, 1, 64)])
def test_f32_attention(ops32, B, heads, Lq, Lk, d):
    C = heads * d
    qkv = frnd(B, Lq, 3 * C, seed=1)
    q = qkv[..., :C]
    kv = frnd(B, Lk, 2 * C, seed=2)
    k, v = kv[..., :C], kv[..., C:]
    out = ops32.attention(q, k, v, heads, d ** -0.5)
    sp = lambda t: t.double().reshape(B, -1, heads, d).permute(0, 2, 1, 3)
    ref = (torch.softmax(sp(q) @ sp(k).transpose(-1, -2) * d ** -0.5, -1) @ sp(v)).permute(0, 2, 1, 3).reshape(B, Lq, C)
    assert relerr(out, ref.float()) < 5e-6


def test_f32_norms_and_elementwise(ops32):
    x = frnd(3, 24, 24, 320, seed=1) * 3 + 5            # a large mean: the double-precision statistics matter
    gamma, beta = frnd(320, seed=2), frnd(320, seed=3)
    for silu in (False, True):
        y = ops32.groupnorm(x, gamma, beta, 32, 1e-5, silu)
        ref = F.group_norm(x.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), 1e-5)
        ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
        assert relerr(y, ref.float()) < 5e-6
    wide = frnd(2, 64, 640, seed=4)
    ys = ops32.groupnorm(wide[..., 320:], gamma, beta, 32, 1e-6, False)      # a channel slice as input
    refs = F.group_norm(wide[..., 320:].double().permute(0, 2, 1), 32, gamma.double(), beta.double(), 1e-6).permute(0, 2, 1)
    assert relerr(ys, refs.float()) < 5e-6
    t = frnd(100, 320, seed=5)
    assert relerr(ops32.layernorm(t, gamma, beta, 1e-5), F.layer_norm(t.double(), (320,), gamma.double(), beta.double(), 1e-5).float()) < 5e-6
    g = frnd(50, 2560, seed=6)
    a, b = g.double().chunk(2, -1)
    assert relerr(ops32.geglu(g), (a * F.gelu(b)).float()) < 5e-6
    assert relerr(ops32.silu(t), F.silu(t.double()).float()) < 5e-6
    nchw = frnd(2, 4, 8, 8, seed=7)
    dst = torch.zeros(2, 8, 8, 8, device="cuda")
    ops32.nchw_to_nhwc(nchw, dst, 4, 0.5)
    assert torch.equal(dst[..., 4:], nchw.permute(0, 2, 3, 1) * 0.5) and dst[..., :4].abs().max().item() == 0
    ts = torch.tensor([200, 50, 999], device="cuda")
    half = 160
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float64, device="cuda") / half)
    arg = ts.double()[:, None] * f[None]
    assert relerr(ops32.timestep_embedding(ts, 320), torch.cat([arg.cos(), arg.sin()], -1).float()) < 2e-4
    with pytest.raises(RuntimeError):
        ops32.gemm(torch.zeros(4, 4), torch.zeros(4, 4))
