"""fp32 mode of the ControlLDM path: the same graph as ``engine.py`` evaluated with fp32 tensors and fp32 accumulation
through the ``edtr_f32_*`` kernels (``ops32.py``), for BASELINE.json's fp32 tolerance (per-step latent max-rel error
<= 1e-4 against the reference's fp32 path).

This is the accuracy mode, selected with ``ControlLDM.set_precision("fp32")``: eager launches, fp32 channels-last
activations, no weight repacking beyond the tap-major convolution layout, the contractions on the CUDA cores.  The
throughput mode is the bf16 tensor-core engine.  Covered: ``ControlLDM.forward`` (ControlNet + controlled UNet,
model/cldm.py:166-194), the sampler loop on top of it (the generic ``SpacedSampler`` loop with the fused update
kernel), ``vae_decode`` (model/cldm.py:136-156) and ``vae_encode`` (model/cldm.py:107-134); the tiled variants and
SwinIR stay bf16-only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import topology as T

F32 = torch.float32


def _pack(sd: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    """fp32 device copies; 3x3 filters [Cout, Cin, 3, 3] -> [Cout, 9 * Cin] tap-major / channel-minor (the K order of
    edtr_f32_gemm's convolution gather), 1x1 filters -> [Cout, Cin]."""
    w = {}
    for k, v in sd.items():
        t = v.detach().to(device=device, dtype=F32)
        if t.dim() == 4:
            co, ci, kh, kw = t.shape
            t = t.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
        w[k] = t.contiguous()
    return w


class _Runner32:
    """Leaf blocks of one UNet-family network on fp32 channels-last tensors; every block writes its result into `out`
    (possibly a channel slice of a concatenation buffer)."""

    def __init__(self, w: Dict[str, torch.Tensor], ops):
        self.w, self.ops = w, ops
        self.emb_silu: Optional[torch.Tensor] = None   # SiLU(emb) [B, 4 * model_channels]
        self.ctx: Optional[torch.Tensor] = None        # c_txt [B, 77, context_dim]

    def time_embedding(self, t: torch.Tensor, mc: int) -> None:
        """model/util.py:98-118, model/unet.py:475-480; the SiLU of every ResBlock's emb_layers (model/unet.py:166-172)
        is applied once."""
        ops, w = self.ops, self.w
        te = ops.timestep_embedding(t, mc)
        e = ops.gemm(te, w["time_embed.0.weight"], bias=w["time_embed.0.bias"], act=ops.ACT_SILU)
        e = ops.gemm(e, w["time_embed.2.weight"], bias=w["time_embed.2.bias"])
        self.emb_silu = ops.silu(e)

    def res(self, p: str, x: torch.Tensor, out: torch.Tensor) -> None:
        """ResBlock.forward (model/unet.py:203-223)."""
        ops, w = self.ops, self.w
        y = ops.groupnorm(x, w[p + "in_layers.0.weight"], w[p + "in_layers.0.bias"], 32, 1e-5, True)
        e = ops.gemm(self.emb_silu, w[p + "emb_layers.1.weight"], bias=w[p + "emb_layers.1.bias"])
        h = ops.conv3x3(y, w[p + "in_layers.2.weight"], bias=w[p + "in_layers.2.bias"], rowvec=e)
        y2 = ops.groupnorm(h, w[p + "out_layers.0.weight"], w[p + "out_layers.0.bias"], 32, 1e-5, True)
        if (p + "skip_connection.weight") in w:
            skip = ops.gemm(x, w[p + "skip_connection.weight"], bias=w[p + "skip_connection.bias"]).view(out.shape)
        else:
            skip = x
        ops.conv3x3(y2, w[p + "out_layers.3.weight"], bias=w[p + "out_layers.3.bias"], residual=skip, out=out)

    def st(self, p: str, x: torch.Tensor, out: torch.Tensor, heads: int) -> None:
        """SpatialTransformer.forward with one BasicTransformerBlock (model/attention.py:283-302, 230-234)."""
        ops, w = self.ops, self.w
        B, H, W, C = x.shape
        L = H * W
        t = p + "transformer_blocks.0."
        scale = float(C // heads) ** -0.5
        y = ops.groupnorm(x, w[p + "norm.weight"], w[p + "norm.bias"], 32, 1e-6, False).view(B, L, C)
        t0 = ops.gemm(y, w[p + "proj_in.weight"], bias=w[p + "proj_in.bias"]).view(B, L, C)
        n = ops.layernorm(t0, w[t + "norm1.weight"], w[t + "norm1.bias"], 1e-5)
        q = ops.gemm(n, w[t + "attn1.to_q.weight"]).view(B, L, C)
        k = ops.gemm(n, w[t + "attn1.to_k.weight"]).view(B, L, C)
        v = ops.gemm(n, w[t + "attn1.to_v.weight"]).view(B, L, C)
        a = ops.attention(q, k, v, heads, scale)
        t1 = ops.gemm(a, w[t + "attn1.to_out.0.weight"], bias=w[t + "attn1.to_out.0.bias"], residual=t0).view(B, L, C)
        n = ops.layernorm(t1, w[t + "norm2.weight"], w[t + "norm2.bias"], 1e-5)
        q = ops.gemm(n, w[t + "attn2.to_q.weight"]).view(B, L, C)
        Lc = self.ctx.shape[1]
        k = ops.gemm(self.ctx, w[t + "attn2.to_k.weight"]).view(B, Lc, C)
        v = ops.gemm(self.ctx, w[t + "attn2.to_v.weight"]).view(B, Lc, C)
        a = ops.attention(q, k, v, heads, scale)
        t2 = ops.gemm(a, w[t + "attn2.to_out.0.weight"], bias=w[t + "attn2.to_out.0.bias"], residual=t1).view(B, L, C)
        n = ops.layernorm(t2, w[t + "norm3.weight"], w[t + "norm3.bias"], 1e-5)
        g = ops.gemm(n, w[t + "ff.net.0.proj.weight"], bias=w[t + "ff.net.0.proj.bias"])
        gg = ops.geglu(g)
        t3 = ops.gemm(gg, w[t + "ff.net.2.weight"], bias=w[t + "ff.net.2.bias"], residual=t2).view(B, L, C)
        ops.gemm(t3, w[p + "proj_out.weight"], bias=w[p + "proj_out.bias"], residual=x, out=out)

    def block(self, prefix: str, layers, x: torch.Tensor, out: torch.Tensor) -> None:
        """TimestepEmbedSequential.forward (model/unet.py:40-48); the last layer writes `out`."""
        ops, w = self.ops, self.w
        n = len(layers)
        for i, layer in enumerate(layers):
            p = f"{prefix}{i}."
            kind = layer[0]
            B, H, W, _ = x.shape
            if kind == "up":
                H, W = 2 * H, 2 * W
            elif kind == "down":
                H, W = (H + 1) // 2, (W + 1) // 2
            cout = layer[2] if kind in ("conv_in", "res") else layer[1]
            dst = out if i == n - 1 else torch.empty((B, H, W, cout), dtype=F32, device=x.device)
            if kind == "conv_in":
                ops.conv3x3(x, w[p + "weight"], bias=w[p + "bias"], out=dst)
            elif kind == "res":
                self.res(p, x, dst)
            elif kind == "st":
                self.st(p, x, dst, layer[2])
            elif kind == "down":       # model/unet.py:99-108: 3x3, stride 2, pad 1
                ops.conv3x3(x, w[p + "op.weight"], bias=w[p + "op.bias"], stride=2, pad=(1, 1), out_hw=(H, W), out=dst)
            elif kind == "up":         # model/unet.py:69-79: nearest x2, then 3x3
                ops.conv3x3(x, w[p + "conv.weight"], bias=w[p + "conv.bias"], up2x=True, out=dst)
            x = dst


class CldmEngineF32:
    """ControlLDM.forward in fp32 (model/cldm.py:166-194, model/controlnet.py:18-38, 263-277)."""

    def __init__(self, unet_cfg: Dict, controlnet_cfg: Dict, unet_sd, controlnet_sd, device, ops=None):
        if ops is None:
            from . import ops32 as ops
        self.ops = ops
        self.device = torch.device(device)
        self.ucfg, self.ccfg = unet_cfg, controlnet_cfg
        self.u_in, self.u_mid, self.u_out = T.unet_blocks(unet_cfg, False)
        self.c_in, self.c_mid, _ = T.unet_blocks(controlnet_cfg, True)
        for k, shp in list(T.unet_param_shapes(unet_cfg, False)) + []:
            if k not in unet_sd or tuple(unet_sd[k].shape) != tuple(shp):
                raise ValueError(f"UNet state-dict: {k} missing or of the wrong shape")
        self.uw = _pack(unet_sd, self.device)
        self.cw = _pack(controlnet_sd, self.device)
        self.zc = unet_cfg["in_channels"]
        self.out_c = unet_cfg["out_channels"]
        # per input block: down-sampling factor and width (the skip tensors hs[s])
        self.in_ds, self.in_ch = [], []
        ds = 1
        for blk in self.u_in:
            if blk[0][0] == "down":
                ds *= 2
            self.in_ds.append(ds)
            self.in_ch.append(T.block_out_channels(blk))
        self.mid_ch = self.in_ch[-1]

    @torch.no_grad()
    def forward(self, x_noisy: torch.Tensor, t: torch.Tensor, c_img: torch.Tensor, c_txt: torch.Tensor,
                control_scales: Optional[Sequence[float]] = None) -> torch.Tensor:
        ops = self.ops
        if getattr(ops, "REQUIRES_CUDA", True) and not x_noisy.is_cuda:
            raise RuntimeError("edtr_b200 has no CPU path: inputs must be CUDA tensors")
        B, zc, H, W = x_noisy.shape
        if zc != self.zc or c_img.shape != x_noisy.shape:
            raise ValueError(f"x_noisy / c_img must be [B, {self.zc}, H, W]")
        n_in = len(self.u_in)
        scales = list(control_scales) if control_scales is not None else [1.0] * (n_in + 1)
        if len(scales) < n_in + 1:     # (ControlLDM keeps the reference's 13 entries whatever the depth, model/cldm.py:34)
            raise ValueError(f"control_scales must have at least {n_in + 1} entries")
        dev = x_noisy.device
        new = lambda *shape: torch.empty(shape, dtype=F32, device=dev)
        with ops.device_guard(dev):
            un, cn = _Runner32(self.uw, ops), _Runner32(self.cw, ops)
            ctx = c_txt.to(F32).contiguous()
            un.ctx = cn.ctx = ctx
            tt = t.long().contiguous()
            un.time_embedding(tt, self.ucfg["model_channels"])
            cn.time_embedding(tt, self.ccfg["model_channels"])
            xu = new(B, H, W, zc)
            ops.nchw_to_nhwc(x_noisy.to(F32).contiguous(), xu, 0)
            xc = new(B, H, W, 2 * zc)                      # cat(x, hint): model/controlnet.py:266
            ops.nchw_to_nhwc(x_noisy.to(F32).contiguous(), xc, 0)
            ops.nchw_to_nhwc(c_img.to(F32).contiguous(), xc, zc)

            # ControlNet (model/controlnet.py:263-277)
            couts: List[torch.Tensor] = []
            h = xc
            for s, blk in enumerate(self.c_in):
                d = self.in_ds[s]
                dst = new(B, H // d, W // d, self.in_ch[s])
                cn.block(f"input_blocks.{s}.", blk, h, dst)
                h = dst
                couts.append(dst)
            d = self.in_ds[-1]
            cmid = new(B, H // d, W // d, self.mid_ch)
            cn.block("middle_block.", self.c_mid, h, cmid)
            couts.append(cmid)

            # decoder input buffers [previous output | skip]; the encoder writes the skips straight into them
            cats = []
            cp = self.mid_ch
            for j, blk in enumerate(self.u_out):
                s = n_in - 1 - j
                d = self.in_ds[s]
                cats.append((new(B, H // d, W // d, cp + self.in_ch[s]), cp))
                cp = T.block_out_channels(blk)

            def hs_view(s: int) -> torch.Tensor:
                buf, c0 = cats[n_in - 1 - s]
                return buf[..., c0:]

            # UNet encoder + middle (model/controlnet.py:25-28)
            h = xu
            for s, blk in enumerate(self.u_in):
                un.block(f"input_blocks.{s}.", blk, h, hs_view(s))
                h = hs_view(s)
            mid = cats[0][0][..., :cats[0][1]]
            un.block("middle_block.", self.u_mid, h, mid)
            # zero-convs accumulate into the UNet tensors (model/controlnet.py:270-275, 31, 37; scales: model/cldm.py:189)
            cwt = self.cw
            for s in range(n_in):
                q = f"zero_convs.{s}.0."
                ops.gemm(couts[s], cwt[q + "weight"], bias=cwt[q + "bias"] * scales[s] if scales[s] != 1.0 else cwt[q + "bias"],
                         residual=hs_view(s), out=hs_view(s), alpha=float(scales[s]))
            q = "middle_block_out.0."
            ops.gemm(couts[n_in], cwt[q + "weight"], bias=cwt[q + "bias"] * scales[n_in] if scales[n_in] != 1.0 else cwt[q + "bias"],
                     residual=mid, out=mid, alpha=float(scales[n_in]))
            # UNet decoder (model/controlnet.py:33-38)
            n_out = len(self.u_out)
            for j, blk in enumerate(self.u_out):
                if j + 1 < n_out:
                    out = cats[j + 1][0][..., :cats[j + 1][1]]
                else:
                    out = new(B, H, W, T.block_out_channels(blk))
                un.block(f"output_blocks.{j}.", blk, cats[j][0], out)
            uw = self.uw
            y = ops.groupnorm(out, uw["out.0.weight"], uw["out.0.bias"], 32, 1e-5, True)
            eps = new(B, self.out_c, H, W)
            ops.conv3x3(y, uw["out.2.weight"], bias=uw["out.2.bias"], out=eps.view(B, self.out_c, H * W), nchw=True)
        return eps


class VaeDecoderF32:
    """ControlLDM.vae_decode in fp32: z / scale -> post_quant_conv -> Decoder.forward
    (model/cldm.py:136-156, model/vae.py:731-734, 527-560, 103-124, 279-308)."""

    def __init__(self, ddconfig: Dict, embed_dim: int, sd: Dict[str, torch.Tensor], device, ops=None):
        if ops is None:
            from . import ops32 as ops
        self.ops = ops
        self.device = torch.device(device)
        self.dd = ddconfig
        self.levels, _ = T.vae_decoder_levels(ddconfig)
        keep = {k: v for k, v in sd.items() if k.startswith("decoder.") or k.startswith("post_quant_conv.")}
        self.w = _pack(keep, self.device)

    def _res(self, p: str, x: torch.Tensor) -> torch.Tensor:
        ops, w = self.ops, self.w
        y = ops.groupnorm(x, w[p + "norm1.weight"], w[p + "norm1.bias"], 32, 1e-6, True)
        h = ops.conv3x3(y, w[p + "conv1.weight"], bias=w[p + "conv1.bias"])
        y2 = ops.groupnorm(h, w[p + "norm2.weight"], w[p + "norm2.bias"], 32, 1e-6, True)
        skip = x
        if (p + "nin_shortcut.weight") in w:
            skip = ops.gemm(x, w[p + "nin_shortcut.weight"], bias=w[p + "nin_shortcut.bias"]).view(h.shape)
        return ops.conv3x3(y2, w[p + "conv2.weight"], bias=w[p + "conv2.bias"], residual=skip)

    def _attn(self, p: str, x: torch.Tensor) -> torch.Tensor:
        ops, w = self.ops, self.w
        B, H, W, C = x.shape
        L = H * W
        y = ops.groupnorm(x, w[p + "norm.weight"], w[p + "norm.bias"], 32, 1e-6, False).view(B, L, C)
        q = ops.gemm(y, w[p + "q.weight"], bias=w[p + "q.bias"]).view(B, L, C)
        k = ops.gemm(y, w[p + "k.weight"], bias=w[p + "k.bias"]).view(B, L, C)
        v = ops.gemm(y, w[p + "v.weight"], bias=w[p + "v.bias"]).view(B, L, C)
        a = ops.attention(q, k, v, 1, float(C) ** -0.5)
        return ops.gemm(a, w[p + "proj_out.weight"], bias=w[p + "proj_out.bias"], residual=x).view(B, H, W, C)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, scale_factor: float) -> torch.Tensor:
        ops, w = self.ops, self.w
        if getattr(ops, "REQUIRES_CUDA", True) and not z.is_cuda:
            raise RuntimeError("edtr_b200 has no CPU path: z must be a CUDA tensor")
        B, zc, H, W = z.shape
        dev = z.device
        with ops.device_guard(dev):
            zin = torch.empty((B, H, W, zc), dtype=F32, device=dev)
            ops.nchw_to_nhwc(z.to(F32).contiguous(), zin, 0, 1.0 / scale_factor)
            h = ops.gemm(zin, w["post_quant_conv.weight"], bias=w["post_quant_conv.bias"]).view(B, H, W, -1)
            h = ops.conv3x3(h, w["decoder.conv_in.weight"], bias=w["decoder.conv_in.bias"])
            h = self._res("decoder.mid.block_1.", h)
            h = self._attn("decoder.mid.attn_1.", h)
            h = self._res("decoder.mid.block_2.", h)
            for level, blocks, has_up in self.levels:
                for i in range(len(blocks)):
                    h = self._res(f"decoder.up.{level}.block.{i}.", h)
                if has_up:   # Upsample: nearest x2 then conv (model/vae.py:36-38)
                    q = f"decoder.up.{level}.upsample.conv."
                    h = ops.conv3x3(h, w[q + "weight"], bias=w[q + "bias"], up2x=True)
            y = ops.groupnorm(h, w["decoder.norm_out.weight"], w["decoder.norm_out.bias"], 32, 1e-6, True)
            Bh, Hh, Wh, _ = y.shape
            out_ch = self.dd["out_ch"]
            img = torch.empty((B, out_ch, Hh, Wh), dtype=F32, device=dev)
            ops.conv3x3(y, w["decoder.conv_out.weight"], bias=w["decoder.conv_out.bias"], out=img.view(B, out_ch, Hh * Wh),
                        nchw=True)
        return img


class VaeEncoderF32(VaeDecoderF32):
    """ControlLDM.vae_encode in fp32: Encoder.forward + quant_conv -> posterior moments
    (model/cldm.py:107-134, model/vae.py:326-446, 725-729; Downsample pads right / bottom only, model/vae.py:54-58)."""

    def __init__(self, ddconfig: Dict, embed_dim: int, sd: Dict[str, torch.Tensor], device, ops=None):
        if ops is None:
            from . import ops32 as ops
        self.ops = ops
        self.device = torch.device(device)
        self.dd = ddconfig
        self.embed_dim = embed_dim
        self.levels, self.top = T.vae_encoder_levels(ddconfig)
        keep = {k: v for k, v in sd.items() if k.startswith("encoder.") or k.startswith("quant_conv.")}
        self.w = _pack(keep, self.device)

    @torch.no_grad()
    def encode(self, image: torch.Tensor) -> torch.Tensor:
        """image [B, in_channels, H, W] fp32 in [-1, 1] -> moments [B, 2 * embed_dim, H/f, W/f] fp32 (mean | logvar)."""
        ops, w = self.ops, self.w
        if getattr(ops, "REQUIRES_CUDA", True) and not image.is_cuda:
            raise RuntimeError("edtr_b200 has no CPU path: image must be a CUDA tensor")
        B, cin, H, W = image.shape
        f = 2 ** (len(self.levels) - 1)
        if cin != self.dd["in_channels"] or H % f or W % f:
            raise ValueError(f"image must be [B, {self.dd['in_channels']}, H, W] with H, W multiples of {f}")
        dev = image.device
        with ops.device_guard(dev):
            x = torch.empty((B, H, W, cin), dtype=F32, device=dev)
            ops.nchw_to_nhwc(image.to(F32).contiguous(), x, 0)
            h = ops.conv3x3(x, w["encoder.conv_in.weight"], bias=w["encoder.conv_in.bias"])
            for level, blocks, has_down in self.levels:
                for i in range(len(blocks)):
                    h = self._res(f"encoder.down.{level}.block.{i}.", h)
                if has_down:
                    q = f"encoder.down.{level}.downsample.conv."
                    Hc, Wc = h.shape[1], h.shape[2]
                    h = ops.conv3x3(h, w[q + "weight"], bias=w[q + "bias"], stride=2, pad=(0, 0), out_hw=(Hc // 2, Wc // 2))
            h = self._res("encoder.mid.block_1.", h)
            h = self._attn("encoder.mid.attn_1.", h)
            h = self._res("encoder.mid.block_2.", h)
            y = ops.groupnorm(h, w["encoder.norm_out.weight"], w["encoder.norm_out.bias"], 32, 1e-6, True)
            m = ops.conv3x3(y, w["encoder.conv_out.weight"], bias=w["encoder.conv_out.bias"])
            Bh, Hh, Wh, _ = m.shape
            mo = torch.empty((B, 2 * self.embed_dim, Hh, Wh), dtype=F32, device=dev)
            ops.gemm(m, w["quant_conv.weight"], bias=w["quant_conv.bias"], out=mo.view(B, 2 * self.embed_dim, Hh * Wh),
                     nchw_hw=Hh * Wh)
        return mo
