"""Generates the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py [--full]

The reference (JaehaKim97/EDTR) is imported from its tree with three import stubs for
packages missing here (omegaconf / ftfy / timm — none is on the hot path).
The synthetic weights of ``oracle.cldm_oracle`` are loaded into the reference modules with
``load_state_dict`` (strict for UNet / ControlNet, which also pins the oracle's key/shape
enumeration), then the reference's own public API is executed on CPU in fp32:
``SpacedSampler.manual_sample_with_timesteps`` -> ``ControlLDM.forward`` and
``ControlLDM.vae_decode``.  ``torch.randn_like`` is patched during the sampler call so the
four per-step noise draws are the seeded tensors the oracle receives explicitly.

Outputs:  golden_tiny.npz  (TINY config, B=2, 16x16 latent; everything in fp32)
          golden_wavelet.npz    (utils.common.wavelet_reconstruction on two random [2,3,40,56] images)
          golden_vae_encode.npz (TINY VAE encoder, B=2, 64x64 image: vae_encode mode / sample, q_sample at t=200)
          golden_vae_encode_tiled.npz (TINY_VAE8 encoder, B=2, 136x160 image (stored fp16), vae_encode(tiled=True, 64))
          golden_vae_tiled.npz  (TINY_VAE8 decoder, B=2, 40x48 latent, vae_decode(tiled=True, tile_size=16))
          golden_s4.npz    (s4 config, B=1, 64x64 latent; --full; image stored as fp16)
          golden_swinir.npz (SWINIR_TINY: model.swinir.SwinIR on [2,3,128,128] and [1,3,64,192] images; --swinir alone)
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("EDTR_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def _stub_missing_packages():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    try:
        import omegaconf  # noqa: F401
    except ImportError:
        mod("omegaconf")
        mod("omegaconf.listconfig", ListConfig=type("ListConfig", (list,), {}))
    try:
        import ftfy  # noqa: F401
    except ImportError:
        mod("ftfy", fix_text=lambda s: s)
    try:
        import timm  # noqa: F401
    except ImportError:
        import torch.nn as nn

        mod("timm")
        mod("timm.models")
        mod("timm.models.layers", DropPath=nn.Identity, to_2tuple=lambda x: (x, x),
            trunc_normal_=lambda t, std=1.0, **kw: torch.nn.init.trunc_normal_(t, std=std))


def build_reference(cfg, weights):
    """ControlLDM shell without CLIP (its ctor would build a 1 GB text tower unrelated to the path)."""
    _stub_missing_packages()
    from model.cldm import ControlLDM
    from model.controlnet import ControlledUnetModel, ControlNet
    from model.vae import AutoencoderKL

    def common(c):
        return dict(image_size=32, in_channels=c["in_channels"], model_channels=c["model_channels"],
                    attention_resolutions=list(c["attention_resolutions"]), num_res_blocks=c["num_res_blocks"],
                    channel_mult=list(c["channel_mult"]), num_head_channels=c["num_head_channels"],
                    use_spatial_transformer=True, use_linear_in_transformer=True, transformer_depth=1,
                    context_dim=c["context_dim"], legacy=False, use_checkpoint=True)

    m = ControlLDM.__new__(ControlLDM)
    torch.nn.Module.__init__(m)
    m.unet = ControlledUnetModel(out_channels=cfg["unet"]["out_channels"], **common(cfg["unet"]))
    m.controlnet = ControlNet(hint_channels=cfg["controlnet"]["hint_channels"], **common(cfg["controlnet"]))
    v = cfg["vae"]
    m.vae = AutoencoderKL(ddconfig=dict(double_z=True, z_channels=v["z_channels"], resolution=256,
                                        in_channels=v["in_channels"], out_ch=v["out_ch"], ch=v["ch"],
                                        ch_mult=list(v["ch_mult"]), num_res_blocks=v["num_res_blocks"],
                                        attn_resolutions=[], dropout=0.0), embed_dim=v["embed_dim"])
    m.scale_factor = cfg["latent_scale_factor"]
    m.control_scales = [1.0] * (len(m.controlnet.zero_convs) + 1)
    m.unet.load_state_dict(weights["unet"], strict=True)
    m.controlnet.load_state_dict(weights["controlnet"], strict=True)
    res = m.vae.load_state_dict(weights["vae"], strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith(("encoder.", "quant_conv.")) for k in res.missing_keys), res.missing_keys
    return m.eval()


def run_reference(cfg, batch, latent_hw):
    from oracle import cldm_oracle as O

    weights = O.make_cldm_weights(cfg, seed=0)
    x_T, cond, noise = O.make_inputs(cfg, batch, latent_hw, seed=1)
    model = build_reference(cfg, weights)
    from utils.sampler import SpacedSampler

    sampler = SpacedSampler(O.make_betas(**cfg["diffusion"]))
    used = list(cfg["used_timesteps"])
    draws = list(noise)
    real_randn_like = torch.randn_like
    torch.randn_like = lambda t, *a, **k: draws.pop(0)
    try:
        with torch.no_grad():
            # per-step x_prev through the public p_sample, exactly as the loop of
            # manual_sample_with_timesteps does (utils/sampler.py:304-314)
            sampler.make_schedule(len(used), used)
            ts = np.flip(sampler.timesteps)
            x = x_T
            xs, x0s, eps0 = [], [], None
            for i, step in enumerate(ts):
                t = torch.full((batch,), int(step), dtype=torch.long)
                index = torch.full_like(t, len(ts) - i - 1)
                if i == 0:
                    eps0 = model(x, t, cond)
                x, x0 = sampler.p_sample(model, x, t, index, cond, None, 1.0)
                xs.append(x)
                x0s.append(x0)
            # and the whole public loop once more, to pin the loop itself
            draws[:] = list(noise)
            z = sampler.manual_sample_with_timesteps(model, "cpu", x_T, len(used), used, batch, cond, None, 1.0,
                                                     progress=False)
            assert torch.equal(z, xs[-1])
            img = model.vae_decode(z)
    finally:
        torch.randn_like = real_randn_like
    sched = {k: getattr(sampler, k).numpy() for k in
             ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
              "posterior_mean_coef1", "posterior_mean_coef2")}
    return dict(eps0=eps0, xs=torch.stack(xs), x0s=torch.stack(x0s), img=img, sched=sched)


def run_reference_tiled_vae(vae_cfg, batch, latent_h, latent_w, tile_size):
    """ControlLDM.vae_decode(z, tiled=True, tile_size) of the unmodified reference (VAEHook, non-fast mode)."""
    from oracle import cldm_oracle as O

    _stub_missing_packages()
    from model.cldm import ControlLDM
    from model.vae import AutoencoderKL

    sd = O.make_weights(O.vae_decoder_param_shapes(vae_cfg), seed=2)
    m = ControlLDM.__new__(ControlLDM)
    torch.nn.Module.__init__(m)
    m.vae = AutoencoderKL(ddconfig=dict(double_z=True, z_channels=vae_cfg["z_channels"], resolution=256,
                                        in_channels=vae_cfg["in_channels"], out_ch=vae_cfg["out_ch"], ch=vae_cfg["ch"],
                                        ch_mult=list(vae_cfg["ch_mult"]), num_res_blocks=vae_cfg["num_res_blocks"],
                                        attn_resolutions=[], dropout=0.0), embed_dim=vae_cfg["embed_dim"])
    res = m.vae.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    m.scale_factor = 0.18215
    m.eval()
    g = torch.Generator().manual_seed(11)
    z = 0.9 * torch.randn(batch, vae_cfg["embed_dim"], latent_h, latent_w, generator=g)
    with torch.no_grad():
        img = m.vae_decode(z, tiled=True, tile_size=tile_size)
        img_untiled = m.vae_decode(z)
    return z, img, img_untiled


def run_reference_vae_encode(vae_cfg, batch, hw):
    """ControlLDM.vae_encode(image, sample=False / True) and Diffusion.q_sample of the unmodified reference."""
    from oracle import cldm_oracle as O

    _stub_missing_packages()
    from model.cldm import ControlLDM
    from model.gaussian_diffusion import Diffusion
    from model.vae import AutoencoderKL

    sd = O.make_weights(O.vae_encoder_param_shapes(vae_cfg), seed=3)
    m = ControlLDM.__new__(ControlLDM)
    torch.nn.Module.__init__(m)
    m.vae = AutoencoderKL(ddconfig=dict(double_z=True, z_channels=vae_cfg["z_channels"], resolution=256,
                                        in_channels=vae_cfg["in_channels"], out_ch=vae_cfg["out_ch"], ch=vae_cfg["ch"],
                                        ch_mult=list(vae_cfg["ch_mult"]), num_res_blocks=vae_cfg["num_res_blocks"],
                                        attn_resolutions=[], dropout=0.0), embed_dim=vae_cfg["embed_dim"])
    res = m.vae.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith(("decoder.", "post_quant_conv.")) for k in res.missing_keys), res.missing_keys
    m.scale_factor = 0.18215
    m.eval()
    g = torch.Generator().manual_seed(21)
    image = torch.rand(batch, 3, hw, hw, generator=g) * 2 - 1
    with torch.no_grad():
        z_mode = m.vae_encode(image, sample=False)
        draw = torch.randn(z_mode.shape, generator=g)
        real = torch.randn
        torch.randn = lambda *a, **k: draw          # DiagonalGaussianDistribution.sample draws torch.randn(mean.shape)
        try:
            z_sample = m.vae_encode(image, sample=True)
        finally:
            torch.randn = real
        diffusion = Diffusion(timesteps=1000, beta_schedule="linear", loss_type="l2", linear_start=0.00085,
                              linear_end=0.0120, cosine_s=8e-3, parameterization="eps")
        t = torch.full((batch,), 200, dtype=torch.long)
        n2 = torch.randn(z_mode.shape, generator=g)
        x_T = diffusion.q_sample(x_start=z_mode, t=t, noise=n2)
    return image, z_mode, draw, z_sample, n2, x_T


def run_reference_tiled_vae_encode(vae_cfg, batch, h, w, tile_size):
    """ControlLDM.vae_encode(image, sample=False, tiled=True, tile_size) of the unmodified reference."""
    from oracle import cldm_oracle as O

    _stub_missing_packages()
    from model.cldm import ControlLDM
    from model.vae import AutoencoderKL

    sd = O.make_weights(O.vae_encoder_param_shapes(vae_cfg), seed=3)
    m = ControlLDM.__new__(ControlLDM)
    torch.nn.Module.__init__(m)
    m.vae = AutoencoderKL(ddconfig=dict(double_z=True, z_channels=vae_cfg["z_channels"], resolution=256,
                                        in_channels=vae_cfg["in_channels"], out_ch=vae_cfg["out_ch"], ch=vae_cfg["ch"],
                                        ch_mult=list(vae_cfg["ch_mult"]), num_res_blocks=vae_cfg["num_res_blocks"],
                                        attn_resolutions=[], dropout=0.0), embed_dim=vae_cfg["embed_dim"])
    res = m.vae.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    m.scale_factor = 0.18215
    m.eval()
    g = torch.Generator().manual_seed(41)
    image = (torch.rand(batch, 3, h, w, generator=g) * 2 - 1).half().float()   # fp16-exact: stored as fp16
    with torch.no_grad():
        z = m.vae_encode(image, sample=False, tiled=True, tile_size=tile_size)
    return image, z


def run_reference_swinir(cfg, shapes, seed=53):
    """model.swinir.SwinIR (the pre-restoration network) built exactly as
    configs/det/voc2012/test/007_edtr-s4.yaml:3-19 does (at the widths of `cfg`), with the oracle's synthetic
    parameters; the relative_position_index / attn_mask buffers stay the reference's own."""
    _stub_missing_packages()
    from model.swinir import SwinIR
    from oracle import swinir_oracle as S

    m = SwinIR(img_size=cfg["img_size"], patch_size=1, in_chans=cfg["in_chans"], embed_dim=cfg["embed_dim"],
               depths=list(cfg["depths"]), num_heads=list(cfg["num_heads"]), window_size=cfg["window_size"],
               mlp_ratio=cfg["mlp_ratio"], sf=cfg["sf"], img_range=cfg["img_range"], upsampler="nearest+conv",
               resi_connection="1conv", unshuffle=True, unshuffle_scale=cfg["sf"])
    assert cfg["num_feat"] == 64   # hard-coded in the reference (model/swinir.py:686)
    sd = S.make_swinir_weights(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith(("relative_position_index", "attn_mask")) for k in missing), missing
    params = {k for k, _ in m.named_parameters()}
    assert params == set(sd), (sorted(params - set(sd))[:5], sorted(set(sd) - params)[:5])
    m.eval()
    g = torch.Generator().manual_seed(seed)
    out = {}
    for i, (b, h, w) in enumerate(shapes):
        x = torch.rand(b, 3, h, w, generator=g)
        with torch.no_grad():
            y = m(x)
        out[f"x{i}"], out[f"y{i}"] = x.numpy(), y.numpy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also generate the s4 (full-size) fixture; ~2 min")
    ap.add_argument("--swinir", action="store_true", help="only (re)generate golden_swinir.npz")
    args = ap.parse_args()
    from oracle import cldm_oracle as O
    from oracle import swinir_oracle as S

    torch.set_num_threads(os.cpu_count())
    sw = run_reference_swinir(S.SWINIR_TINY, [(2, 128, 128), (1, 64, 192)])
    np.savez_compressed(os.path.join(HERE, "golden_swinir.npz"), **sw)
    print("swinir:", {k: tuple(v.shape) for k, v in sw.items()})
    if args.swinir:
        return

    torch.set_num_threads(os.cpu_count())
    r = run_reference(O.TINY, batch=2, latent_hw=16)
    np.savez_compressed(os.path.join(HERE, "golden_tiny.npz"), eps0=r["eps0"].numpy(), xs=r["xs"].numpy(),
                        x0s=r["x0s"].numpy(), img=r["img"].numpy(), **{"sched_" + k: v for k, v in r["sched"].items()})
    print("tiny:", {k: tuple(v.shape) for k, v in r.items() if hasattr(v, "shape")})
    z, img, img_u = run_reference_tiled_vae(O.TINY_VAE8, batch=2, latent_h=40, latent_w=48, tile_size=16)
    np.savez_compressed(os.path.join(HERE, "golden_vae_tiled.npz"), z=z.numpy(), img=img.numpy().astype(np.float32),
                        tile_size=np.int64(16))
    _stub_missing_packages()
    from utils.common import wavelet_reconstruction as wr_ref
    gg = torch.Generator().manual_seed(31)
    content, style = torch.rand(2, 3, 40, 56, generator=gg), torch.rand(2, 3, 40, 56, generator=gg)
    np.savez_compressed(os.path.join(HERE, "golden_wavelet.npz"), content=content.numpy(), style=style.numpy(),
                        out=wr_ref(content, style).numpy())
    image, z_mode, draw, z_sample, n2, x_T = run_reference_vae_encode(O.TINY["vae"], batch=2, hw=64)
    np.savez_compressed(os.path.join(HERE, "golden_vae_encode.npz"), image=image.numpy(), z_mode=z_mode.numpy(),
                        draw=draw.numpy(), z_sample=z_sample.numpy(), q_noise=n2.numpy(), x_T=x_T.numpy())
    print("vae encode:", tuple(image.shape), "->", tuple(z_mode.shape))
    timg, tz = run_reference_tiled_vae_encode(O.TINY_VAE8, batch=2, h=136, w=160, tile_size=64)
    np.savez_compressed(os.path.join(HERE, "golden_vae_encode_tiled.npz"), image=timg.numpy().astype(np.float16),
                        z=tz.numpy(), tile_size=np.int64(64))
    print("tiled vae encode:", tuple(timg.shape), "->", tuple(tz.shape))
    print("tiled vae:", tuple(z.shape), tuple(img.shape), "tiled vs untiled max diff",
          float((img - img_u).abs().max()))
    if args.full:
        r = run_reference(O.S4, batch=1, latent_hw=64)
        np.savez_compressed(os.path.join(HERE, "golden_s4.npz"), eps0=r["eps0"].numpy(), xs=r["xs"].numpy(),
                            x0s=r["x0s"].numpy(), img=r["img"].numpy().astype(np.float16),
                            **{"sched_" + k: v for k, v in r["sched"].items()})
        print("s4:", {k: tuple(v.shape) for k, v in r.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
