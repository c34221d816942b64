"""CPU tests: the oracle restatement against fixtures recorded from the live reference
(tests/golden/make_golden.py), plus the host-side schedule arithmetic."""
import os

import numpy as np
import pytest
import torch

from oracle import cldm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tiny():
    w = O.make_cldm_weights(O.TINY, seed=0)
    x_T, cond, noise = O.make_inputs(O.TINY, 2, 16, seed=1)
    return w, x_T, cond, noise, np.load(os.path.join(GOLD, "golden_tiny.npz"))


def test_oracle_matches_reference_tiny(tiny):
    w, x_T, cond, noise, g = tiny
    with torch.no_grad():
        t = torch.full((2,), 200, dtype=torch.long)
        eps0 = O.cldm_forward(w, O.TINY, x_T, t, cond)
        z, xs, x0s = O.sample(w, O.TINY, x_T, cond, noise)
        img = O.vae_decode(w["vae"], O.TINY["vae"], z, O.TINY["latent_scale_factor"])
    assert np.abs(g["eps0"]).max() > 0.05  # the fixture is not the all-zero trap of zero_module
    assert O.max_rel_err(eps0, torch.from_numpy(g["eps0"])) < 1e-5
    for i in range(4):
        assert O.max_rel_err(xs[i], torch.from_numpy(g["xs"][i])) < 1e-5
        assert O.max_rel_err(x0s[i], torch.from_numpy(g["x0s"][i])) < 1e-5
    assert O.max_rel_err(img, torch.from_numpy(g["img"])) < 1e-5


def test_schedule_matches_reference(tiny):
    g = tiny[-1]
    s = O.make_schedule(O.make_betas(**O.TINY["diffusion"]), 4, O.TINY["used_timesteps"])
    for k in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
              "posterior_mean_coef1", "posterior_mean_coef2"):
        assert np.array_equal(s[k], g["sched_" + k]), k
    assert list(s["timesteps"]) == [50, 100, 150, 200]
    # SURVEY.md §3.2 measured values
    assert np.allclose(s["posterior_mean_coef1"], [1.0, 0.5558, 0.4079, 0.3307], atol=1e-4)
    assert np.allclose(s["posterior_variance"], [0, 0.02759, 0.04562, 0.06259], atol=1e-5)


def test_space_timesteps():
    assert O.space_timesteps(1000, "4") == {0, 333, 666, 999}
    assert O.space_timesteps(300, [10, 15, 20]) is not None
    assert len(O.space_timesteps(1000, "ddim50")) == 50
    with pytest.raises(ValueError):
        O.space_timesteps(10, "20")


def test_param_enumeration_counts():
    n = lambda shapes: sum(int(np.prod(s)) for _, s in shapes)
    # SURVEY.md / BASELINE.md §2: UNet 865.9 M, ControlNet 363.2 M, VAE decoder 49.5 M parameters
    assert abs(n(O.unet_param_shapes(O.S4["unet"])) / 1e6 - 865.9) < 0.1
    assert abs(n(O.unet_param_shapes(O.S4["controlnet"], True)) / 1e6 - 363.2) < 0.5
    assert abs(n(O.vae_decoder_param_shapes(O.S4["vae"])) / 1e6 - 49.5) < 0.1


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "golden_s4.npz")), reason="full-size fixture absent")
def test_s4_fixture_is_sane():
    g = np.load(os.path.join(GOLD, "golden_s4.npz"))
    assert g["xs"].shape == (4, 1, 4, 64, 64) and g["img"].shape == (1, 3, 512, 512)
    assert np.isfinite(g["xs"]).all() and np.abs(g["eps0"]).max() > 0.05


def test_oracle_tiled_vae_matches_reference():
    """vae_decode(tiled=True) of the live reference (VAEHook, pooled GroupNorm statistics) vs the restatement."""
    g = np.load(os.path.join(GOLD, "golden_vae_tiled.npz"))
    sd = O.make_weights(O.vae_decoder_param_shapes(O.TINY_VAE8), seed=2)
    z = torch.from_numpy(g["z"])
    with torch.no_grad():
        img = O.vae_decode_tiled(sd, O.TINY_VAE8, z, 0.18215, int(g["tile_size"]))
        img_untiled = O.vae_decode(sd, O.TINY_VAE8, z, 0.18215)
    assert O.max_rel_err(img, torch.from_numpy(g["img"])) < 1e-5
    # the pooled statistics are an approximation: tiled != untiled, so the test is not vacuous
    assert O.max_rel_err(img, img_untiled) > 1e-3


def test_vae_split_tiles_matches_survey():
    """a 256x256 latent with decoder tile 64 -> 16 tiles of <= 86x86 latent pixels."""
    ib, ob = O.vae_split_tiles(256, 256, 64)
    assert len(ib) == 16 and ib[5] == [64, 150, 64, 150] and ob[5] == [600, 1112, 600, 1112]
    assert ob[-1] == [1624, 2048, 1624, 2048]


def test_oracle_vae_encode_and_q_sample_match_reference():
    """ControlLDM.vae_encode (mode and sample) and Diffusion.q_sample of the live reference vs the restatement."""
    g = np.load(os.path.join(GOLD, "golden_vae_encode.npz"))
    sd = O.make_weights(O.vae_encoder_param_shapes(O.TINY["vae"]), seed=3)
    image = torch.from_numpy(g["image"])
    with torch.no_grad():
        z_mode = O.vae_encode(sd, O.TINY["vae"], image, 0.18215)
        z_sample = O.vae_encode(sd, O.TINY["vae"], image, 0.18215, noise=torch.from_numpy(g["draw"]))
    assert O.max_rel_err(z_mode, torch.from_numpy(g["z_mode"])) < 1e-5
    assert O.max_rel_err(z_sample, torch.from_numpy(g["z_sample"])) < 1e-5
    t = torch.full((2,), 200, dtype=torch.long)
    x_T = O.q_sample(O.make_betas(**O.TINY["diffusion"]), z_mode, t, torch.from_numpy(g["q_noise"]))
    assert O.max_rel_err(x_T, torch.from_numpy(g["x_T"])) < 1e-5


def test_oracle_wavelet_matches_reference():
    g = np.load(os.path.join(GOLD, "golden_wavelet.npz"))
    out = O.wavelet_reconstruction(torch.from_numpy(g["content"]), torch.from_numpy(g["style"]))
    assert O.max_rel_err(out, torch.from_numpy(g["out"])) < 1e-6


def test_oracle_tiled_vae_encode_matches_reference():
    """vae_encode(tiled=True) of the live reference (VAEHook on the encoder, pad 32) vs the restatement."""
    g = np.load(os.path.join(GOLD, "golden_vae_encode_tiled.npz"))
    sd = O.make_weights(O.vae_encoder_param_shapes(O.TINY_VAE8), seed=3)
    image = torch.from_numpy(g["image"].astype(np.float32))
    with torch.no_grad():
        z = O.vae_encode_tiled(sd, O.TINY_VAE8, image, 0.18215, int(g["tile_size"]))
        z_untiled = O.vae_encode(sd, O.TINY_VAE8, image, 0.18215)
    assert O.max_rel_err(z, torch.from_numpy(g["z"])) < 1e-5
    assert O.max_rel_err(z, z_untiled) > 1e-3


def test_oracle_swinir_matches_reference():
    """SwinIR pre-restoration network (model/swinir.py:856-894): the restatement against the
    live-reference fixture (tests/golden/make_golden.py --swinir), square multi-window and rectangular one-window-high
    inputs (plain and shifted window attention, relative position bias, nearest+conv x8 upsampler)."""
    from oracle import swinir_oracle as S

    d = np.load(os.path.join(GOLD, "golden_swinir.npz"))
    sd = S.make_swinir_weights(S.SWINIR_TINY)
    for i in range(2):
        x, ref = torch.from_numpy(d[f"x{i}"]), torch.from_numpy(d[f"y{i}"])
        with torch.no_grad():
            y = S.swinir_forward(sd, S.SWINIR_TINY, x)
        assert y.shape == ref.shape == x.shape
        assert O.max_rel_err(y, ref) < 1e-5


def test_swinir_parameter_enumeration_and_flops():
    from oracle import swinir_oracle as S

    n = sum(int(np.prod(s)) for _, s in S.swinir_param_shapes(S.SWINIR_EDTR))
    assert 15.5e6 < n < 16.5e6          # 15.8 M parameters at the EDTR widths (embed 180, 8 x 6 blocks)
    # 48 Swin blocks at 64x64 tokens (111 GF) + RSTB / body convs (21) + conv_first + the x8 nearest+conv upsampler (47)
    assert abs(S.swinir_gflops(S.SWINIR_EDTR, 512, 512) - 181.5) < 1.0
    assert S.block_geometry(S.SWINIR_EDTR, 0) == (8, 0) and S.block_geometry(S.SWINIR_EDTR, 1) == (8, 4)
    assert S.block_geometry(dict(S.SWINIR_TINY, img_size=8), 1) == (8, 0)
