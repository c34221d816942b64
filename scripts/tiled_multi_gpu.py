"""Config C4 check (run under torchrun on N GPUs): a cldm-tiled step with the latent tiles spread over the ranks
(one all-reduce of the partial blend per step) must equal the single-rank result.  Also times a 256x256-latent
(2048x2048 image) 4-step tiled sample at s4 widths."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from edtr_b200.sampler import SpacedSampler  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model = bench.build_model(dev)
    g = torch.Generator().manual_seed(5)
    L = int(os.environ.get("LATENT", "128"))
    x = torch.randn(1, 4, L, L, generator=g).to(dev)
    cond = {"c_img": (0.8 * torch.randn(1, 4, L, L, generator=g)).to(dev), "c_txt": torch.randn(1, 77, 1024, generator=g).to(dev)}
    t = torch.full((1,), 200, dtype=torch.long, device=dev)
    eps_multi = model.forward_tiled(x, t, cond, 64, 32)
    # single-rank reference on every rank (no process-group sharding)
    eps_single = model.engine().forward_tiled(x, t, cond["c_img"], cond["c_txt"], 64, 32, control_scales=model.control_scales)
    err = float((eps_multi - eps_single).abs().max() / eps_single.abs().max())
    betas = (torch.linspace(0.00085 ** 0.5, 0.0120 ** 0.5, 1000, dtype=torch.float64) ** 2).numpy()
    sampler = SpacedSampler(betas)
    torch.manual_seed(7)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.time()
    z = sampler.manual_sample_with_timesteps(model, dev, x, 4, [50, 100, 150, 200], 1, cond, None, 1.0, tiled=True,
                                             tile_size=64, tile_stride=32, progress=False)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.time() - t0
    # tiled VAE decode of the sampled latent (VAEHook semantics): tiles over ranks, statistics + image all-reduced
    vt = int(os.environ.get("VAE_TILE", "64"))
    img_multi = model.vae_decode(z, tiled=True, tile_size=vt)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.time()
    img_multi = model.vae_decode(z, tiled=True, tile_size=vt)
    torch.cuda.synchronize()
    dist.barrier()
    dt_vae = time.time() - t0
    img_single = model.vae._decoder_engine().decode_tiled(z.float().contiguous(), float(model.scale_factor), vt)
    mse = torch.mean(((img_multi - img_single).double() / 2) ** 2)
    psnr = float(10.0 * torch.log10(1.0 / (mse + 1e-8)))
    if rank == 0:
        print(f"tiled VAE decode: latent {L}x{L} -> image {8 * L}x{8 * L}, tile {vt}, world {world}: {dt_vae * 1e3:.1f} ms, "
              f"multi-rank vs single-rank PSNR {psnr:.1f} dB, finite {bool(torch.isfinite(img_multi).all())}")
        assert psnr > 45.0
    if rank == 0:
        print(f"tiled C4 check: world {world}, latent {L}x{L}, multi-rank vs single-rank max-rel {err:.2e}, "
              f"4-step tiled sample (first call, includes graph capture) {dt:.2f} s, finite {bool(torch.isfinite(z).all())}")
        assert err < 1e-2  # batch composition changes split-K / summation order only
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
