"""Image-parallel sharding of the restore path: one process per GPU, every rank restores a
contiguous slice of the batch with replicated weights and no per-step traffic; restored images and metric
scalars are all-gathered once at the end of a batch (the reference's ``accelerator.gather_for_metrics``,
main/det/test_edtr.py:163, main/cls/test_edtr.py:134-138)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def tile_sharding(tile_group):
    """(rank, world, reduce_fn) of the OPT-IN tile-parallel mode of the tiled paths (config C4).

    `tile_group` is None (default: no sharding, no communication — what the reference does, and the only valid choice
    when ranks hold different images, as in the image-parallel evaluation loops), True (the default process group) or a
    ``torch.distributed`` ProcessGroup.  Tile-parallelism requires that EVERY rank of the group calls with the SAME
    image / latent: rank r then evaluates tiles r, r+world, ... and the partial results are summed over the group."""
    if tile_group is None or tile_group is False:
        return 0, 1, None
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("tile_group is set but torch.distributed is not initialised")
    group = None if tile_group is True else tile_group
    world = dist.get_world_size(group)
    if world == 1:
        return 0, 1, None
    return dist.get_rank(group), world, (lambda buf: dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group))


def check_same_across_ranks(t: torch.Tensor, tile_group, what: str) -> None:
    """Debug guard for the tile-parallel mode: raises when the ranks of `tile_group` do not hold the same tensor
    (a 3-number fingerprint is max-/min-reduced; one tiny collective)."""
    rank, world, _ = tile_sharding(tile_group)
    if world == 1:
        return
    group = None if tile_group is True else tile_group
    tf = t.detach().float()
    fp = torch.stack([tf.sum(), tf.abs().sum(), tf.reshape(-1)[:: max(1, tf.numel() // 97)].sum()]).to(torch.float64)
    lo, hi = fp.clone(), fp.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if not torch.equal(lo, hi):
        raise RuntimeError(f"tile-parallel mode needs the same {what} on every rank of tile_group (they differ); "
                           "leave tile_group=None for image-parallel runs")


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` images for `rank`; the first total % world ranks take one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sliced_noise(shape, seed: int, steps: int, lo: int, hi: int, device) -> List[torch.Tensor]:
    """Per-step noise of the full batch drawn with one seed on every rank, then sliced, so an N-GPU run
    reproduces the 1-GPU result image by image (RNG note)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn(shape, generator=g)[lo:hi].to(device) for _ in range(steps)]


def gather_images(local: torch.Tensor, counts: List[int]) -> torch.Tensor:
    """All-gather variable-size image shards [b_r, C, H, W] into the full batch on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat([b[:n] for b, n in zip(bufs, counts)], 0)


class ImageGatherer:
    """The end-of-batch exchange of the image-parallel mode: ONE ``all_gather_into_tensor`` of the
    rank's restored images on a side stream, so that the NVLink transfer of batch i runs under the sampling of batch
    i+1 (the reference's ``accelerator.gather_for_metrics`` blocks the compute stream, main/cls/test_edtr.py:134-138).

    ``dtype``: wire / result dtype (bf16 halves the bytes; a [0, 1] image rounded to bf16 keeps > 55 dB PSNR against
    the fp32 image, far above any metric's resolution).  ``submit(img)`` returns at once; ``result()`` waits for the
    last submitted gather and returns the [world * B, C, H, W] tensor (valid until the next ``submit``)."""

    def __init__(self, shape, device, dtype=torch.bfloat16, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device)
        self.dtype = dtype
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.local = torch.empty(tuple(shape), dtype=dtype, device=self.device)
        self.out = torch.empty((self.world * shape[0],) + tuple(shape[1:]), dtype=dtype, device=self.device)
        self.bytes_per_rank = self.local.numel() * self.local.element_size()
        self._work = None

    def submit(self, img: torch.Tensor) -> None:
        if self.world == 1:
            self.out.copy_(img)
            return
        if self.stream is None:           # CPU / gloo (tests)
            self.local.copy_(img)
            dist.all_gather_into_tensor(self.out, self.local, group=self.group)
            return
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)       # the images are ready; also orders after the previous gather's readers
        with torch.cuda.stream(self.stream):
            self.local.copy_(img)          # cast to the wire dtype into a buffer the compute stream never touches
            self._work = dist.all_gather_into_tensor(self.out, self.local, group=self.group, async_op=True)
        img.record_stream(self.stream)

    def result(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.out
