"""Dev tool (GPU box): time the SwinIR (EDTR configuration) forward graph at B=8, 512x512 and list kernel time by name."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edtr_b200.swinir import SwinIR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = SwinIR(img_size=64, patch_size=1, in_chans=3, embed_dim=180, depths=[6] * 8, num_heads=[6] * 8, window_size=8, mlp_ratio=2,
           sf=8, img_range=1.0, upsampler="nearest+conv", resi_connection="1conv", unshuffle=True, unshuffle_scale=8).cuda().eval()
x = torch.rand(B, 3, 512, 512, device="cuda")
for _ in range(3):
    y = m(x)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    y = m(x)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print(f"SwinIR forward B={B} 512x512: {ms:.3f} ms  ({181.5 * B / ms:.1f} TFLOP/s on 181.5 GF/img)")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    y = m(x)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name[:70]][0] += 1
        agg[ev.name[:70]][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"kernel time {tot / 1e3:.3f} ms in {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%  x{v[0]:<5d} avg {v[1] / v[0]:8.1f} us  {k}")
