"""Dev tool (GPU box): time the fp32 mode (CldmEngineF32 + VaeDecoderF32, s4 widths) at B = 1 and B = 4 and compare one
step with the bf16 engine on the same inputs."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cldm_oracle as O  # noqa: E402  (dev tool: synthetic weights / inputs only)
from edtr_b200.engine import CldmEngine  # noqa: E402
from edtr_b200.engine_f32 import CldmEngineF32, VaeDecoderF32  # noqa: E402

cfg = O.S4
w = O.make_cldm_weights(cfg, seed=0)
dev = torch.device("cuda")
e32 = CldmEngineF32(cfg["unet"], cfg["controlnet"], w["unet"], w["controlnet"], dev)
e16 = CldmEngine(cfg["unet"], cfg["controlnet"], w["unet"], w["controlnet"], dev)
dd = dict(double_z=True, z_channels=4, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, out_ch=3, in_channels=3, attn_resolutions=[])
v32 = VaeDecoderF32(dd, 4, w["vae"], dev)
for B in (1, 4):
    x_T, cond, _ = O.make_inputs(cfg, B, 64, seed=1)
    x, ci, ct = x_T.to(dev), cond["c_img"].to(dev), cond["c_txt"].to(dev)
    t = torch.full((B,), 200, dtype=torch.long, device=dev)
    eps = e32.forward(x, t, ci, ct)
    torch.cuda.synchronize()
    t0 = time.time()
    eps = e32.forward(x, t, ci, ct)
    torch.cuda.synchronize()
    ms_step = (time.time() - t0) * 1e3
    ref = e16.forward(x, t, ci, ct)
    rel = ((eps - ref).abs().max() / eps.abs().max()).item()
    z = torch.randn(B, 4, 64, 64, device=dev)
    img = v32.decode(z, 0.18215)
    torch.cuda.synchronize()
    t0 = time.time()
    img = v32.decode(z, 0.18215)
    torch.cuda.synchronize()
    ms_dec = (time.time() - t0) * 1e3
    gf = 1073.38 * B
    print(f"fp32 mode B={B}: ControlLDM step {ms_step:.1f} ms ({gf / ms_step:.2f} TFLOP/s), VAE decode {ms_dec:.1f} ms "
          f"({2514.52 * B / ms_dec:.2f} TFLOP/s) -> restore {4 * ms_step + ms_dec:.0f} ms = {B / (4 * ms_step + ms_dec) * 1e3:.2f} img/s; "
          f"bf16 engine vs fp32 engine eps max-rel {rel:.2e}; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
