"""Dev tool (GPU box): time the attention launches of the path (B=8) and check them against a fp32 torch softmax.
EDTR_ATT_VARIANT selects the kernel variant (read once per process)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edtr_b200 import ops  # noqa: E402

SHAPES = [  # B, heads, Lq, Lk
    (8, 5, 4096, 4096), (8, 10, 1024, 1024), (8, 20, 256, 256), (8, 20, 64, 64),
    (8, 5, 4096, 77), (8, 10, 1024, 77), (8, 20, 256, 77),
]


def run(B, h, Lq, Lk, iters=20):
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, Lq, h * 64, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Lk, h * 64, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Lk, h * 64, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(q)
    for _ in range(3):
        ops.attention(q, k, v, h, 0.125, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.attention(q, k, v, h, 0.125, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # parity on the first image / first two heads against fp32 softmax
    qf, kf, vf = (t[0].float().view(-1, h, 64).transpose(0, 1)[:2] for t in (q, k, v))
    ref = torch.softmax(qf @ kf.transpose(1, 2) * 0.125, -1) @ vf
    got = out[0].float().view(-1, h, 64).transpose(0, 1)[:2]
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    return ms, 4.0 * B * h * Lq * Lk * 64 / ms / 1e9, err


if __name__ == "__main__":
    print("variant", os.environ.get("EDTR_ATT_VARIANT", "default"))
    for s in SHAPES:
        ms, tf, err = run(*s)
        print(f"attention B={s[0]} h={s[1]:2d} Lq={s[2]:5d} Lk={s[3]:5d}: {ms * 1e3:8.1f} us {tf:8.1f} TFLOP/s  max-rel err {err:.2e}")
