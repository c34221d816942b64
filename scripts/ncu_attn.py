"""Dev tool: one attention launch (L=4096, 5 heads, B=8) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edtr_b200 import ops  # noqa: E402

B, h, L = 8, 5, 4096
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(B, L, h * 64, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3))
out = torch.empty_like(q)
for _ in range(4):
    ops.attention(q, k, v, h, 0.125, out=out)
torch.cuda.synchronize()
