"""Dev tool: one GroupNorm(+SiLU) at the 64x64x320 shape (two-pass kernels) for an ncu capture."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edtr_b200 import ops
B, HW, C = 8, 4096, 320
x = torch.randn(B, HW, C, device="cuda").to(torch.bfloat16)
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
out = torch.empty_like(x)
for _ in range(3):
    ops.groupnorm(x, g, b, 32, 1e-5, True, out=out)
torch.cuda.synchronize()
