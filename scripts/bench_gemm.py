"""Dev tool (GPU box): time individual GEMM / conv shapes of the path on both tensor-core kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edtr_b200 import ops  # noqa: E402

BF = torch.bfloat16
SHAPES = [
    # kind, args  (conv: B,H,W,Cin,Cout ; gemm: M,N,K)
    ("conv", (8, 64, 64, 320, 320)), ("conv", (8, 32, 32, 640, 640)), ("conv", (8, 16, 16, 1280, 1280)),
    ("conv", (8, 8, 8, 1280, 1280)), ("conv", (8, 64, 64, 960, 320)), ("conv", (8, 16, 16, 2560, 1280)),
    ("conv", (8, 512, 512, 128, 128)), ("conv", (8, 256, 256, 256, 256)), ("conv", (8, 128, 128, 512, 512)),
    ("conv", (8, 64, 64, 512, 512)),
    ("gemm", (32768, 320, 320)), ("gemm", (32768, 960, 320)), ("gemm", (32768, 2560, 320)), ("gemm", (32768, 320, 1280)),
    ("gemm", (8192, 640, 640)), ("gemm", (8192, 5120, 640)), ("gemm", (8192, 640, 2560)),
    ("gemm", (2048, 1280, 1280)), ("gemm", (2048, 10240, 1280)), ("gemm", (2048, 1280, 5120)),
    ("gemm", (512, 1280, 1280)), ("gemm", (512, 10240, 1280)),
    ("gemm_nores", (32768, 320, 320)), ("gemm_nores", (32768, 960, 320)), ("gemm_nores", (8192, 640, 640)),
    ("geglu", (32768, 2560, 320)), ("geglu", (8192, 5120, 640)), ("geglu", (2048, 10240, 1280)),
]


def run(kind, a, iters=20):
    g = torch.Generator(device="cuda").manual_seed(0)
    if kind == "conv":
        B, H, W, Ci, Co = a
        x = torch.randn(B, H, W, Ci, generator=g, device="cuda").to(BF)
        w = (torch.randn(Co, 9 * Ci, generator=g, device="cuda") * (9 * Ci) ** -0.5).to(BF)
        bias = torch.randn(Co, device="cuda")
        out = torch.empty(B, H, W, Co, dtype=BF, device="cuda")
        fn = lambda: ops.conv3x3(x, w, bias=bias, out=out)
        fl = 2.0 * B * H * W * Co * 9 * Ci
    elif kind == "geglu":
        M, N, K = a
        x = torch.randn(M, K, generator=g, device="cuda").to(BF)
        w = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(BF)
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N // 2, dtype=BF, device="cuda")
        fn = lambda: ops.gemm(x, w, bias=bias, act=ops.ACT_GEGLU, out=out)
        fl = 2.0 * M * N * K
    elif kind == "gemm_nores":
        M, N, K = a
        x = torch.randn(M, K, generator=g, device="cuda").to(BF)
        w = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(BF)
        out = torch.empty(M, N, dtype=BF, device="cuda")
        fn = lambda: ops.gemm(x, w, out=out)
        fl = 2.0 * M * N * K
    else:
        M, N, K = a
        x = torch.randn(M, K, generator=g, device="cuda").to(BF)
        w = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(BF)
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, generator=g, device="cuda").to(BF)
        out = torch.empty(M, N, dtype=BF, device="cuda")
        fn = lambda: ops.gemm(x, w, bias=bias, residual=res, out=out)
        fl = 2.0 * M * N * K
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    return ms, fl / ms / 1e9


def bench_attention():
    print(f"{'attention B,heads,Lq,Lk':40s} {'ms':>9s} {'TF/s':>9s}")
    for B, h, Lq, Lk in ((8, 5, 4096, 4096), (8, 10, 1024, 1024), (8, 20, 256, 256), (8, 5, 4096, 77), (8, 20, 64, 64)):
        g = torch.Generator(device="cuda").manual_seed(0)
        C = h * 64
        q = torch.randn(B, Lq, C, generator=g, device="cuda").to(BF)
        k = torch.randn(B, Lk, C, generator=g, device="cuda").to(BF)
        v = torch.randn(B, Lk, C, generator=g, device="cuda").to(BF)
        out = torch.empty(B, Lq, C, dtype=BF, device="cuda")
        fn = lambda: ops.attention(q, k, v, h, 0.125, out=out)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print(f"{str((B, h, Lq, Lk)):40s} {ms:9.3f} {4.0 * B * h * Lq * Lk * 64 / ms / 1e9:9.1f}", flush=True)


def main():
    only = sys.argv[1:]
    if not only or "attention" in only:
        bench_attention()
    print(f"{'shape':40s} {'v2 ms':>9s} {'v2 TF/s':>9s} {'v1 ms':>9s} {'v1 TF/s':>9s}")
    for kind, a in SHAPES:
        if only and kind not in only:
            continue
        os.environ["EDTR_GEMM_V1"] = "0"
        m2, t2 = run(kind, a)
        os.environ["EDTR_GEMM_V1"] = "1"
        m1, t1 = run(kind, a)
        os.environ["EDTR_GEMM_V1"] = "0"
        print(f"{kind + str(a):40s} {m2:9.3f} {t2:9.1f} {m1:9.3f} {t1:9.1f}", flush=True)


if __name__ == "__main__":
    main()
