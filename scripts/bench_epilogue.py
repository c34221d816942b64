"""Dev tool (GPU box): cost of the epilogue features of the CTA-pair GEMM on the short-K shapes of the 64x64 level
(graph of 20 launches per variant, so the host launch path does not count)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edtr_b200 import ops  # noqa: E402
from edtr_b200.engine import fold_layernorm, geglu_permutation  # noqa: E402

BF = torch.bfloat16


def variants(M, N, K, geglu):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, K, generator=g, device="cuda").to(BF)
    w = torch.randn(N, K, generator=g, device="cuda") * K ** -0.5
    b = torch.randn(N, generator=g, device="cuda") * 0.1
    gamma = 1 + 0.1 * torch.randn(K, generator=g, device="cuda")
    beta = 0.1 * torch.randn(K, generator=g, device="cuda")
    if geglu:
        perm = geglu_permutation(N // 2, ops.geglu_tile_n()).cuda()
        w, b = w[perm].contiguous(), b[perm].contiguous()
    wg, cs, bf = fold_layernorm(w, b, gamma, beta, "cuda")
    wb = w.to(BF).contiguous()
    parts = ops.row_stats_parts(M, K, K)
    xf = x.float()
    stats = torch.stack([xf.sum(-1, keepdim=True).expand(M, parts) / parts,
                         (xf * xf).sum(-1, keepdim=True).expand(M, parts) / parts], -1).contiguous()
    n_out = N // 2 if geglu else N
    out = torch.empty(M, n_out, dtype=BF, device="cuda")
    res = torch.randn(M, n_out, generator=g, device="cuda").to(BF)
    act = ops.ACT_GEGLU if geglu else ops.ACT_NONE
    v = {
        "plain": lambda: ops.gemm(x, wb, act=act, out=out) if not geglu else ops.gemm(x, wb, bias=b, act=act, out=out),
        "bias": lambda: ops.gemm(x, wb, bias=b, act=act, out=out),
        "bias+ln": lambda: ops.gemm(x, wg, bias=bf, ln=(stats, K, 1e-5, cs), act=act, out=out),
    }
    if not geglu:
        rs = torch.empty(M, ops.row_stats_parts(M, N, K), 2, device="cuda")
        v["bias+res"] = lambda: ops.gemm(x, wb, bias=b, residual=res, out=out)
        v["bias+res+rowstats"] = lambda: ops.gemm(x, wb, bias=b, residual=res, row_stats=rs, out=out)
    return v


def time_graph(fn, reps=20):
    fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (5 * reps) * 1e3


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    shapes = [(32768, 960, 320, False), (32768, 320, 320, False), (32768, 2560, 320, True), (8192, 1920, 640, False),
              (8192, 5120, 640, True), (2048, 3840, 1280, False)]
    for M, N, K, geglu in shapes:
        vs = variants(M, N, K, geglu)
        if only:   # one launch for ncu
            if (M, N) == (32768, 2560 if "geglu" in only else 960):
                vs["bias+ln"]()
                vs["bias"]()
                torch.cuda.synchronize()
            continue
        line = f"gemm {M}x{N}x{K}{' geglu' if geglu else ''}:"
        for name, fn in vs.items():
            line += f"  {name} {time_graph(fn):7.1f} us"
        print(line, flush=True)


if __name__ == "__main__":
    main()
