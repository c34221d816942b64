"""Dev tool (GPU box): time the captured sample / decode graphs and aggregate kernel time by
(kernel, grid) with torch.profiler.  Writes gpurun_out/profile_step.txt.  Not part of the product."""
import argparse
import collections
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edtr_b200 import topology as T  # noqa: E402
from edtr_b200.engine import CldmEngine, VaeDecoderEngine  # noqa: E402

S4_NET = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1), num_res_blocks=2,
              channel_mult=(1, 2, 4, 4), num_head_channels=64, context_dim=1024)
S4_DD = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4),
             num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def rand_sd(shapes, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    sd = {}
    for k, shp in shapes:
        if k.endswith("weight") and len(shp) >= 2:
            fan = 1
            for d in shp[1:]:
                fan *= d
            sd[k] = torch.randn(shp, generator=g, device="cuda") * fan ** -0.5
        elif k.endswith("weight"):
            sd[k] = 1 + 0.1 * torch.randn(shp, generator=g, device="cuda")
        else:
            sd[k] = 0.05 * torch.randn(shp, generator=g, device="cuda")
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/profile_step.txt")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--shapes", action="store_true", help="per-shape kernel time of one graph replay (single stream)")
    ap.add_argument("--no-overlap", action="store_true", help="run the ControlNet on the main stream")
    a = ap.parse_args()
    B = a.batch
    cn_cfg = dict(S4_NET, hint_channels=4)
    t0 = time.time()
    eng = CldmEngine(S4_NET, cn_cfg, rand_sd(T.unet_param_shapes(S4_NET), 0),
                     rand_sd(T.unet_param_shapes(cn_cfg, True), 1), "cuda")
    vd = VaeDecoderEngine(S4_DD, 4, rand_sd(T.vae_decoder_param_shapes(S4_DD, 4), 2), "cuda")
    print("pack s", time.time() - t0, flush=True)
    if a.no_overlap or a.shapes:   # launch order == time order, so kernels can be attributed to ops
        eng.overlap = False
    g = torch.Generator(device="cuda").manual_seed(3)
    x_T = torch.randn(B, 4, 64, 64, generator=g, device="cuda")
    c_img = 0.8 * torch.randn(B, 4, 64, 64, generator=g, device="cuda")
    c_txt = torch.randn(B, 77, 1024, generator=g, device="cuda")
    noise = [torch.randn(B, 4, 64, 64, generator=g, device="cuda") for _ in range(4)]
    # s4 schedule tables (values irrelevant for timing)
    tabs = {k: torch.rand(4, device="cuda") for k in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                                                      "posterior_mean_coef1", "posterior_mean_coef2",
                                                      "posterior_variance")}
    ts = [200, 150, 100, 50]
    lines = []

    def timeit(name, fn):
        fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(a.iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        ms = sorted(s.elapsed_time(e) for s, e in evs)
        lines.append(f"{name}: median {ms[len(ms) // 2]:.3f} ms  min {ms[0]:.3f} ms  (B={B})")
        print(lines[-1], flush=True)
        return ms[len(ms) // 2]

    t0 = time.time()
    z = eng.sample(x_T, ts, tabs, c_img, c_txt, noise)
    torch.cuda.synchronize()
    print("first sample (warm-up + capture) s", time.time() - t0, "finite", bool(torch.isfinite(z).all()), flush=True)
    t0 = time.time()
    img = vd.decode(z, 0.18215)
    torch.cuda.synchronize()
    print("first decode s", time.time() - t0, "finite", bool(torch.isfinite(img).all()), flush=True)
    ms_s = timeit("sample(4 steps) graph", lambda: eng.sample(x_T, ts, tabs, c_img, c_txt, noise))
    ms_d = timeit("vae decode graph", lambda: vd.decode(z, 0.18215))
    lines.append(f"restore: {ms_s + ms_d:.3f} ms per batch -> {B / (ms_s + ms_d) * 1e3:.2f} img/s; "
                 f"tensor roofline frac {B / (ms_s + ms_d) * 1e3 * 6808.0e9 / 1417.6e12:.3f} (sustained 1417.6 TF)")
    print(lines[-1], flush=True)
    ws = eng.workspace(B, 64, 64)
    lines.append(f"weights {eng.unet.nbytes() / 1e9:.2f}+{eng.cnet.nbytes() / 1e9:.2f} GB, workspace {ws.nbytes() / 1e9:.2f} GB, "
                 f"mem allocated {torch.cuda.memory_allocated() / 1e9:.2f} GB")
    if a.shapes:
        # 1) eager pass: record the op sequence with shapes; 2) profile one graph replay and attribute the
        # kernel durations (CUPTI) to the ops by launch order.
        from torch.profiler import ProfilerActivity, profile

        from edtr_b200 import ops

        seq = []
        names = ("gemm", "conv3x3", "conv3x3_up2x", "attention", "groupnorm", "groupnorm_fold", "layernorm", "softmax_rows", "upsample2x",
                 "im2col", "nchw_to_nhwc", "pointwise_nchw_to_nhwc", "cast_bf16", "timestep_embedding", "sampler_update",
                 "tile_blend")
        orig = {n: getattr(ops, n) for n in names}
        # kernels an op launches first ("primary", counted) — the rest (split-K reduce, GN apply) follow their primary
        PRIMARY = {"gemm": ("gemm2_kernel", "gemm_conv_kernel"), "conv3x3": ("gemm2_kernel", "gemm_conv_kernel"),
                   "conv3x3_up2x": ("gemm2_kernel",), "attention": ("attention_ts_kernel", "cross_attention_small_kernel"),
                   "groupnorm": ("groupnorm_stats_kernel", "groupnorm_fused_kernel"),
                   "groupnorm_fold": ("groupnorm_fold_kernel",),   # + the apply kernel (secondary); statistics from the producer
                   "layernorm": ("layernorm_kernel",),
                   "softmax_rows": ("softmax_rows_kernel",), "upsample2x": ("upsample2x_kernel",),
                   "im2col": ("im2col_kernel",), "nchw_to_nhwc": ("nchw_to_nhwc_kernel",),
                   "pointwise_nchw_to_nhwc": ("pointwise_nchw_kernel",), "cast_bf16": ("cast_f32_bf16_kernel",),
                   "timestep_embedding": ("timestep_embedding_kernel",), "sampler_update": ("sampler_update_kernel",),
                   "tile_blend": ("tile_blend_kernel",)}
        SECONDARY = ("splitk_reduce_kernel", "groupnorm_apply_kernel")

        def wrap(name, fn):
            def w(*args, **kw):
                nk = 1
                if name == "gemm":
                    A, Wt = args[0], args[1]
                    M = A.numel() // A.shape[-1]
                    flag = "geglu" if kw.get("act") == 2 else "res" if kw.get("residual") is not None else ""
                    key = (name, M, Wt.shape[0], Wt.shape[1], flag)
                    fl = 2.0 * M * Wt.shape[0] * Wt.shape[1]
                elif name == "conv3x3":
                    X, Wt = args[0], args[1]
                    M = X.numel() // X.shape[-1]
                    key = (name, M, Wt.shape[0], Wt.shape[1], "res" if kw.get("residual") is not None else "")
                    fl = 2.0 * M * Wt.shape[0] * Wt.shape[1]
                elif name == "conv3x3_up2x":   # 4 phase GEMMs; FLOPs counted as the reference's 3x3 conv on the 2x grid
                    X, Wt = args[0], args[1]
                    M = X.numel() // X.shape[-1]
                    key = (name, 4 * M, Wt.shape[1], 9 * X.shape[-1], "")
                    fl = 2.0 * 4 * M * Wt.shape[1] * 9 * X.shape[-1]
                    nk = 4
                elif name == "attention":
                    q, k_ = args[0], args[1]
                    key = (name, q.shape[0] * q.shape[1], k_.shape[1], q.shape[2], "")
                    fl = 4.0 * q.shape[0] * q.shape[1] * k_.shape[1] * q.shape[2]
                elif name == "groupnorm_fold":
                    pt = args[0]            # [B, slabs, C/unit, 2], C
                    key = ("groupnorm_fold", pt.shape[0] * pt.shape[1] * 32, args[1], 0, "+apply")
                    fl = 4.0 * pt.shape[0] * pt.shape[1] * 32 * args[1]
                elif name in ("groupnorm", "layernorm"):
                    x = args[0]
                    key = (name, x.numel() // x.shape[-1], x.shape[-1], 0, "")
                    fl = 4.0 * x.numel()   # bytes: bf16 in + out
                else:
                    key = (name, 0, 0, 0, "")
                    fl = 0.0
                seq.append((key, fl, nk))
                return fn(*args, **kw)
            return w

        for phase, eager, replay in (
                ("sample", lambda: eng.sample(x_T, ts, tabs, c_img, c_txt, noise, use_graph=False),
                 lambda: eng.sample(x_T, ts, tabs, c_img, c_txt, noise)),
                ("decode", lambda: vd.decode(z, 0.18215, use_graph=False), lambda: vd.decode(z, 0.18215))):
            seq.clear()
            for n, fn in orig.items():
                setattr(ops, n, wrap(n, fn))
            try:
                eager()
            finally:
                for n, fn in orig.items():
                    setattr(ops, n, fn)
            torch.cuda.synchronize()
            replay()
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                replay()
                torch.cuda.synchronize()
            evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "edtr::" in e.name]
            evs.sort(key=lambda e: e.time_range.start)
            agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
            i, left, bad, last = 0, (seq[0][2] if seq else 0), 0, None
            for e in evs:
                us = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
                if any(k in e.name for k in SECONDARY):
                    if last is not None:
                        agg[last][1] += us
                    continue
                while i < len(seq) and not any(k in e.name for k in PRIMARY[seq[i][0][0]]):
                    i += 1
                    bad += 1
                    left = seq[i][2] if i < len(seq) else 0
                if i >= len(seq):
                    bad += 1
                    continue
                key, fl, nk = seq[i]
                agg[key][1] += us
                last = key
                left -= 1
                if left == 0:
                    agg[key][0] += 1
                    agg[key][2] += fl
                    i += 1
                    left = seq[i][2] if i < len(seq) else 0
            tot = sum(v[1] for v in agg.values())
            span = (max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)) if evs else 0.0
            lines.append(f"--- {phase}: graph-replay kernel time by op/shape: {tot / 1e3:.2f} ms busy, {span / 1e3:.2f} ms span, "
                         f"{len(evs)} kernels, {len(seq)} ops, {bad} unmatched  (op, M, N, K|Lk, flag): count, ms, TFLOP/s "
                         f"(norms: GB/s), ms at the sustained bf16 peak")
            for key, (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
                ideal = fl / 1382.9e12 * 1e3 if key[0] not in ("groupnorm", "groupnorm_fold", "layernorm") else fl / 6539.2e9 * 1e3
                lines.append(f"{us / 1e3:8.3f} ms {100 * us / tot:5.1f}% x{n:<4d} avg {us / max(n, 1):7.1f} us "
                             f"{fl / max(us, 1e-9) / 1e6:8.1f}  ideal {ideal:7.3f} ms  {key}")
    if not a.no_profile:
        from torch.profiler import ProfilerActivity, profile

        for name, fn in (("sample", lambda: eng.sample(x_T, ts, tabs, c_img, c_txt, noise)),
                         ("decode", lambda: vd.decode(z, 0.18215))):
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                fn()
                torch.cuda.synchronize()
            agg = collections.defaultdict(lambda: [0, 0.0])
            total = 0.0
            for ev in prof.events():
                if ev.device_type == torch.autograd.DeviceType.CUDA:
                    k = ev.name[:70]
                    agg[k][0] += 1
                    agg[k][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
                    total += agg[k][1] * 0
            tot = sum(v[1] for v in agg.values())
            lines.append(f"--- {name}: kernel time {tot / 1e3:.3f} ms in {sum(v[0] for v in agg.values())} launches")
            for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
                lines.append(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  x{n:<5d} {k}")
            try:
                prof.export_chrome_trace(f"gpurun_out/trace_{name}.json")
            except Exception as ex:  # noqa: BLE001
                lines.append(f"trace export failed: {ex}")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
