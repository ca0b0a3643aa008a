"""Per-layer microbenchmark of the tcgen05 engine on the UNet's distinct contraction shapes
(SURVEY.md Appendix D).  Run under gpurun; prints one line per (shape, pass) with TFLOP/s.
    python tests/gpu_bench_layers.py [64|256] [batch]
"""
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")

# (out_res_div, Cin, Cout, stride) with res_div relative to the full resolution
SHAPES = [
    (1, 64, 64, 1), (1, 128, 64, 1), (1, 320, 64, 1), (1, 128, 256, 1),
    (2, 64, 64, 1), (2, 128, 128, 1), (2, 192, 128, 1), (2, 576, 128, 1), (2, 256, 512, 1),
    (4, 128, 128, 1), (4, 256, 256, 1), (4, 192, 256, 1), (4, 384, 256, 1), (4, 512, 64, 1),
    (8, 256, 256, 1), (8, 256, 512, 1), (8, 512, 512, 1), (8, 768, 512, 1),
    (2, 64, 64, 2), (4, 64, 128, 2), (8, 128, 256, 2),
]


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def gn_fusion(res, B):
    """dgrad + two-pass GroupNorm backward against the fused dgrad epilogue, per layer shape."""
    print(f"GN-backward fusion, res={res} B={B}: ms for [dgrad | gn_bwd 2-pass] vs [dgrad_gn | gn_bwd_dz]", flush=True)
    t = [0.0, 0.0, 0.0, 0.0]
    for div, cin, cout, stride in SHAPES:
        ho = res // div
        if stride != 1 or ho * ho < 128:
            continue
        x = (torch.randn(B, ho, ho, cin, device=dev) * 1.5 + 0.3).bfloat16()
        w = (torch.randn(3, 3, cin, cout, device=dev) / (3 * cin ** 0.5)).bfloat16()
        dy = torch.randn(B, ho, ho, cout, device=dev).bfloat16()
        da = torch.empty_like(x)
        dx = torch.empty_like(x)
        g, b = torch.ones(cin, device=dev), torch.zeros(cin, device=dev)
        dg, db = torch.zeros(cin, device=dev), torch.zeros(cin, device=dev)
        st = ops.groupnorm_stats(x, 8)
        ab = ops.groupnorm_coeffs(st, g, b, ho * ho, 1e-4)
        ws = ops.conv3x3_dgrad_gn(dy, w, da, x, ab)
        m0 = timeit(lambda: ops.conv3x3_dgrad(dy, w, da))
        m1 = timeit(lambda: ops.groupnorm_bwd(x, da, 8, st, g, b, 1e-4, True, dg, db, dx))
        m2 = timeit(lambda: (ops.groupnorm_coeffs(st, g, b, ho * ho, 1e-4), ops.conv3x3_dgrad_gn(dy, w, da, x, ab)))
        m3 = timeit(lambda: ops.groupnorm_bwd_dz(x, da, 8, st, g, 1e-4, ws, dg, db, dx))
        for i, m in enumerate((m0, m1, m2, m3)):
            t[i] += m
        print(f"{ho:4d}x{ho:<4d} {cin:4d}<-{cout:<4d} | {m0:7.3f} {m1:7.3f} = {m0+m1:7.3f} | {m2:7.3f} {m3:7.3f} = {m2+m3:7.3f}",
              flush=True)
        del x, w, dy, da, dx
    print(f"TOTAL | {t[0]:7.3f} {t[1]:7.3f} = {t[0]+t[1]:7.3f} | {t[2]:7.3f} {t[3]:7.3f} = {t[2]+t[3]:7.3f}", flush=True)


def gn_only(res, B):
    """GroupNorm forward / backward streaming rates over the UNet's tensor sizes."""
    tot = [0.0, 0.0, 0.0]
    for c, r in ((64, res), (128, res), (320, res), (128, res // 2), (192, res // 2), (576, res // 2),
                 (256, res // 4), (384, res // 4), (512, res // 4), (512, res // 8), (768, res // 8), (1024, res // 8)):
        x = (torch.randn(B, r, r, c, device=dev) * 1.5 + 0.3).bfloat16()
        dy = torch.randn(B, r, r, c, device=dev).bfloat16()
        dx = torch.empty_like(x)
        g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        st = ops.groupnorm_stats(x, 8)
        y = torch.empty_like(x)
        m1 = timeit(lambda: ops.groupnorm_stats(x, 8))
        m2 = timeit(lambda: ops.groupnorm_apply(x, 8, st, g, b, 1e-4, True, out=y))
        m3 = timeit(lambda: ops.groupnorm_bwd(x, dy, 8, st, g, b, 1e-4, True, dg, db, dx))
        nb = x.numel() * 2 / 1e6
        for i, m in enumerate((m1, m2, m3)):
            tot[i] += m
        print(f"GN {r:3d}x{r:<3d}x{c:<4d} {nb:7.1f} MB | stats {m1*1e3:7.1f} us {nb/m1/1e3:5.2f} TB/s | apply {m2*1e3:7.1f} us "
              f"{2*nb/m2/1e3:5.2f} TB/s | bwd {m3*1e3:7.1f} us {5*nb/m3/1e3:5.2f} TB/s", flush=True)
    print(f"TOTAL stats {tot[0]:.3f} ms apply {tot[1]:.3f} ms bwd {tot[2]:.3f} ms", flush=True)


def main():
    if len(sys.argv) > 3 and sys.argv[3] == "gnonly":
        return gn_only(int(sys.argv[1]), int(sys.argv[2]))
    if len(sys.argv) > 3 and sys.argv[3] == "gn":
        return gn_fusion(int(sys.argv[1]), int(sys.argv[2]))
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (256 if res == 64 else 64)
    only = sys.argv[3] if len(sys.argv) > 3 else None
    print(f"res={res} B={B}", flush=True)
    tot = {"fwd": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    for div, cin, cout, stride in SHAPES:
        ho = res // div
        hi = ho * stride
        x = torch.randn(B, hi, hi, cin, device=dev).bfloat16()
        w = (torch.randn(3, 3, cin, cout, device=dev) / (3 * cin ** 0.5)).bfloat16()
        y = torch.empty(B, ho, ho, cout, device=dev, dtype=torch.bfloat16)
        dy = torch.randn(B, ho, ho, cout, device=dev).bfloat16()
        dx = torch.empty_like(x)
        dw = torch.zeros(3, 3, cin, cout, device=dev)
        bias = torch.zeros(cout, device=dev)
        flops = 2.0 * B * ho * ho * 9 * cin * cout
        runs = {"fwd": lambda: ops.conv3x3_fwd(x, w, bias, out=y, stride=stride),
                "dgrad": lambda: ops.conv3x3_dgrad(dy, w, dx, stride=stride),
                "wgrad": lambda: ops.conv3x3_wgrad(x, dy, dw, stride=stride)}
        line = f"{ho:4d}x{ho:<4d} {cin:4d}->{cout:<4d} s{stride} {flops/1e9:8.1f} GF |"
        for k, fn in runs.items():
            if only and k != only:
                continue
            ms = timeit(fn)
            tf = flops / ms / 1e9
            tot[k][0] += flops
            tot[k][1] += ms
            line += f" {k} {ms:7.3f} ms {tf:7.1f} TF/s |"
        print(line, flush=True)
        del x, w, y, dy, dx, dw
    for k, (f, ms) in tot.items():
        if ms > 0:
            print(f"TOTAL {k}: {f/1e12:.2f} TFLOP in {ms:.2f} ms = {f/ms/1e9:.1f} TFLOP/s", flush=True)
    # GroupNorm streaming rate
    for c, r in ((64, res), (128, res // 2), (512, res // 8)):
        x = torch.randn(B, r, r, c, device=dev).bfloat16()
        g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        st = ops.groupnorm_stats(x, 8)
        y = torch.empty_like(x)
        ms1 = timeit(lambda: ops.groupnorm_stats(x, 8))
        ms2 = timeit(lambda: ops.groupnorm_apply(x, 8, st, g, b, 1e-4, True, out=y))
        nb = x.numel() * 2
        print(f"GN {r}x{r}x{c}: stats {nb/ms1/1e6:.0f} GB/s, apply {2*nb/ms2/1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
