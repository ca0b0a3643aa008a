"""Transposed small-Cout kernels (fdx_tct.cu; baseline = FDX_NO_TCT=1) against the generic engine on full-size layers:
outputs must agree to bf16 rounding (same products, different summation order), plus per-layer timing.
    python tests/gpu_tct_check.py [res] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")
# (res divisor, Cin, Cout): forward needs Cout in {64,128}; dgrad needs Cin in {64,128}
SHAPES = [(1, 64, 64), (1, 128, 64), (1, 320, 64), (2, 128, 128), (2, 192, 128), (2, 576, 128), (2, 64, 128),
          (4, 128, 128), (1, 64, 128), (2, 128, 192), (4, 384, 256), (1, 64, 320)]


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


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    tot = [0.0, 0.0, 0.0, 0.0]
    for div, cin, cout in SHAPES:
        h = res // div
        x = torch.randn(B, h, h, cin, device=dev).bfloat16()
        w = (torch.randn(3, 3, cin, cout, device=dev) / (3 * cin ** 0.5)).bfloat16()
        bias = torch.randn(cout, device=dev)
        row = torch.randn(B, cout, device=dev)
        resid = torch.randn(B, h, h, cout, device=dev).bfloat16()
        dy = torch.randn(B, h, h, cout, device=dev).bfloat16()
        outs, times, stats = [], [], []
        for tct in (False, True):
            if tct:
                os.environ.pop("FDX_NO_TCT", None)
            else:
                os.environ["FDX_NO_TCT"] = "1"
            y = torch.empty(B, h, h, cout, device=dev, dtype=torch.bfloat16)
            dx = torch.empty_like(x)
            cs = ops.ColStats(B, cout, dev)
            ops.conv3x3_fwd(x, w, bias, rowvec=row, res=resid, out=y, colstats=(cs, 0))
            ops.conv3x3_dgrad(dy, w, dx)
            torch.cuda.synchronize()
            outs.append((y.clone(), dx.clone()))
            stats.append((ops.groupnorm_stats_from_cols(cs, 8), ops.groupnorm_stats(y, 8)))
            times.append((timeit(lambda: ops.conv3x3_fwd(x, w, bias, rowvec=row, res=resid, out=y)),
                          timeit(lambda: ops.conv3x3_dgrad(dy, w, dx))))
        ef, ed = rel(outs[1][0], outs[0][0]), rel(outs[1][1], outs[0][1])
        es = rel(stats[1][0], stats[1][1])
        fl = 2.0 * B * h * h * 9 * cin * cout / 1e9
        print(f"{h:4d}x{h:<4d} {cin:4d}->{cout:<4d} | fwd {times[0][0]:7.3f} -> {times[1][0]:7.3f} ms ({fl/times[1][0]:6.0f} TF/s) err {ef:.1e} "
              f"stats {es:.1e} | dgrad {times[0][1]:7.3f} -> {times[1][1]:7.3f} ms ({fl/times[1][1]:6.0f} TF/s) err {ed:.1e}", flush=True)
        for i, v in enumerate((times[0][0], times[1][0], times[0][1], times[1][1])):
            tot[i] += v
        del x, w, dy, resid
    print(f"TOTAL fwd {tot[0]:.3f} -> {tot[1]:.3f} ms | dgrad {tot[2]:.3f} -> {tot[3]:.3f} ms")


if __name__ == "__main__":
    main()
