"""CTA-pair (cta_group::2, FDX_PAIR=1) tensor-core kernels against the single-CTA ones on full-size layers:
bit-identical outputs expected (same accumulation order), and the per-layer timing of both.
    python tests/gpu_pair_check.py [res] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")
SHAPES = [(1, 128, 256), (1, 320, 128), (2, 128, 128), (2, 192, 128), (2, 576, 128), (2, 256, 512), (2, 256, 256),
          (4, 256, 256), (4, 384, 256), (4, 512, 512), (8, 512, 512), (8, 768, 512), (1, 64, 128), (1, 128, 192)]


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


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    tot = [0.0, 0.0, 0.0, 0.0]
    bad = 0
    for div, cin, cout in SHAPES:
        h = res // div
        x = torch.randn(B, h, h, cin, device=dev).bfloat16()
        w = (torch.randn(3, 3, cin, cout, device=dev) / (3 * cin ** 0.5)).bfloat16()
        bias = torch.randn(cout, device=dev)
        resid = torch.randn(B, h, h, cout, device=dev).bfloat16()
        dy = torch.randn(B, h, h, cout, device=dev).bfloat16()
        outs, times = [], []
        for pair in (False, True):
            if pair:
                os.environ["FDX_PAIR"] = "1"
            else:
                os.environ.pop("FDX_PAIR", None)
            y = torch.empty(B, h, h, cout, device=dev, dtype=torch.bfloat16)
            dx = torch.empty_like(x)
            ops.conv3x3_fwd(x, w, bias, res=resid, out=y)
            ops.conv3x3_dgrad(dy, w, dx)
            torch.cuda.synchronize()
            outs.append((y.clone(), dx.clone()))
            times.append((timeit(lambda: ops.conv3x3_fwd(x, w, bias, res=resid, out=y)),
                          timeit(lambda: ops.conv3x3_dgrad(dy, w, dx))))
        same_f = torch.equal(outs[0][0], outs[1][0])
        same_d = torch.equal(outs[0][1], outs[1][1])
        bad += (not same_f) + (not same_d)
        fl = 2.0 * B * h * h * 9 * cin * cout / 1e9
        print(f"{h:4d}x{h:<4d} {cin:4d}->{cout:<4d} | fwd {times[0][0]:7.3f} -> {times[1][0]:7.3f} ms ({fl/times[1][0]:6.0f} TF/s) "
              f"{'same' if same_f else 'DIFF'} | dgrad {times[0][1]:7.3f} -> {times[1][1]:7.3f} ms ({fl/times[1][1]:6.0f} TF/s) "
              f"{'same' if same_d else 'DIFF'}", flush=True)
        for i, v in enumerate((times[0][0], times[1][0], times[0][1], times[1][1])):
            tot[i] += v
        del x, w, dy, resid
    print(f"TOTAL fwd {tot[0]:.3f} -> {tot[1]:.3f} ms | dgrad {tot[2]:.3f} -> {tot[3]:.3f} ms | mismatches {bad}")


if __name__ == "__main__":
    main()
