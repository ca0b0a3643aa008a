"""1x1 residual convolution (ResidualBlock.residual_conv) forward and data gradient: flat fdx_gemm launch vs the
convolution-geometry launch (fdx_conv1x1_*, eligible for the transposed engine).  Checks both agree, prints
time and effective HBM rate.   python tests/gpu_bench_res1x1.py [res] [batch]"""
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402
from flaxdiff_b200._lib import GEMM_KK, GEMM_KMN  # noqa: E402

dev = torch.device("cuda")
res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
# (resolution divisor, Cin, Cout) of the UNet's residual convs (up path + middle)
SHAPES = [(1, 128, 64), (1, 192, 64), (1, 320, 64), (2, 192, 128), (2, 384, 128), (2, 576, 256), (4, 384, 256), (4, 768, 512),
          (8, 1024, 512)]


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


tot = [0.0, 0.0, 0.0, 0.0]
for div, cin, cout in SHAPES:
    h = res // div
    M = B * h * h
    x = torch.randn(B, h, h, cin, device=dev).bfloat16()
    w = (torch.randn(cin, cout, device=dev) / cin ** 0.5).bfloat16()
    bias = torch.randn(cout, device=dev)
    dout = torch.randn(B, h, h, cout, device=dev).bfloat16()
    base = torch.randn(B, h, h, cin, device=dev).bfloat16()
    r1 = torch.empty(B, h, h, cout, device=dev, dtype=torch.bfloat16)
    r2 = torch.empty_like(r1)
    f_g = lambda: ops.gemm(GEMM_KMN, x, w, r1, M, cout, cin, cin, cout, cout, bias=bias)
    f_c = lambda: ops.conv1x1_fwd(x, w, bias, out=r2)
    dx1, dx2 = base.clone(), base.clone()
    d_g = lambda: ops.gemm(GEMM_KK, dout, w, dx1, M, cin, cout, cout, cout, cin, res=dx1, r_ld=cin)
    d_c = lambda: ops.conv1x1_dgrad(dout, w, dx2, accumulate=True)
    f_g(); f_c(); d_g(); d_c()
    torch.cuda.synchronize()
    ef = ((r1.float() - r2.float()).norm() / r1.float().norm()).item()
    ed = ((dx1.float() - dx2.float()).norm() / dx1.float().norm()).item()
    t = [timeit(f_g), timeit(f_c), timeit(d_g), timeit(d_c)]
    for i in range(4):
        tot[i] += t[i]
    fb = M * (cin + cout) * 2 / 1e9
    db = M * (cout + 2 * cin) * 2 / 1e9
    print(f"{h:4d}x{h:<4d} {cin:4d}->{cout:<4d} | fwd gemm {t[0]*1e3:7.1f} us conv {t[1]*1e3:7.1f} us ({fb/t[1]:5.2f} TB/s) diff {ef:.1e} | "
          f"dgrad gemm {t[2]*1e3:7.1f} us conv {t[3]*1e3:7.1f} us ({db/t[3]:5.2f} TB/s) diff {ed:.1e}", flush=True)
print(f"TOTAL fwd gemm {tot[0]:.3f} ms conv {tot[1]:.3f} ms | dgrad gemm {tot[2]:.3f} ms conv {tot[3]:.3f} ms")
