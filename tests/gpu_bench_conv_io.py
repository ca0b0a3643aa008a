"""First / last convolution (3 <-> 64 channels): the im2col / col2im + tensor-core GEMM path against the direct
CUDA-core kernels (fdx_conv_in_fwd, fdx_conv_out_fwd, fdx_conv_out_dgrad).  python tests/gpu_bench_conv_io.py"""
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")


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
    return s.elapsed_time(e) / iters * 1e3


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


for B, R in ((256, 64), (64, 256)):
    x3 = torch.randn(B, R, R, 3, device=dev).bfloat16()
    w_in = torch.randn(3, 3, 3, 64, device=dev) * 0.2
    b_in = torch.randn(64, device=dev) * 0.1
    o1 = torch.empty(B, R, R, 64, device=dev, dtype=torch.bfloat16)
    o2 = torch.empty_like(o1)
    t_g = timeit(lambda: ops.conv_in_fwd(x3, w_in, b_in, o1))
    t_d = timeit(lambda: ops.conv_in_fwd_direct(x3, w_in, b_in, o2))
    print(f"{R}x{R} B={B} conv_in  fwd  : gemm {t_g:8.1f} us  direct {t_d:8.1f} us  diff {rel(o2, o1):.1e}", flush=True)
    a = torch.randn(B, R, R, 64, device=dev).bfloat16()
    w_out = torch.randn(3, 3, 64, 3, device=dev) * 0.05
    b_out = torch.randn(3, device=dev) * 0.1
    y1 = ops.conv_out_fwd(a, w_out, b_out)
    y2 = ops.conv_out_fwd_direct(a, w_out, b_out)
    t_g = timeit(lambda: ops.conv_out_fwd(a, w_out, b_out))
    t_d = timeit(lambda: ops.conv_out_fwd_direct(a, w_out, b_out))
    print(f"{R}x{R} B={B} conv_out fwd  : gemm {t_g:8.1f} us  direct {t_d:8.1f} us  diff {rel(y2, y1):.1e}", flush=True)
    dF = torch.randn(B, R, R, 3, device=dev)
    d1 = torch.empty(B, R, R, 64, device=dev, dtype=torch.bfloat16)
    d2 = torch.empty_like(d1)
    t_g = timeit(lambda: ops.conv_out_dgrad(dF, w_out, d1))
    t_d = timeit(lambda: ops.conv_out_dgrad_direct(dF, w_out, d2))
    print(f"{R}x{R} B={B} conv_out dgrad: gemm {t_g:8.1f} us  direct {t_d:8.1f} us  diff {rel(d2, d1):.1e}", flush=True)
