"""Fused attention micro-benchmark at the BASELINE shapes (C3 self-attention at 32x32, C4 cross-attention to
77 text tokens at 128x128 / 64x64 / 32x32): fdx_attention_fwd / bwd time, TFLOP/s of the four (fwd) / ten
(bwd, incl. recomputation) GEMM-shaped products, and the HBM bytes a materialising path would have moved.
    python tests/gpu_bench_attention.py
"""
import sys

import torch

sys.path.insert(0, ".")
from flaxdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda")
# B, heads, L, Lk, dh(stored), d(true), label
SHAPES = [(64, 8, 1024, 1024, 32, 32, "C3 down_3 (256ch) self"), (64, 8, 1024, 1024, 64, 64, "C3 mid/up_0 (512ch) self"),
          (32, 8, 1024, 1024, 64, 64, "C5 B=32 self"), (128, 8, 16384, 77, 32, 16, "C4 level1 cross (128x128)"),
          (128, 8, 4096, 77, 32, 32, "C4 level2 cross (64x64)"), (128, 8, 1024, 77, 64, 64, "C4 level3 cross (32x32)")]


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


for B, h, L, Lk, dh, d, label in SHAPES:
    HD = h * dh
    q = torch.randn(B, L, HD, device=dev).bfloat16()
    k = torch.randn(B, Lk, HD, device=dev).bfloat16()
    v = torch.randn(B, Lk, HD, device=dev).bfloat16()
    do = torch.randn(B, L, HD, device=dev).bfloat16()
    o, lse = ops.attention_fwd(q, k, v, h, dh, d ** -0.5)
    tf = timeit(lambda: ops.attention_fwd(q, k, v, h, dh, d ** -0.5))
    tb = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, do, h, dh, d ** -0.5))
    fl = 4.0 * B * h * L * Lk * d                      # QK^T + PV at the true head width
    sp_bytes = B * h * L * Lk * (4 + 4 + 2 + 2)         # f32 S written+read, bf16 P written+read (forward only)
    io = (2 * q.numel() + 2 * k.numel()) * 2            # q, o, k, v
    print(f"{label:28s} B={B:3d} L={L:5d} Lk={Lk:4d} d={d:2d}(stored {dh}) | fwd {tf*1e3:8.1f} us {fl/tf/1e9:7.1f} TF/s "
          f"io {io/tf/1e6:6.0f} GB/s | bwd {tb*1e3:8.1f} us {2*fl/tb/1e9:7.1f} TF/s | avoided S/P traffic {sp_bytes/1e9:6.2f} GB "
          f"(= {sp_bytes/6.5e12*1e6:7.1f} us at 6.5 TB/s)", flush=True)
