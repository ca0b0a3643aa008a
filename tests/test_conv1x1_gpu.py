"""fdx_conv1x1_fwd / fdx_conv1x1_dgrad (ResidualBlock.residual_conv in convolution geometry) against fp32 torch
and against the flat fdx_gemm launch, incl. concat-slot (strided) inputs, the fused bias and accumulation."""
import pytest
import torch

from flaxdiff_b200 import ops
from flaxdiff_b200._lib import GEMM_KK, GEMM_KMN

pytestmark = pytest.mark.gpu
dev = torch.device("cuda")


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


@pytest.mark.parametrize("B,h,w,cin,cout", [(2, 32, 32, 128, 64), (3, 16, 16, 192, 128), (2, 64, 64, 320, 64), (2, 8, 8, 768, 512),
                                            (1, 16, 32, 384, 256), (5, 4, 4, 1024, 512)])
def test_conv1x1_forward_and_dgrad(B, h, w, cin, cout):
    torch.manual_seed(0)
    big = torch.randn(B, h, w, cin + 64, device=dev).bfloat16()
    x = big[..., :cin]                                   # a channel slot of a wider buffer
    wt = (torch.randn(cin, cout, device=dev) / cin ** 0.5).bfloat16()
    bias = torch.randn(cout, device=dev)
    y = ops.conv1x1_fwd(x, wt, bias)
    want = x.float() @ wt.float() + bias
    assert rel(y, want) < 5e-3
    yg = torch.empty_like(y)
    M = B * h * w
    ops.gemm(GEMM_KMN, x, wt, yg, M, cout, cin, x.stride(2), cout, cout, bias=bias)
    assert rel(y, yg) < 2e-3
    dy = torch.randn(B, h, w, cout, device=dev).bfloat16()
    base = torch.randn(B, h, w, cin, device=dev).bfloat16()
    dx = base.clone()
    ops.conv1x1_dgrad(dy, wt, dx, accumulate=True)
    assert rel(dx, dy.float() @ wt.float().t() + base.float()) < 5e-3
    dx2 = torch.empty_like(base)
    ops.conv1x1_dgrad(dy, wt, dx2)
    assert rel(dx2, dy.float() @ wt.float().t()) < 5e-3
