"""Thin Python wrappers over the libfdx C-ABI (one function per entry point).

Tensors are NHWC bf16 activations / HWIO conv kernels exactly as the reference's flax
modules store them (flaxdiff/models/common.py:166-172).  No arithmetic happens in
Python: each wrapper marshals pointers, shapes and the current CUDA stream.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import GEMM_KK, GEMM_KMN, GEMM_MNMN, act, check, load, ptr, stream_ptr


def _opt_act(t: Optional[torch.Tensor], name: str):
    return None if t is None else ctypes.byref(act(t, name))


# --------------------------------------------------------------------------- conv
def conv3x3_fwd(x: torch.Tensor, w_hwio: torch.Tensor, bias: Optional[torch.Tensor] = None,
                rowvec: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, stride: int = 1) -> torch.Tensor:
    """3x3 SAME conv, NHWC bf16, HWIO bf16 weights; out = conv + bias + rowvec[n] + res."""
    n, h, w, cin = x.shape
    cout = w_hwio.shape[-1]
    assert w_hwio.dtype == torch.bfloat16 and w_hwio.is_contiguous()
    assert tuple(w_hwio.shape) == (3, 3, cin, cout), (w_hwio.shape, cin, cout)
    ho, wo = (h + stride - 1) // stride, (w + stride - 1) // stride
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    check(load().fdx_conv3x3_fwd(ctypes.byref(act(x, "x")), ptr(w_hwio), ptr(bias), ptr(rowvec),
                                 _opt_act(res, "res"), ctypes.byref(act(out, "out")),
                                 ctypes.c_int(stride), stream_ptr()), "conv3x3_fwd")
    return out


def conv3x3_dgrad(dy: torch.Tensor, w_hwio: torch.Tensor, dx: torch.Tensor, stride: int = 1,
                  accumulate: bool = False) -> torch.Tensor:
    check(load().fdx_conv3x3_dgrad(ctypes.byref(act(dy, "dy")), ptr(w_hwio),
                                   ctypes.byref(act(dx, "dx")), ctypes.c_int(stride),
                                   ctypes.c_int(1 if accumulate else 0), stream_ptr()),
          "conv3x3_dgrad")
    return dx


def conv3x3_wgrad(x: torch.Tensor, dy: torch.Tensor, dw_hwio: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """Accumulates d(loss)/dw (f32 HWIO) into dw_hwio."""
    assert dw_hwio.dtype == torch.float32 and dw_hwio.is_contiguous()
    check(load().fdx_conv3x3_wgrad(ctypes.byref(act(x, "x")), ctypes.byref(act(dy, "dy")),
                                   ptr(dw_hwio), ctypes.c_int(stride), stream_ptr()),
          "conv3x3_wgrad")
    return dw_hwio


# --------------------------------------------------------------------------- gemm
def gemm(mode: int, A: torch.Tensor, B: torch.Tensor, D: torch.Tensor, M: int, N: int, K: int,
         a_ld: int, b_ld: int, d_ld: int, batch1: int = 1, batch2: int = 1,
         a_s=(0, 0), b_s=(0, 0), d_s=(0, 0), alpha: float = 1.0,
         bias: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
         r_ld: int = 0, r_s=(0, 0), atomic: bool = False, reduce_batch: bool = False) -> torch.Tensor:
    """Generic (batched) bf16 GEMM on the tcgen05 engine; see include/fdx.h fdx_gemm_desc."""
    g = _lib.fdx_gemm_desc()
    g.mode = mode
    g.M, g.N, g.K = M, N, K
    g.batch1, g.batch2 = batch1, batch2
    g.A, g.a_ld, g.a_s1, g.a_s2 = ptr(A), a_ld, a_s[0], a_s[1]
    g.B, g.b_ld, g.b_s1, g.b_s2 = ptr(B), b_ld, b_s[0], b_s[1]
    g.D, g.d_ld, g.d_s1, g.d_s2 = ptr(D), d_ld, d_s[0], d_s[1]
    g.d_f32 = 1 if D.dtype == torch.float32 else 0
    g.d_atomic = 1 if atomic else 0
    g.reduce_batch = 1 if reduce_batch else 0
    g.alpha = alpha
    g.bias = ptr(bias)
    g.res, g.r_ld, g.r_s1, g.r_s2 = ptr(res), r_ld, r_s[0], r_s[1]
    check(load().fdx_gemm(ctypes.byref(g), stream_ptr()), "gemm")
    return D


def linear_fwd(x2d: torch.Tensor, w_kn: torch.Tensor, out: Optional[torch.Tensor] = None,
               bias: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
               out_dtype=torch.bfloat16) -> torch.Tensor:
    """y[m][n] = x[m][k] @ w[k][n] (flax Dense kernel layout (in, out)); x row stride may exceed K."""
    M, K = x2d.shape
    Kw, N = w_kn.shape
    assert K == Kw and x2d.stride(1) == 1 and w_kn.is_contiguous()
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x2d.device)
    return gemm(GEMM_KMN, x2d, w_kn, out, M, N, K, x2d.stride(0), N, out.stride(0), bias=bias,
                res=res, r_ld=(res.stride(0) if res is not None else 0))
