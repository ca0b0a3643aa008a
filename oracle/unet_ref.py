"""ORACLE (test infrastructure, not product code) - CPU fp32 restatement of the reference UNet.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; flaxdiff_b200 never does.

PARITY UNPINNED: the reference ships no tests or golden vectors (SURVEY.md $4) and JAX / Flax
cannot be installed in this image, so this file restates the reference algorithm from its
sources plus the upstream Flax semantics listed in SURVEY.md Appendix B.  The only external
pin is the FourierEmbedding frequency vector, whose generator (threefry2x32 + JAX normal) is
checked against published Random123 / JAX quick-start vectors in tests/test_prng.py.

Follows, line by line:
  Unet.__call__                 flaxdiff/models/simple_unet.py:33-222
  ResidualBlock.__call__        flaxdiff/models/common.py:284-338
  Upsample / Downsample         flaxdiff/models/common.py:210-249
  FourierEmbedding / TimeProjection   flaxdiff/models/common.py:97-124
  TransformerBlock / NormalAttention  flaxdiff/models/attention.py:117-177, 321-380
Arithmetic is plain torch fp32 on the CPU with autograd for the backward pass; parameters
are the flax tree ({'a/b/kernel': tensor} flat names, HWIO conv kernels, (in,out) dense).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F


def conv_same(x, kernel_hwio, bias, stride=1):
    """flax nn.Conv, NHWC, padding='SAME' (jax: stride 1 -> (1,1); stride 2, even size -> (0,1))."""
    kh = kernel_hwio.shape[0]
    xn = x.permute(0, 3, 1, 2)
    w = kernel_hwio.permute(3, 2, 0, 1)
    if kh == 3:
        h, wd = x.shape[1], x.shape[2]

        def pads(n):
            out = -(-n // stride)
            total = max((out - 1) * stride + 3 - n, 0)
            return total // 2, total - total // 2
        pt, pb = pads(h)
        pl, pr = pads(wd)
        xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w, bias, stride=stride)
    return y.permute(0, 2, 3, 1)


def group_norm(x, scale, bias, groups, eps):
    """flax nn.GroupNorm: stats over (H, W, C/G) per sample, fast variance, fp32."""
    B, H, W, C = x.shape
    xg = x.reshape(B, H * W, groups, C // groups)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    mean2 = (xg * xg).mean(dim=(1, 3), keepdim=True)
    var = torch.clamp(mean2 - mean * mean, min=0.0)
    y = (xg - mean) * torch.rsqrt(var + eps)
    return y.reshape(B, H, W, C) * scale + bias


def rms_norm(x, scale, eps):
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * scale


def swish(x):
    return x * torch.sigmoid(x)


def gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def time_embedding(t, freqs, P):
    """FourierEmbedding (common.py:104-108) + TimeProjection (common.py:114-124)."""
    t = t.to(torch.float32)
    # jnp: (2 * jnp.pi * freqs) with f32 freqs -> weak-typed scalar times f32 array, f32 result
    w = (torch.tensor(2 * math.pi, dtype=torch.float32) * freqs.to(torch.float32))
    emb = t[:, None] * w[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    h = emb @ P["TimeProjection_0/DenseGeneral_0/kernel"] + P["TimeProjection_0/DenseGeneral_0/bias"]
    h = gelu_tanh(h)
    h = h @ P["TimeProjection_0/DenseGeneral_1/kernel"] + P["TimeProjection_0/DenseGeneral_1/bias"]
    return gelu_tanh(h)


def residual_block(P, name, x, temb, groups, n1="norm1", n2="norm2", eps=1e-4):
    residual = x
    out = group_norm(x, P[f"{name}/{n1}/scale"], P[f"{name}/{n1}/bias"], groups, eps)
    out = swish(out)
    out = conv_same(out, P[f"{name}/conv1/conv/kernel"], P[f"{name}/conv1/conv/bias"])
    tp = temb @ P[f"{name}/temb_projection/kernel"] + P[f"{name}/temb_projection/bias"]
    out = out + tp[:, None, None, :]
    out = group_norm(out, P[f"{name}/{n2}/scale"], P[f"{name}/{n2}/bias"], groups, eps)
    out = swish(out)
    out = conv_same(out, P[f"{name}/conv2/conv/kernel"], P[f"{name}/conv2/conv/bias"])
    if residual.shape != out.shape:
        residual = conv_same(residual, P[f"{name}/residual_conv/conv/kernel"], P[f"{name}/residual_conv/conv/bias"])
    return out + residual


def attention_block(P, name, x, heads, context=None, eps=1e-4):
    """TransformerBlock with only_pure_attention / norm_inputs / no projection (attention.py:321-380):
    x <- RMSNorm(x); out = x + NormalAttention(x, context or x)."""
    B, H, W, C = x.shape
    xn = rms_norm(x, P[f"{name}/RMSNorm_0/scale"], eps)
    base = f"{name}/Attention/Attention2"
    d = C // heads
    q_in = xn.reshape(B, H * W, C)
    ctx = q_in if context is None else context
    q = torch.einsum("blc,chd->blhd", q_in, P[f"{base}/to_q/kernel"])
    k = torch.einsum("bkc,chd->bkhd", ctx, P[f"{base}/to_k/kernel"])
    v = torch.einsum("bkc,chd->bkhd", ctx, P[f"{base}/to_v/kernel"])
    logits = torch.einsum("blhd,bkhd->bhlk", q / math.sqrt(d), k)
    w = torch.softmax(logits, dim=-1)
    o = torch.einsum("bhlk,bkhd->blhd", w, v)
    proj = torch.einsum("blhd,hdc->blc", o, P[f"{base}/to_out_0/kernel"]).reshape(B, H, W, C)
    return xn + proj


def _mha(P, base, xq, ctx, heads):
    """NormalAttention (attention.py:117-177) on (B, L, C) tokens."""
    d = xq.shape[-1] // heads
    q = torch.einsum("blc,chd->blhd", xq, P[f"{base}/to_q/kernel"])
    k = torch.einsum("bkc,chd->bkhd", ctx, P[f"{base}/to_k/kernel"])
    v = torch.einsum("bkc,chd->bkhd", ctx, P[f"{base}/to_v/kernel"])
    w = torch.softmax(torch.einsum("blhd,bkhd->bhlk", q / math.sqrt(d), k), dim=-1)
    o = torch.einsum("bhlk,bkhd->blhd", w, v)
    return torch.einsum("blhd,hdc->blc", o, P[f"{base}/to_out_0/kernel"])


def transformer_block(P, name, x, acfg, context=None, eps=1e-4):
    """TransformerBlock + BasicTransformerBlock + FlaxFeedForward / FlaxGEGLU in full generality
    (attention.py:179-303, 305-380): input RMSNorm, optional Dense project_in / project_out (no bias),
    self-attention (Attention1), cross-attention (Attention2; context = projected x when none is given),
    GEGLU feed-forward, residual on the NORMALISED input."""
    B, H, W, C = x.shape
    heads = acfg["heads"]
    xn = rms_norm(x, P[f"{name}/RMSNorm_0/scale"], eps)
    t = xn.reshape(B, H * W, C)
    proj = acfg.get("use_projection", False)
    px = t @ P[f"{name}/project_in/kernel"] if proj else t
    ctx = px if context is None else context
    pre = f"{name}/Attention"
    if acfg.get("only_pure_attention", True):
        h = _mha(P, f"{pre}/Attention2", px, ctx, heads)
    else:
        h = px
        if acfg.get("use_self_and_cross", True):
            n1 = rms_norm(h, P[f"{pre}/norm1/scale"], eps)
            h = h + _mha(P, f"{pre}/Attention1", n1, n1, heads)
        h = h + _mha(P, f"{pre}/Attention2", rms_norm(h, P[f"{pre}/norm2/scale"], eps), ctx, heads)
        u = rms_norm(h, P[f"{pre}/norm3/scale"], eps) @ P[f"{pre}/ff/net_0/proj/kernel"] + P[f"{pre}/ff/net_0/proj/bias"]
        lin, gate = torch.chunk(u, 2, dim=-1)
        h = h + (lin * gelu_tanh(gate)) @ P[f"{pre}/ff/net_2/kernel"] + P[f"{pre}/ff/net_2/bias"]
    if proj:
        h = h @ P[f"{name}/project_out/kernel"]
    return xn + h.reshape(B, H, W, C)


def _attn(P, name, x, acfg, context):
    if acfg.get("only_pure_attention", True) and not acfg.get("use_projection", False):
        return attention_block(P, name, x, acfg["heads"], context)
    return transformer_block(P, name, x, acfg, context)


def unet_forward(P: Dict[str, torch.Tensor], x, t, freqs, feature_depths: Sequence[int] = (64, 128, 256, 512),
                 attention_configs=(None, None, None, None), num_res_blocks=2, num_middle_res_blocks=1,
                 norm_groups=8, textcontext=None, named_norms=False):
    """x (B,H,W,3) fp32 NHWC, t (B,) -> (B,H,W,3). P: flat-name param dict."""
    n1, n2 = ("GroupNorm_0", "GroupNorm_1") if named_norms else ("norm1", "norm2")
    nout = "GroupNorm_0" if named_norms else "conv_out_norm"
    temb = time_embedding(t, freqs, P)
    x = conv_same(x, P["ConvLayer_0/conv/kernel"], P["ConvLayer_0/conv/bias"])
    downs = [x]
    L = len(feature_depths)
    for i, (dim_out, acfg) in enumerate(zip(feature_depths, attention_configs)):
        for j in range(num_res_blocks):
            x = residual_block(P, f"down_{i}_residual_{j}", x, temb, norm_groups, n1, n2)
            if acfg is not None and j == num_res_blocks - 1:
                x = _attn(P, f"down_{i}_attention_{j}", x, acfg, textcontext)
            downs.append(x)
        if i != L - 1:
            x = conv_same(x, P[f"down_{i}_downsample/ConvLayer_0/conv/kernel"],
                          P[f"down_{i}_downsample/ConvLayer_0/conv/bias"], stride=2)
    macfg = attention_configs[-1]
    for j in range(num_middle_res_blocks):
        x = residual_block(P, f"middle_res1_{j}", x, temb, norm_groups, n1, n2)
        if macfg is not None and j == num_middle_res_blocks - 1:
            x = _attn(P, f"middle_attention_{j}", x, macfg, textcontext)
        x = residual_block(P, f"middle_res2_{j}", x, temb, norm_groups, n1, n2)
    for i, (dim_out, acfg) in enumerate(zip(reversed(feature_depths), reversed(attention_configs))):
        for j in range(num_res_blocks):
            x = torch.cat([x, downs.pop()], dim=-1)
            x = residual_block(P, f"up_{i}_residual_{j}", x, temb, norm_groups, n1, n2)
            if acfg is not None and j == num_res_blocks - 1:
                x = _attn(P, f"up_{i}_attention_{j}", x, acfg, textcontext)
        if i != L - 1:
            x = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)   # jax.image.resize nearest x2
            x = conv_same(x, P[f"up_{i}_upsample/ConvLayer_0/conv/kernel"],
                          P[f"up_{i}_upsample/ConvLayer_0/conv/bias"])
    x = conv_same(x, P["ConvLayer_1/conv/kernel"], P["ConvLayer_1/conv/bias"])
    x = torch.cat([x, downs.pop()], dim=-1)
    x = residual_block(P, "final_residual", x, temb, norm_groups, n1, n2)
    x = group_norm(x, P[f"{nout}/scale"], P[f"{nout}/bias"], norm_groups, 1e-6)
    x = swish(x)
    return conv_same(x, P["ConvLayer_2/conv/kernel"], P["ConvLayer_2/conv/bias"])
