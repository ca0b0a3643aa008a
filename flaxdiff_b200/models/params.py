"""Flat parameter storage with a Flax-compatible pytree view.

The reference keeps parameters as a nested dict pytree (`{'params': {...}}`) of separate
arrays (flax `model.init`, trainer/diffusion_trainer.py:107-109).  Here the same tree is a
set of *views* into one flat f32 buffer so that
  * the optimiser + EMA + bf16 shadow refresh is ONE fused kernel over the buffer,
  * the data-parallel gradient all-reduce is one (bucketed) NCCL call on a flat buffer,
  * tensor-core kernels read a flat bf16 shadow with identical offsets.
Names and layouts follow flax: conv kernels HWIO, dense kernels (in, out).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, Tuple

import torch

ALIGN = 64  # elements: 256 B in f32, 128 B in bf16 (TMA needs 16 B)


class ParamLayout:
    """Ordered table name -> (offset, shape) over a flat buffer."""

    def __init__(self, specs: Iterable[Tuple[str, Tuple[int, ...]]]):
        self.table: "OrderedDict[str, Tuple[int, Tuple[int, ...]]]" = OrderedDict()
        off = 0
        for name, shape in specs:
            n = 1
            for s in shape:
                n *= int(s)
            self.table[name] = (off, tuple(int(s) for s in shape))
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self.num_params = sum(_numel(s) for _, s in self.table.values())

    def views(self, flat: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict()
        for name, (off, shape) in self.table.items():
            out[name] = flat[off:off + _numel(shape)].view(shape)
        return out


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def nest(flat_views: Dict[str, torch.Tensor]) -> dict:
    """'a/b/kernel' -> {'a': {'b': {'kernel': t}}} (the flax param tree)."""
    root: dict = {}
    for name, t in flat_views.items():
        parts = name.split('/')
        d = root
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = t
    return root


def flatten_tree(tree: dict, prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in tree.items():
        name = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            out.update(flatten_tree(v, name))
        else:
            out[name] = v
    return out


class FlatParams(dict):
    """`{'params': tree}`-style dict whose leaves are views of `.flat` (f32)."""

    def __init__(self, layout: ParamLayout, flat: torch.Tensor):
        self.layout = layout
        self.flat = flat
        self.named = layout.views(flat)
        super().__init__({'params': nest(self.named)})
        self._shadow = None
        self._shadow_version = -1
        # Generation of `.flat`'s CONTENTS.  libfdx kernels write the buffer through raw pointers, which
        # torch's `_version` counter never sees, so every writer outside torch (fdx_adamw_ema_step, checkpoint
        # loads) calls `touch()`; `shadow()` re-casts when either counter moved.
        self.gen = 0
        self._shadow_gen = -1

    # bf16 shadow of the whole buffer (weights read by the tensor-core kernels)
    def shadow(self) -> "OrderedDict[str, torch.Tensor]":
        from .. import ops
        if self._shadow is None:
            self._shadow_flat = torch.empty(self.layout.total, dtype=torch.bfloat16, device=self.flat.device)
            self._shadow = self.layout.views(self._shadow_flat)
        if self._shadow_version != self.flat._version or self._shadow_gen != self.gen:
            ops.cast_f32_bf16(self.flat, self._shadow_flat)
            self._shadow_version = self.flat._version
            self._shadow_gen = self.gen
        return self._shadow

    def touch(self):
        """`.flat` was modified behind torch's back (a libfdx kernel or an external copy)."""
        self.gen += 1

    def shadow_is_fresh(self) -> bool:
        return (self._shadow is not None and self._shadow_version == self.flat._version
                and self._shadow_gen == self.gen)

    def shadow_flat(self) -> torch.Tensor:
        self.shadow()
        return self._shadow_flat

    def shadow_flat_noupdate(self) -> torch.Tensor:
        """The shadow buffer itself (allocated if needed) without refreshing it."""
        if self._shadow is None:
            self._shadow_flat = torch.empty(self.layout.total, dtype=torch.bfloat16, device=self.flat.device)
            self._shadow = self.layout.views(self._shadow_flat)
        return self._shadow_flat

    def mark_shadow_fresh(self):
        """The fused optimiser kernel rewrote `.flat` AND refreshed the shadow itself."""
        self.gen += 1
        self._shadow_version = self.flat._version
        self._shadow_gen = self.gen

    def clone(self) -> "FlatParams":
        return FlatParams(self.layout, self.flat.clone())

    def zeros_like(self) -> "FlatParams":
        return FlatParams(self.layout, torch.zeros_like(self.flat))


def from_tree(layout: ParamLayout, tree: dict, device) -> FlatParams:
    """Pack an arbitrary (e.g. checkpoint-loaded) flax-style tree into a FlatParams."""
    if 'params' in tree and isinstance(tree['params'], dict):
        tree = tree['params']
    named = flatten_tree(tree)
    flat = torch.zeros(layout.total, dtype=torch.float32, device=device)
    fp = FlatParams(layout, flat)
    missing = [k for k in layout.table if k not in named]
    extra = [k for k in named if k not in layout.table]
    if missing or extra:
        raise KeyError(f"param tree mismatch: missing={missing[:5]} extra={extra[:5]}")
    for k, v in named.items():
        fp.named[k].copy_(torch.as_tensor(v).to(device=device, dtype=torch.float32).reshape(fp.named[k].shape))
    return fp
