"""Checkpoint interchange with the reference's orbax layout (flaxdiff/trainer/simple_trainer.py:341-389).

The reference saves, through `orbax.checkpoint.CheckpointManager(PyTreeCheckpointer())` with
`save_args_from_target` (every leaf aggregated), the pytree
    {'rngs', 'state': {step, params: {'params': tree}, ema_params, opt_state: (ScaleByAdamState{count, mu, nu},
      EmptyState), rngs [, dynamic_scale]}, 'best_state': ..., 'best_loss', 'epoch'}
as the directory  <base>/<step>/default/{_METADATA, checkpoint}  +  <base>/<step>/_CHECKPOINT_METADATA  (the real
FlaxDiff checkpoints under /root/reference/pretrained/*/<step>/ have exactly these files; their `checkpoint`
payloads are git-LFS pointers).  `checkpoint` is ONE msgpack document in the flax.serialization wire format
(nested maps with string keys; an ndarray is ExtType(1, msgpack((shape, dtype.name, C-order bytes)));
tuples become maps keyed '0', '1', ...), `_METADATA` lists every leaf's key path
(key_type 2 = dict key, 1 = sequence index) with `skip_deserialize: true` = "the value lives in the aggregate
file".  This module writes and reads that layout, with HWIO conv kernels / (in, out) dense kernels passed
through untouched, so a tree restored by orbax from a FlaxDiff run drops into `FlatParams` and vice versa.

Not verified against orbax itself (not installable in this image): the structure is checked against the
reference's own `_METADATA` files (tests/golden/orbax_metadata_keys.json).
"""
from __future__ import annotations

import json
import os
import time
from typing import Any, Dict, Iterable, List, Tuple

import msgpack
import numpy as np

_EXT_NDARRAY = 1          # flax.serialization._MsgpackExtType.ndarray
_EXT_NPSCALAR = 3         # flax.serialization._MsgpackExtType.npscalar


# --------------------------------------------------------------------------- flax msgpack wire format
def _ndarray_to_bytes(arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)
    if arr.dtype.hasobject or arr.dtype.isalignedstruct:
        raise ValueError("object / struct arrays cannot be serialised")
    tpl = (list(arr.shape), arr.dtype.name, arr.tobytes("C"))
    return msgpack.packb(tpl, use_bin_type=True)


def _ndarray_from_bytes(data: bytes) -> np.ndarray:
    shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
    return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape).copy()


def _default(obj):
    if isinstance(obj, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, _ndarray_to_bytes(obj))
    if isinstance(obj, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(obj)))
    raise TypeError(f"cannot serialise {type(obj)}")


def _ext_hook(code, data):
    if code == _EXT_NDARRAY:
        return _ndarray_from_bytes(data)
    if code == _EXT_NPSCALAR:
        return _ndarray_from_bytes(data)[()]
    return msgpack.ExtType(code, data)


def to_state_dict(tree):
    """flax.serialization.to_state_dict for plain containers: tuples / lists -> {'0': ..., '1': ...}."""
    if isinstance(tree, dict):
        return {str(k): to_state_dict(v) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return {str(i): to_state_dict(v) for i, v in enumerate(tree)}
    if tree is None:
        return None
    if isinstance(tree, (np.ndarray, np.generic)):
        return tree
    if hasattr(tree, "detach"):                      # torch tensor
        return tree.detach().cpu().numpy()
    if isinstance(tree, (bool, int, float)):
        return np.asarray(tree)
    raise TypeError(f"unsupported leaf {type(tree)}")


def msgpack_serialize(state_dict) -> bytes:
    return msgpack.packb(state_dict, default=_default, use_bin_type=True, strict_types=True)


def msgpack_restore(data: bytes):
    return msgpack.unpackb(data, ext_hook=_ext_hook, raw=False, strict_map_key=False)


# --------------------------------------------------------------------------- _METADATA
def _walk(tree, path=(), types=()) -> Iterable[Tuple[Tuple[str, ...], Tuple[int, ...], Any]]:
    if isinstance(tree, dict):
        for k, v in tree.items():
            yield from _walk(v, path + (str(k),), types + (2,))
    elif isinstance(tree, (tuple, list)):
        for i, v in enumerate(tree):
            yield from _walk(v, path + (str(i),), types + (1,))
    else:
        yield path, types, tree


def tree_metadata(tree) -> Dict[str, Any]:
    """orbax `_METADATA` content for an all-aggregated pytree (key strings are python tuple reprs)."""
    meta = {}
    for path, types, leaf in _walk(tree):
        key = "(" + ", ".join(repr(p) for p in path) + ("," if len(path) == 1 else "") + ")"
        meta[key] = {
            "key_metadata": [{"key": p, "key_type": t} for p, t in zip(path, types)],
            "value_metadata": {"value_type": "None" if leaf is None else "jax.Array", "skip_deserialize": True},
        }
    return {"tree_metadata": meta, "use_zarr3": False}


# --------------------------------------------------------------------------- directory layout
def save_tree(base: str, step: int, tree) -> str:
    """Write <base>/<step>/{_CHECKPOINT_METADATA, default/{_METADATA, checkpoint}}; returns the step dir."""
    d = os.path.join(base, str(int(step)))
    os.makedirs(os.path.join(d, "default"), exist_ok=True)
    t0 = time.time_ns()
    with open(os.path.join(d, "default", "checkpoint"), "wb") as f:
        f.write(msgpack_serialize(to_state_dict(tree)))
    with open(os.path.join(d, "default", "_METADATA"), "w") as f:
        json.dump(tree_metadata(tree), f)
    with open(os.path.join(d, "_CHECKPOINT_METADATA"), "w") as f:
        json.dump({"init_timestamp_nsecs": t0, "commit_timestamp_nsecs": time.time_ns()}, f)
    return d


def latest_step(base: str):
    steps = [int(n) for n in os.listdir(base) if n.isdigit() and
             os.path.exists(os.path.join(base, n, "default", "checkpoint"))] if os.path.isdir(base) else []
    return max(steps) if steps else None


def load_tree(base: str, step: int = None, load_directly_from_dir: bool = False):
    """-> (step, nested dict of numpy arrays) - sequences come back as {'0': ..., '1': ...} maps, exactly what
    orbax's aggregate restore without a target returns."""
    if load_directly_from_dir:
        d = base
    else:
        if step is None:
            step = latest_step(base)
            if step is None:
                raise FileNotFoundError(f"no checkpoint under {base}")
        d = os.path.join(base, str(int(step)))
    path = os.path.join(d, "default", "checkpoint")
    with open(path, "rb") as f:
        head = f.read(64)
        if head.startswith(b"version https://git-lfs"):
            raise ValueError(f"{path} is a git-LFS pointer, not the checkpoint payload")
        data = head + f.read()
    return step, msgpack_restore(data)


def flatten_names(tree: dict, prefix: str = "") -> List[str]:
    out = []
    for k, v in tree.items():
        name = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, dict):
            out += flatten_names(v, name)
        else:
            out.append(name)
    return out
