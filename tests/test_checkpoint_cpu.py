"""Checkpoint interchange (SURVEY.md $8 f2): the orbax aggregate layout written / read by
flaxdiff_b200/checkpoint.py against the STRUCTURE of a real FlaxDiff checkpoint
(tests/golden/orbax_metadata_keys.json, generated from /root/reference/pretrained by
tests/golden/make_orbax_metadata_fixture.py), and the flax msgpack wire format round trip."""
import json
import os

import msgpack
import numpy as np
import pytest
import torch

from flaxdiff_b200 import checkpoint as C
from flaxdiff_b200.models.params import FlatParams, from_tree, nest
from flaxdiff_b200.models.simple_unet import Unet

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "orbax_metadata_keys.json")))


def _state_tree(model, seed=0):
    lay = model.layout()
    g = torch.Generator().manual_seed(seed)
    def rand():
        fp = FlatParams(lay, torch.randn(lay.total, generator=g))
        return {"params": nest({k: v.numpy() for k, v in fp.named.items()})}
    return {"step": np.asarray(7, np.int32), "params": rand(), "ema_params": rand(),
            "opt_state": ({"count": np.asarray(7, np.int32), "mu": rand(), "nu": rand()}, None),
            "rngs": np.asarray([1, 2], np.uint32)}


def test_metadata_structure_matches_reference_checkpoint():
    """Every non-parameter leaf of `state`, the key-type pattern of every parameter / moment leaf and the
    key-string spelling agree with the reference's own _METADATA."""
    model = Unet(attention_configs=(None,) * 4, named_norms=True)
    tree = {"state": _state_tree(model)}
    ours = C.tree_metadata(tree)
    assert sorted(ours.keys()) == GOLD["doc_keys"] and ours["use_zarr3"] == GOLD["use_zarr3"]
    om, ref = ours["tree_metadata"], GOLD["state"]

    def is_param(k):
        return "'params'" in k or "'mu'" in k or "'nu'" in k
    ref_other = {k: v for k, v in ref.items() if not is_param(k)}
    our_other = {k: v for k, v in om.items() if not is_param(k)}
    assert set(ref_other) == set(our_other), (sorted(ref_other), sorted(our_other))
    for k, v in ref_other.items():
        assert [m["key_type"] for m in our_other[k]["key_metadata"]] == v["key_types"], k
        assert our_other[k]["value_metadata"] == v["value_metadata"], k
    # parameter leaves that exist in both trees: identical key strings, key types and value metadata
    common = [k for k in ref if is_param(k) and k in om]
    assert len(common) > 600, len(common)          # 4 subtrees x the shared ResidualBlock / conv leaves
    for k in common:
        assert [m["key_type"] for m in om[k]["key_metadata"]] == ref[k]["key_types"], k
        assert [m["key"] for m in om[k]["key_metadata"]] == list(eval(k)), k
        assert om[k]["value_metadata"] == ref[k]["value_metadata"], k
    # key-type pattern per subtree over ALL reference leaves: dict keys = 2, the optax chain tuple index = 1
    for k, v in ref.items():
        parts = eval(k)
        want = [1 if (i == 2 and parts[1] == "opt_state") else 2 for i in range(len(parts))]
        assert v["key_types"] == want, k


def test_directory_layout_and_roundtrip(tmp_path):
    model = Unet(attention_configs=(None, None, None, {"heads": 8}))
    tree = {"rngs": {"rng": np.asarray([3, 4], np.uint32)}, "state": _state_tree(model, 1),
            "best_state": _state_tree(model, 2), "best_loss": np.asarray(0.125), "epoch": np.asarray(3)}
    d = C.save_tree(str(tmp_path), 1948, tree)
    files = sorted(os.listdir(os.path.join(d, "default"))) + sorted(
        f for f in os.listdir(d) if not os.path.isdir(os.path.join(d, f)))
    assert files == GOLD["sibling_files"]
    assert C.latest_step(str(tmp_path)) == 1948
    step, back = C.load_tree(str(tmp_path))
    assert step == 1948 and float(back["best_loss"]) == 0.125 and int(back["epoch"]) == 3
    assert back["state"]["opt_state"]["1"] is None                      # optax EmptyState
    k = "down_0_residual_0"
    a = tree["state"]["params"]["params"][k]["conv1"]["conv"]["kernel"]
    b = back["state"]["params"]["params"][k]["conv1"]["conv"]["kernel"]
    assert b.shape == (3, 3, 64, 64) and b.dtype == np.float32 and np.array_equal(a, b)      # HWIO pass-through
    fp = from_tree(model.layout(), back["state"]["ema_params"], "cpu")
    want = from_tree(model.layout(), tree["state"]["ema_params"], "cpu")
    assert torch.equal(fp.flat, want.flat)
    # the raw document is plain flax msgpack: ndarray = ExtType(1, packb((shape, dtype.name, bytes)))
    raw = msgpack.unpackb(open(os.path.join(d, "default", "checkpoint"), "rb").read(), raw=False,
                          strict_map_key=False)
    ext = raw["state"]["step"]
    assert isinstance(ext, msgpack.ExtType) and ext.code == 1
    shape, dtype, buf = msgpack.unpackb(ext.data, raw=False)
    assert shape == [] and dtype == "int32" and np.frombuffer(buf, np.int32)[0] == 7


def test_lfs_pointer_is_rejected(tmp_path):
    d = tmp_path / "5" / "default"
    d.mkdir(parents=True)
    (d / "checkpoint").write_bytes(b"version https://git-lfs.github.com/spec/v1\noid sha256:00\nsize 1\n")
    with pytest.raises(ValueError, match="git-LFS"):
        C.load_tree(str(tmp_path), 5)


def test_from_tree_rejects_other_architectures():
    a = Unet(attention_configs=(None,) * 4)
    b = Unet(attention_configs=(None, None, None, {"heads": 8}))
    tree = _state_tree(b)["params"]
    with pytest.raises(KeyError):
        from_tree(a.layout(), tree, "cpu")
