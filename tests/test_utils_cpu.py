"""Host-side helpers of flaxdiff/utils.py:40-90, 239-263 (config mapping, model serialisation, latest checkpoint,
tokenizer wrapper)."""
import os

import pytest
import torch

from flaxdiff_b200 import utils


def test_map_nested_config_matches_reference_rules():
    cfg = {"dtype": "bfloat16", "precision": "high", "activation": "swish", "nothing": "None",
           "ignored.path": "jax.nn.swish", "number": 3,
           "nested": {"dtype": "float32", "activation": "gelu", "other": "keep.me.not"}}
    out = utils.map_nested_config(cfg)
    assert out == {"dtype": torch.bfloat16, "precision": None, "activation": "swish", "nothing": None,
                   "nested": {"dtype": torch.float32, "activation": "gelu"}}


def test_serialize_model_names_callables_and_dtypes():
    class M:
        def __init__(self):
            self.features = (64, 128)
            self.dtype = torch.bfloat16
            self.activation = torch.nn.functional.silu
            self.attention_configs = [None, {"heads": 8, "kernel_init": len}]
            self._private = 1
    d = utils.serialize_model(M())
    assert d == {"features": (64, 128), "dtype": "bfloat16", "activation": "silu",
                 "attention_configs": [None, {"heads": 8, "kernel_init": "len"}]}


def test_get_latest_checkpoint(tmp_path):
    for s in (10, 9, 200, 31):
        os.makedirs(tmp_path / str(s))
    os.makedirs(tmp_path / "tmp.orbax-checkpoint")       # non-numeric entries are ignored
    assert utils.get_latest_checkpoint(str(tmp_path)) == str(tmp_path / "200")
    with pytest.raises(FileNotFoundError):
        utils.get_latest_checkpoint(str(tmp_path / "10"))


def test_auto_text_tokenizer_contract():
    class Tok:
        model_max_length = 77

        def __call__(self, inputs, padding=None, max_length=None, truncation=None, return_tensors=None):
            assert (padding, max_length, truncation, return_tensors) == ("max_length", 77, True, "pt")
            n = len(inputs)
            return {"input_ids": torch.zeros(n, 77, dtype=torch.long), "attention_mask": torch.ones(n, 77, dtype=torch.long)}
    t = utils.AutoTextTokenizer(tokenizer=Tok())
    out = t(["a", "b"])
    assert out["input_ids"].shape == (2, 77) and out["caption"] == ["a", "b"] and repr(t) == "AutoTextTokenizer()"
