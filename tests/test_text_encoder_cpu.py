"""CLIPTextEncoder (flaxdiff/inputs/encoders.py:53-94 on the torch backend): tokenizer contract, output shape /
dtype, frozen + deterministic, serialisation, and the loud failure when the weights are not in the local cache."""
import pytest
import torch

from flaxdiff_b200._lib import FdxError
from flaxdiff_b200.inputs import CLIPTextEncoder, ConditionalInputConfig, DiffusionInputConfig
from flaxdiff_b200.inputs.encoders import CONDITIONAL_ENCODERS_REGISTRY


class _Tok:
    """Stand-in for AutoTokenizer: records the arguments CLIPTextEncoder must pass (encoders.py:31-34)."""
    model_max_length = 77

    def __call__(self, data, padding=None, max_length=None, truncation=None, return_tensors=None):
        assert padding == "max_length" and max_length == 77 and truncation is True and return_tensors == "pt"
        ids = torch.zeros((len(data), 77), dtype=torch.long)
        mask = torch.zeros((len(data), 77), dtype=torch.long)
        for i, s in enumerate(data):
            toks = [1 + (hash(w) % 900) for w in str(s).split()][:75]
            ids[i, :len(toks)] = torch.tensor(toks, dtype=torch.long)
            mask[i, :len(toks) + 1] = 1
        return {"input_ids": ids, "attention_mask": mask}


def _tiny_clip():
    from transformers import CLIPTextConfig, CLIPTextModel
    torch.manual_seed(0)
    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                         num_attention_heads=4, max_position_embeddings=77)
    return CLIPTextModel(cfg)


def test_clip_text_encoder_contract():
    enc = CLIPTextEncoder(_tiny_clip(), _Tok(), modelname="tiny-random")
    out = enc(["a photo of a cat", "a painting of a dog in the style of monet", ""])
    assert out.shape == (3, 77, 64) and out.dtype == torch.bfloat16 and not out.requires_grad
    assert torch.equal(out, enc(["a photo of a cat", "a painting of a dog in the style of monet", ""]))
    assert not torch.equal(out[0], out[1])
    assert enc.key == "text" and enc.serialize() == {"modelname": "tiny-random", "backend": "torch"}
    assert CONDITIONAL_ENCODERS_REGISTRY["text"] is CLIPTextEncoder
    # an embedding tensor passes through untouched (the trainer feeds cached embeddings)
    assert enc(out) is out


def test_clip_text_encoder_drives_the_conditioning_config():
    enc = CLIPTextEncoder(_tiny_clip(), _Tok())
    cond = ConditionalInputConfig(encoder=enc, conditioning_data_key="text", pretokenized=False,
                                  unconditional_input="", model_key_override="textcontext")
    cfg = DiffusionInputConfig("image", (64, 64, 3), [cond])
    null = cfg.get_unconditionals()[0]
    assert null.shape[-2:] == (77, 64)
    out = cfg.process_conditioning({"text": ["a", "b"]})
    assert out[0].shape == (2, 77, 64)


def test_from_modelname_fails_loudly_offline():
    with pytest.raises(FdxError, match="local Hugging Face cache"):
        CLIPTextEncoder.from_modelname("openai/clip-vit-large-patch14")
    with pytest.raises(FdxError, match="only the torch backend"):
        CLIPTextEncoder.from_modelname(backend="jax")


def test_input_config_serialize_deserialize_roundtrip():
    from flaxdiff_b200.inputs import RandomEmbeddingEncoder
    enc = RandomEmbeddingEncoder(seq_len=77, features=32, seed=5)
    cfg = DiffusionInputConfig("image", (64, 64, 3), [ConditionalInputConfig(
        encoder=enc, conditioning_data_key="caption", unconditional_input="", model_key_override="textcontext")])
    ser = cfg.serialize()
    back = DiffusionInputConfig.deserialize(ser, registry={"text": RandomEmbeddingEncoder})
    assert back.sample_data_key == "image" and back.sample_data_shape == (64, 64, 3)
    c = back.conditions[0]
    assert (c.conditioning_data_key, c.unconditional_input, c.model_key_override) == ("caption", "", "textcontext")
    assert torch.equal(c.encoder(["x y"]), enc(["x y"])) and back.serialize() == ser
    with pytest.raises(ValueError, match="Unknown encoder type"):
        ConditionalInputConfig.deserialize(dict(ser["conditions"][0], encoder_key="audio"))
