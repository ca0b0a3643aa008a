"""Options outside the supported hot path must fail loudly at construction time (never silently run
something else): SURVEY.md section 8 (b) 'ownership / errors', DESIGN.md section 7."""
import pytest
import torch

from flaxdiff_b200._lib import FdxError
from flaxdiff_b200.inputs import DiffusionInputConfig
from flaxdiff_b200.models.simple_unet import Unet
from flaxdiff_b200.predictors import KarrasPredictionTransform
from flaxdiff_b200.samplers import EulerSampler
from flaxdiff_b200.schedulers import EDMNoiseScheduler, KarrasVENoiseScheduler
from flaxdiff_b200.trainer import GeneralDiffusionTrainer, adamw


@pytest.mark.parametrize("kw,needle", [
    (dict(norm_groups=0), "norm_groups"),
    (dict(attention_configs=(None, None, None, {"heads": 8, "use_projection": True})), "use_projection"),
    (dict(attention_configs=(None, None, None, {"heads": 8, "only_pure_attention": False})), "only_pure_attention"),
    (dict(attention_configs=(None, None, None)), "one entry per level"),
    (dict(output_channels=4), "output_channels"),
    (dict(dtype=torch.float16), "dtype"),
    (dict(dtype=torch.float32), "dtype"),
    (dict(activation="relu"), "activation"),
    (dict(activation=torch.nn.functional.gelu), "activation"),
])
def test_unet_rejects_unsupported_options(kw, needle):
    with pytest.raises(FdxError, match=needle):
        Unet(**kw)


def test_unet_accepts_the_reference_defaults_and_swish_callables():
    Unet()                                                   # reference default: attention at every level
    Unet(attention_configs=(None,) * 4, dtype=torch.bfloat16, activation=torch.nn.functional.silu)

    def swish(x):
        return x
    Unet(attention_configs=(None,) * 4, activation=swish)    # jax.nn.swish is a function named "swish"


@pytest.mark.parametrize("kw,needle", [(dict(autoencoder=object()), "autoencoder"),
                                       (dict(use_dynamic_scale=True), "DynamicScale")])
def test_trainer_rejects_unsupported_options(kw, needle):
    with pytest.raises(FdxError, match=needle):
        GeneralDiffusionTrainer(Unet(attention_configs=(None,) * 4), adamw(1e-3), EDMNoiseScheduler(1),
                                DiffusionInputConfig("image", (16, 16, 3), []), rngs=0, device="cpu", **kw)


def test_sampler_rejects_video_and_raw_conditioning():
    smp = EulerSampler(Unet(attention_configs=(None,) * 4), KarrasVENoiseScheduler(1), KarrasPredictionTransform(0.5),
                       DiffusionInputConfig("image", (16, 16, 3), []))
    with pytest.raises(FdxError, match="sequence_length"):
        smp.generate_samples(None, 1, 16, sequence_length=8)
    with pytest.raises(FdxError, match="conditioning"):
        smp.generate_samples(None, 1, 16, conditioning=["a photo"])
