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
    (dict(attention_configs=(None, None, None, {"heads": 8, "flash_attention": True})), "flash_attention"),
    (dict(attention_configs=(None, None, None, {"heads": 8, "norm_inputs": False})), "norm_inputs"),
    (dict(attention_configs=(None, None, None, {"heads": 8, "only_pure_attention": False,
                                                "explicitly_add_residual": False})), "explicitly_add_residual"),
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
    # the full transformer block and the projections are supported (SURVEY $8 f3)
    m = Unet(attention_configs=(None, None, None, {"heads": 8, "only_pure_attention": False, "use_projection": True}),
             context_dim=768)
    names = [n for n, _ in m.param_specs()]
    for leaf in ("project_in/kernel", "Attention/Attention1/to_q/kernel", "Attention/Attention2/to_k/kernel",
                 "Attention/ff/net_0/proj/kernel", "Attention/ff/net_0/proj/bias", "Attention/ff/net_2/kernel",
                 "Attention/norm1/scale", "Attention/norm3/scale", "project_out/kernel"):
        assert f"middle_attention_0/{leaf}" in names, leaf
    lay = dict(m.param_specs())
    assert lay["middle_attention_0/Attention/ff/net_0/proj/kernel"] == (512, 4096)       # dim -> 2 * 4 * dim
    assert lay["middle_attention_0/Attention/ff/net_2/kernel"] == (2048, 512)
    assert lay["middle_attention_0/Attention/Attention2/to_k/kernel"] == (768, 8, 64)


@pytest.mark.parametrize("kw,needle", [(dict(autoencoder=object()), "autoencoder")])
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


def test_trainer_namespace_matches_reference_exports():
    """flaxdiff/trainer/__init__.py exports SimpleTrainer, SimpleTrainState, Metrics, DiffusionTrainer, TrainState,
    GeneralDiffusionTrainer, ConditionalInputConfig; the generic SimpleTrainer refuses loudly."""
    import pytest
    from flaxdiff_b200 import trainer as T
    from flaxdiff_b200._lib import FdxError
    for name in ("SimpleTrainer", "SimpleTrainState", "Metrics", "DiffusionTrainer", "TrainState",
                 "GeneralDiffusionTrainer", "ConditionalInputConfig"):
        assert hasattr(T, name), name
    with pytest.raises(FdxError, match="outside the supported hot path"):
        T.SimpleTrainer(None, None)
    m = T.Metrics.empty()
    for v in (1.0, 2.0, 6.0):
        m = m.merge(m.single_from_model_output(loss=v))
    assert m.compute() == {"loss": 3.0}
