"""flaxdiff.samplers-compatible surface (flaxdiff/samplers/__init__.py)."""
from .base import DiffusionSampler, linspace_int16
from .steps import (DDIMSampler, DDPMSampler, EulerAncestralSampler, EulerSampler, HeunSampler,
                    MultiStepDPM, RK4Sampler, SimpleDDPMSampler, SimplifiedEulerSampler)
