"""flaxdiff.schedulers-compatible surface (flaxdiff/schedulers/__init__.py)."""
from .base import (ContinuousNoiseScheduler, GeneralizedNoiseScheduler, NoiseScheduler,
                   get_coeff_shapes_tuple, reshape_rates)
from .vp import (CosineContinuousNoiseScheduler, CosineNoiseScheduler, DiscreteNoiseScheduler,
                 ExpNoiseSchedule, LinearNoiseSchedule, SqrtContinuousNoiseScheduler,
                 cosine_beta_schedule, exp_beta_schedule, linear_beta_schedule)
from .ve import (CosineGeneralNoiseScheduler, EDMNoiseScheduler, KarrasVENoiseScheduler,
                 SimpleExpNoiseScheduler)
