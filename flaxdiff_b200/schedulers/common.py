"""Alias of flaxdiff/schedulers/common.py's module path; the implementation lives in .base."""
from .base import *  # noqa: F401,F403
from .base import get_coeff_shapes_tuple, reshape_rates  # noqa: F401
