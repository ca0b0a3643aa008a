"""Alias of flaxdiff/schedulers/cosine.py's module path; the implementation lives in .vp."""
from .vp import *  # noqa: F401,F403
from .ve import CosineGeneralNoiseScheduler  # noqa: F401
