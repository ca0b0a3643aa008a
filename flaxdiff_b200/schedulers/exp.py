"""Alias of flaxdiff/schedulers/exp.py's module path; the implementation lives in .vp."""
from .vp import *  # noqa: F401,F403
