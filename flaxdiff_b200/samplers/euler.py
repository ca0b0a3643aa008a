"""Alias of flaxdiff/samplers/euler.py's module path; the implementation lives in .steps."""
from .steps import *  # noqa: F401,F403
