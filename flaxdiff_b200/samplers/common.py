"""Alias of flaxdiff/samplers/common.py's module path; the implementation lives in .base."""
from .base import *  # noqa: F401,F403
