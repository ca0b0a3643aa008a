"""Alias of flaxdiff/schedulers/karras.py's module path; the implementation lives in .ve."""
from .ve import *  # noqa: F401,F403
