"""Shim: only what the reference touches at import time / on the forward path."""
from . import filters  # noqa: F401
