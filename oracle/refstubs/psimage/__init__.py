"""Stub of the (absent) psimage package so the UNMODIFIED reference can be imported for golden-vector
generation (oracle/make_golden.py). numpy-backed; `path` is looked up in a registry of arrays."""
from .core.image import PSImage  # noqa: F401
