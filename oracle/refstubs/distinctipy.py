"""Stub: only imported by anno/utils.py for palette generation (not on the hot path)."""
