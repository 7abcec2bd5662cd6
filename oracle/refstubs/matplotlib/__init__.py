"""Stub: the reference imports matplotlib at module level for plots that the oracle never draws."""
