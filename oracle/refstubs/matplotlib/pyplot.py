def __getattr__(name):
    raise RuntimeError("matplotlib stub: plotting is not part of the oracle")
