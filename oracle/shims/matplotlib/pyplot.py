"""Stub of matplotlib.pyplot (module-level import of Training_DQN_pytorch.py; never called)."""


def __getattr__(name):
    raise RuntimeError("matplotlib stub: plotting is not available in the golden-vector harness")
