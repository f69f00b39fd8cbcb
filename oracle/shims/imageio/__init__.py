"""Stub of imageio for importing the reference's training / player scripts headless (TEST INFRASTRUCTURE ONLY).
Training_DQN_pytorch.py imports it at module level for its mp4 checkpoints; nothing here is ever called by the
golden-vector generator."""


def get_writer(*a, **k):
    raise RuntimeError("imageio stub: video output is not available in the golden-vector harness")
