"""Stub of matplotlib (module-level import of Training_DQN_pytorch.py; never called).  TEST INFRASTRUCTURE ONLY."""
