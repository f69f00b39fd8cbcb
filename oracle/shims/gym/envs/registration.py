"""gym.envs.registration stand-in (TEST INFRASTRUCTURE ONLY)."""
from .. import register, make, spec, EnvSpec  # noqa: F401
