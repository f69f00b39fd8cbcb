"""Minimal stand-in for OpenAI gym 0.17 (classic 4-tuple API) used by the reference.

TEST INFRASTRUCTURE ONLY.  gym is not installed in the build container.  Semantics kept
(SURVEY.md Appendix A): ``Env`` has class-level ``action_space = observation_space = None``;
``register``/``make``/``spec`` import ``module:Class``, instantiate with no arguments, attach
``.spec`` and wrap in the 0.17 ``TimeLimit`` (``elapsed >= max_episode_steps`` => done).
"""
import importlib

from . import spaces  # noqa: F401
from .utils import seeding  # noqa: F401


class Env:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = env.reward_range
        self.metadata = env.metadata

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def spec(self):
        return self.env.spec

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def render(self, mode="human", **kw):
        return self.env.render(mode, **kw)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)


class TimeLimit(Wrapper):
    """gym 0.17 wrappers/time_limit.py semantics."""

    def __init__(self, env, max_episode_steps=None):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return observation, reward, done, info

    def reset(self, **kw):
        self._elapsed_steps = 0
        return self.env.reset(**kw)


class EnvSpec:
    def __init__(self, id, entry_point=None, max_episode_steps=None, reward_threshold=None,
                 nondeterministic=False, kwargs=None, **_):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self.reward_threshold = reward_threshold
        self.nondeterministic = nondeterministic
        self._kwargs = kwargs or {}

    def make(self, **kwargs):
        mod_name, cls_name = self.entry_point.split(":")
        cls = getattr(importlib.import_module(mod_name), cls_name)
        kw = dict(self._kwargs)
        kw.update(kwargs)
        env = cls(**kw)
        env.spec = self
        if self.max_episode_steps is not None:
            env = TimeLimit(env, max_episode_steps=self.max_episode_steps)
        return env


_registry = {}


def register(id, **kwargs):
    _registry[id] = EnvSpec(id, **kwargs)


def spec(id):
    return _registry[id]


def make(id, **kwargs):
    return _registry[id].make(**kwargs)
