"""gymnasium if it is installed, otherwise a minimal in-repo stand-in with the same surface.

The reference depends on gymnasium (`gym_genesis/__init__.py:1`, `gym_genesis/env.py:1-4`) for
`gym.Env`, `spaces.Box/Dict`, `register` and `make`.  gymnasium is not installable in this image,
so the few pieces the path needs are restated here and used only on ImportError.
"""
from __future__ import annotations

try:  # pragma: no cover - exercised only where gymnasium exists
    import gymnasium as gym
    from gymnasium import spaces
    from gymnasium.envs.registration import register
    HAVE_GYMNASIUM = True
except ImportError:
    import importlib
    import types

    import numpy as np

    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape=None, dtype=None):
            self.shape, self.dtype = shape, dtype
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            super().__init__(tuple(shape), np.dtype(dtype))
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict(_Space):
        def __init__(self, spaces_dict):
            super().__init__()
            self.spaces = dict(spaces_dict)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def seed(self, seed=None):
            for i, s in enumerate(self.spaces.values()):
                s.seed(None if seed is None else seed + i)
            return [seed]

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {v!r}" for k, v in self.spaces.items()) + ")"

    class Env:
        metadata = {"render_modes": []}
        render_mode = None
        spec = None
        observation_space = None
        action_space = None
        np_random = None

        def reset(self, *, seed=None, options=None):
            if seed is not None or self.np_random is None:
                self.np_random = np.random.default_rng(seed)

        def step(self, action):
            raise NotImplementedError

        def render(self):
            return None

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def reset(self, **kw):
            return self.env.reset(**kw)

        def step(self, action):
            return self.env.step(action)

        def close(self):
            return self.env.close()

    class TimeLimit(Wrapper):
        """Same behaviour as gymnasium.wrappers.TimeLimit: after max_episode_steps the (scalar) truncated flag is set."""

        def __init__(self, env, max_episode_steps):
            super().__init__(env)
            self._max_episode_steps = max_episode_steps
            self._elapsed_steps = 0

        def reset(self, **kw):
            self._elapsed_steps = 0
            return self.env.reset(**kw)

        def step(self, action):
            obs, reward, terminated, truncated, info = self.env.step(action)
            self._elapsed_steps += 1
            if self._elapsed_steps >= self._max_episode_steps:
                truncated = True
            return obs, reward, terminated, truncated, info

    _REGISTRY = {}

    class EnvSpec:
        def __init__(self, id, entry_point, max_episode_steps=None, nondeterministic=False, kwargs=None):
            self.id, self.entry_point, self.max_episode_steps = id, entry_point, max_episode_steps
            self.nondeterministic, self.kwargs = nondeterministic, dict(kwargs or {})

    def register(id, entry_point, max_episode_steps=None, nondeterministic=False, kwargs=None, **_):
        _REGISTRY[id] = EnvSpec(id, entry_point, max_episode_steps, nondeterministic, kwargs)

    def make(id, **kwargs):
        if id not in _REGISTRY:
            raise KeyError(f"unknown environment id {id!r}; registered: {sorted(_REGISTRY)}")
        spec = _REGISTRY[id]
        ep = spec.entry_point
        if isinstance(ep, str):
            mod, attr = ep.split(":")
            ep = getattr(importlib.import_module(mod), attr)
        kw = dict(spec.kwargs)
        kw.update(kwargs)
        env = ep(**kw)
        env.spec = spec
        if spec.max_episode_steps is not None:
            env = TimeLimit(env, spec.max_episode_steps)
        return env

    spaces = types.SimpleNamespace(Box=Box, Dict=Dict, Space=_Space)
    gym = types.SimpleNamespace(Env=Env, Wrapper=Wrapper, make=make, register=register, spaces=spaces,
                                wrappers=types.SimpleNamespace(TimeLimit=TimeLimit))
