"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the *unmodified* reference (``/root/reference``) in this dev container so that golden
fixtures can be generated from the reference itself (see ``oracle/gen_golden.py``).  The
reference's third-party imports that are absent here (``gym==0.18.0``, ``control==0.9.1``,
``matplotlib``, ``serial`` and the git-ignored ``elegantrl/logger.py``) are replaced by the
minimal stub modules below; the reference's own arithmetic is executed as shipped.

Third-party arithmetic on the path and how the stubs restate it:

* ``control.tf2ss`` / ``control.c2d`` (reference ``gym_control/envs/ph.py:118-119``):
  control==0.9.1 without slycot realises the transfer function with ``scipy.signal.tf2ss``
  and discretises with ``scipy.signal.cont2discrete(method='zoh')`` -- the stub calls exactly
  those two scipy functions.
* ``gym.utils.seeding.np_random`` (``ph.py:124``): gym 0.18 hashes the seed (sha512) into a
  ``numpy.random.RandomState``; the stub reproduces ``hash_seed``/``_bigint_from_bytes``/
  ``_int_list_from_bigint`` from gym 0.18.0's published ``gym/utils/seeding.py``.
* ``gym.wrappers.TimeLimit`` (``gym_control/__init__.py:6``): done=True once the elapsed
  step count reaches ``max_episode_steps``.

This file cannot travel to the GPU box (``/root/reference`` does not exist there); only the
fixtures it produces (``tests/golden``) do.
"""
from __future__ import annotations

import hashlib
import importlib
import os
import struct
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PIME_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_control"))


# --------------------------------------------------------------------------- gym stub
def _build_gym_stub() -> types.ModuleType:
    gym = types.ModuleType("gym")

    class Env:
        metadata: dict = {}
        reward_range = (-float("inf"), float("inf"))
        spec = None
        action_space = None
        observation_space = None

        @property
        def unwrapped(self):
            return self

        def seed(self, seed=None):
            return [seed]

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env
            self.action_space = env.action_space
            self.observation_space = env.observation_space
            self.metadata = getattr(env, "metadata", {})

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

        def seed(self, seed=None):
            return self.env.seed(seed)

        def close(self):
            return self.env.close()

    class TimeLimit(Wrapper):
        """gym 0.18.0 gym/wrappers/time_limit.py semantics."""

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

        def reset(self, **kwargs):
            self._elapsed_steps = 0
            return self.env.reset(**kwargs)

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            low = np.asarray(low)
            high = np.asarray(high)
            if shape is None:
                shape = low.shape
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(low, shape).astype(self.dtype)
            self.high = np.broadcast_to(high, shape).astype(self.dtype)

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = n

    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Discrete, spaces.Space = Box, Discrete, Space

    # ---- seeding (gym 0.18.0 gym/utils/seeding.py) ----
    seeding = types.ModuleType("gym.utils.seeding")

    def create_seed(a=None, max_bytes=8):
        if a is None:
            a = int.from_bytes(os.urandom(max_bytes), "big")
        elif isinstance(a, int):
            a = a % 2 ** (8 * max_bytes)
        else:
            raise ValueError(a)
        return a

    def _bigint_from_bytes(b):
        sizeof_int = 4
        padding = sizeof_int - len(b) % sizeof_int
        b += b"\0" * padding
        int_count = int(len(b) / sizeof_int)
        unpacked = struct.unpack("{}I".format(int_count), b)
        accum = 0
        for i, val in enumerate(unpacked):
            accum += 2 ** (sizeof_int * 8 * i) * val
        return accum

    def _int_list_from_bigint(bigint):
        if bigint < 0:
            raise ValueError(bigint)
        elif bigint == 0:
            return [0]
        ints = []
        while bigint > 0:
            bigint, mod = divmod(bigint, 2 ** 32)
            ints.append(mod)
        return ints

    def hash_seed(seed=None, max_bytes=8):
        if seed is None:
            seed = create_seed(max_bytes=max_bytes)
        h = hashlib.sha512(str(seed).encode("utf8")).digest()
        return _bigint_from_bytes(h[:max_bytes])

    def np_random(seed=None):
        if seed is not None and not (isinstance(seed, int) and 0 <= seed):
            raise ValueError("Seed must be a non-negative integer or omitted, not {}".format(seed))
        seed = create_seed(seed)
        rng = np.random.RandomState()
        rng.seed(_int_list_from_bigint(hash_seed(seed)))
        return rng, seed

    seeding.np_random, seeding.hash_seed, seeding.create_seed = np_random, hash_seed, create_seed

    utils = types.ModuleType("gym.utils")
    utils.seeding = seeding
    error = types.ModuleType("gym.error")

    # ---- registration ----
    registry: dict = {}

    class EnvSpec:
        def __init__(self, id, entry_point=None, max_episode_steps=None, kwargs=None, reward_threshold=None):
            self.id = id
            self.entry_point = entry_point
            self.max_episode_steps = max_episode_steps
            self._kwargs = {} if kwargs is None else kwargs
            self.reward_threshold = reward_threshold

    def register(id, **kw):
        registry[id] = EnvSpec(id, **kw)

    def make(id, **overrides):
        spec = registry[id]
        mod_name, cls_name = spec.entry_point.split(":")
        cls = getattr(importlib.import_module(mod_name), cls_name)
        kwargs = dict(spec._kwargs)
        kwargs.update(overrides)
        env = cls(**kwargs)
        env.spec = spec
        if spec.max_episode_steps is not None:
            env = TimeLimit(env, max_episode_steps=spec.max_episode_steps)
        return env

    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register, registration.registry, registration.EnvSpec = register, registry, EnvSpec
    envs.registration = registration
    wrappers = types.ModuleType("gym.wrappers")
    wrappers.TimeLimit = TimeLimit
    logger = types.ModuleType("gym.logger")
    logger.set_level = lambda *_a, **_k: None

    gym.Env, gym.Wrapper, gym.spaces, gym.utils, gym.error = Env, Wrapper, spaces, utils, error
    gym.envs, gym.wrappers, gym.logger = envs, wrappers, logger
    gym.make, gym.register = make, register
    gym.__version__ = "0.18.0-stub"
    mods = {
        "gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding, "gym.error": error,
        "gym.envs": envs, "gym.envs.registration": registration, "gym.wrappers": wrappers, "gym.logger": logger,
    }
    return mods


def _build_control_stub():
    import scipy.signal

    control = types.ModuleType("control")

    class _SS:
        def __init__(self, A, B, C, D, dt=None):
            self.A, self.B, self.C, self.D = (np.atleast_2d(np.asarray(m, dtype=float)) for m in (A, B, C, D))
            self.dt = dt

    def tf2ss(num, den):
        A, B, C, D = scipy.signal.tf2ss(num, den)
        return _SS(A, B, C, D)

    def c2d(sys_c, Ts, method="zoh"):
        Ad, Bd, Cd, Dd, _ = scipy.signal.cont2discrete((sys_c.A, sys_c.B, sys_c.C, sys_c.D), Ts, method=method)
        return _SS(Ad, Bd, Cd, Dd, dt=Ts)

    control.tf2ss, control.c2d, control.ss = tf2ss, c2d, _SS
    matlab = types.ModuleType("control.matlab")
    control.matlab = matlab
    return {"control": control, "control.matlab": matlab}


def _build_misc_stubs():
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    mpl.use = lambda *_a, **_k: None
    serial = types.ModuleType("serial")
    return {"matplotlib": mpl, "matplotlib.pyplot": plt, "serial": serial}


def _build_logger_stub():
    logger = types.ModuleType("elegantrl.logger")
    logger.record = lambda *_a, **_k: None
    logger.dump = lambda *_a, **_k: None
    logger.configure = lambda *_a, **_k: None

    class Figure:  # noqa: D401 - SB3-style container
        def __init__(self, figure=None, close=True):
            self.figure, self.close = figure, close

    logger.Figure = Figure
    return logger


_LOADED = None


def load_reference():
    """Import the reference packages (unchanged) and return a namespace of the pieces used."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    stubs = {}
    stubs.update(_build_gym_stub())
    stubs.update(_build_control_stub())
    stubs.update(_build_misc_stubs())
    for name, mod in stubs.items():
        if name not in sys.modules:
            sys.modules[name] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import elegantrl  # the reference package (its __init__ is empty)

    logger = _build_logger_stub()
    sys.modules["elegantrl.logger"] = logger
    elegantrl.logger = logger

    ns = types.SimpleNamespace()
    ns.gym = sys.modules["gym"]
    ns.gym_control = importlib.import_module("gym_control")
    ns.wt = importlib.import_module("gym_control.envs.nonlinear_watertank")
    ns.ph = importlib.import_module("gym_control.envs.ph")
    ns.net = importlib.import_module("elegantrl.net")
    ns.net_residual = importlib.import_module("elegantrl.net_residual")
    ns.agent = importlib.import_module("elegantrl.agent")
    ns.agent_residual = importlib.import_module("elegantrl.agent_residual")
    ns.replay = importlib.import_module("elegantrl.replay")
    ns.env = importlib.import_module("elegantrl.env")
    ns.run = importlib.import_module("elegantrl.run")
    _LOADED = ns
    return ns
