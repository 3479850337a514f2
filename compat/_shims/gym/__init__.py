"""Minimal stand-in for ``gym==0.18`` (absent from this image, no network): just what train.py, elegantrl/env.py and the env
classes touch -- Env, Wrapper, spaces.Box, the registry with make(id, **overrides), logger.set_level.  Put compat/_shims on
PYTHONPATH only when the real package is missing."""
from gym import logger, spaces  # noqa: F401
from gym.core import Env, Wrapper  # noqa: F401
from gym.envs.registration import make, register, registry, spec  # noqa: F401

__version__ = "0.18.0+pime-shim"
