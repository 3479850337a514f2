"""The reference's ``gym_control`` package name: importing it registers the seven env ids (gym_control/__init__.py:3-142)
with whichever ``gym`` is importable (the real one, or compat/_shims/gym) -- backed by the CUDA envs of pime_b200.gym_api."""
from gym.envs.registration import register

from pime_b200.gym_api import REGISTRY as _REGISTRY

for _id, (_cls, _kwargs, _limit) in _REGISTRY.items():
    try:
        # the pH time limit (max_episode_steps=50, gym_control/__init__.py:6) is folded into the kernels' done flag
        # (kwargs["time_limit"]), so no TimeLimit wrapper is requested from gym
        register(id=_id, entry_point=f"gym_control.envs:{_cls.__name__}", kwargs=dict(_kwargs))
    except Exception as _e:  # noqa: BLE001  (re-import / already registered)
        if "re-register" not in str(_e).lower() and "already" not in str(_e).lower():
            raise
