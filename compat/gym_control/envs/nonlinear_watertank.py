"""gym_control/envs/nonlinear_watertank.py of the reference: the changing-parameter water-tank env classes (:828-1208) and
the frame-stacking wrapper train.py imports (:12-60)."""
from collections import deque

import numpy as np

from pime_b200.gym_api import (Box, NonLinearWaterTankChangingParamUniformGoal,  # noqa: F401
                               NonLinearWaterTankChangingParamUniformGoalIntegrator,
                               NonLinearWaterTankChangingParamUniformGoalStacking)


class StackingHistoryPreprocessing:
    """Observation = the last ``num_stack`` observations of the wrapped env, oldest first (nonlinear_watertank.py:12-60).
    (The registered Stacking1/4/10 ids do their own stacking inside the kernels; this wrapper is the generic form.)"""

    def __init__(self, env, num_stack: int):
        assert num_stack > 0
        self.env, self.num_stack = env, int(num_stack)
        self.frames = deque(maxlen=self.num_stack)
        sp = env.observation_space
        self.observation_space = Box(low=np.repeat(sp.low[np.newaxis, ...], num_stack), high=np.repeat(sp.high[np.newaxis, ...], num_stack),
                                     dtype=sp.dtype)
        self.action_space = env.action_space

    def observation(self, observation=None):
        assert len(self.frames) == self.num_stack, (len(self.frames), self.num_stack)
        return np.array(self.frames)

    def step(self, action):
        observation, reward, terminated, info = self.env.step(action)
        self.frames.append(observation)
        return self.observation(), reward, terminated, info

    def reset(self, **kwargs):
        observation = self.env.reset(**kwargs)
        for _ in range(self.num_stack):
            self.frames.append(observation)
        return self.observation()

    def __getattr__(self, name):
        if name.startswith("__") or "env" not in self.__dict__:
            raise AttributeError(name)
        return getattr(self.__dict__["env"], name)
