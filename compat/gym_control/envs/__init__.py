"""gym_control/envs/__init__.py of the reference exports the env classes; same names, CUDA-backed."""
from gym_control.envs.nonlinear_watertank import (NonLinearWaterTankChangingParamUniformGoal,  # noqa: F401
                                                   NonLinearWaterTankChangingParamUniformGoalIntegrator,
                                                   NonLinearWaterTankChangingParamUniformGoalStacking,
                                                   StackingHistoryPreprocessing)
from gym_control.envs.ph import (PH1DChangingParamUniformGoal, PH1DChangingParamUniformGoalIntegrator,  # noqa: F401
                                 PH1DChangingParamUniformGoalIntegrator_NoBound)
