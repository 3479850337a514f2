"""gym_control/envs/ph.py of the reference: the changing-parameter pH env classes (:350-485), CUDA-backed."""
from pime_b200.gym_api import (PH1DChangingParamUniformGoal, PH1DChangingParamUniformGoalIntegrator,  # noqa: F401
                               PH1DChangingParamUniformGoalIntegrator_NoBound)
