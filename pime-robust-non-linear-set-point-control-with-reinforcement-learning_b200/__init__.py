"""pime_b200 -- B200-native (sm_100a) batched plant step + P/PI prior + integrated-error observation + residual
actor forward of PIME (ruoqizzz/PIME-Robust-Non-linear-Set-point-control-with-Reinforcement-Learning).

The directory name follows the project naming rule and is not a Python identifier; import it through the
``pime_b200`` alias package at the repository root.

  pime_b200._lib       ctypes binding of the C ABI (include/pime_b200.h, libpime_b200.so)
  pime_b200.vec        WaterTankVec / PHVec / ActorPack: millions of envs, SoA torch tensors, fused rollout
  pime_b200.gym_api    the reference's seven env ids as gym-style envs (N = 1 by default, num_envs for the batched form)
  pime_b200.rl         agents / replay buffer / trainer / evaluator with the reference's names; FusedLearner
  pime_b200.scenarios  staircase set-point tests and robust-test sweeps, batched
  pime_b200.logger     the SB3-style logger module the reference's elegantrl/logger.py has to be
  pime_b200.build      nvcc build of the library (sm_100a only)
(compat/ at the repository root re-exports all of it under the reference's own module names.)
"""
__version__ = "0.2.0"
