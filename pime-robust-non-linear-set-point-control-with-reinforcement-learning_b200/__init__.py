"""pime_b200 -- B200-native (sm_100a) batched plant step + P/PI prior + integrated-error observation + residual
actor forward of PIME (ruoqizzz/PIME-Robust-Non-linear-Set-point-control-with-Reinforcement-Learning).

The directory name follows the project naming rule and is not a Python identifier; import it through the
``pime_b200`` alias package at the repository root.

  pime_b200._lib   ctypes binding of the C ABI (include/pime_b200.h, libpime_b200.so)
  pime_b200.vec    WaterTankVec / PHVec / ActorPack: millions of envs, SoA torch tensors, fused rollout
  pime_b200.build  nvcc build of the library (sm_100a only)
"""
__version__ = "0.1.0"
