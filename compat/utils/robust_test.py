"""utils/robust_test.py:4-47: the three off-nominal (a1, a2, Kp) water-tank tests, all in the same launches."""
import os

import numpy as np

from pime_b200 import scenarios as SC
from utils.test import _agent_pack, _save_run, _wt_K


def robust_test_nonlinear_watertank(env, agent, save_path, if_uniform):
    base = getattr(env, "env", env)
    cfg = base._cfg_kwargs()
    cfg.pop("max_step", None)
    kw = dict(params=SC.ROBUST_TESTS, obs_mode=base._obs_mode, num_stack=base.num_stack, max_step=500, dtype=base._dtype,
              setpoints=SC.WT_SETPOINTS, seed=base.vec.seed, **cfg)       # test_watertank -> test_policy_uniform: 2, 6, 9, 4, 1
    res, params = SC.robust_sweep(_wt_K(base), actor=_agent_pack(agent), **kw)
    lin, _ = SC.robust_sweep(_wt_K(base), policy="linear", **kw)
    for i, (a1, a2, Kp) in enumerate(params):
        save_dir = os.path.join(save_path, f"robust_test/test{i + 1}")
        os.makedirs(save_dir, exist_ok=True)
        _save_run(save_dir, {k: (v[:, i] if v is not None else None) for k, v in res.items()},
                  {k: (v[:, i] if v is not None else None) for k, v in lin.items()}, water_tank=True)
        with open(os.path.join(save_dir, "params.txt"), "w") as f:
            f.write(str({"a1": float(a1), "a2": float(a2), "Kp": float(Kp)}))
    return res, np.asarray(params)
