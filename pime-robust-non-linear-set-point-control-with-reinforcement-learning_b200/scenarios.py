"""Batched evaluation scenarios: the reference's staircase set-point tests and robust-test parameter sweeps, every env
of a vec object at once, returning the arrays the reference's plot scripts consume (no plotting here).

  utils/test.py:70-207        test_policy_uniform            water tank, set-points 2, 6, 9, 4, 1
  utils/test.py:209-347       test_policy_uniform_integrator water tank + integrator, set-points 3, 6, 9, 4, 2
  utils/test.py:1369-1407     test_ph_policy_uniform_integrator   pH, set-points 10, 6, 3, 8, 5
  utils/test.py:1056-1067     test_watertank: agent (deterministic) vs env.get_linear_action (the CLIPPED prior)
  utils/robust_test.py:4-47   robust_test_nonlinear_watertank: off-nominal (a1, a2, Kp), max_step = 500, if_reset_all = False

One segment of the agent policy is ONE fused launch (deterministic rollout with replay rows and plant actions kept);
the linear policy is the prior kernel + the step kernel per step (its action is clipped, which the fused kernel's
prior term is not).  Between segments the reference calls reset() (new ensemble member unless if_reset_all is False,
t = 0, I = 0), restores the plant state and sets the next set-point -- reproduced in that order.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import vec as V

WT_SETPOINTS = (2.0, 6.0, 9.0, 4.0, 1.0)
WT_INTEGRATOR_SETPOINTS = (3.0, 6.0, 9.0, 4.0, 2.0)
PH_SETPOINTS = (10.0, 6.0, 3.0, 8.0, 5.0)


def _is_wt(vec):
    return isinstance(vec, V.WaterTankVec)


def _begin_segment(vec, r, resample, first, start):
    """reset(); set_state(previous plant state); set_r(r)  (utils/test.py:102-106, :1386-1388)."""
    if _is_wt(vec):
        keep = (vec.h1.clone(), vec.h2.clone())
        vec.reset(resample_params=resample)
        if first:
            vec.set_state(*start)
        else:
            vec.set_state(*keep)
        vec.set_r(r)
    else:
        keep = vec.x.clone()
        vec.reset(resample_params=resample)
        vec.x.copy_(torch.full_like(keep, float(start[0])) if first else keep)
        idx = torch.round(vec.C * vec.x * 1e5).long().clamp_(0, vec.table.numel() - 1)
        vec.y.copy_(vec.table[idx])                       # observe_state (ph.py:187-189) of the restored state
        vec.r.fill_(float(r))


def staircase(vec, policy: str, K, actor: Optional[V.ActorPack] = None, setpoints: Optional[Sequence[float]] = None,
              steps: Optional[int] = None, resample_params: bool = False, start=(0.0, 0.0)):
    """Run the staircase on every env of ``vec``.

    policy = 'agent': deterministic ``tanh(net(obs)) + obs @ (-K)`` (``actor`` = the agent's ActorPack; None = prior only);
    policy = 'linear': ``clip(-obs @ K, -1, 1)`` for the water tank (get_P_action), ``-obs @ K`` for pH (ph.py:227-231).
    Returns time-major arrays over all segments: obs [T, n, S] (the states the policy saw), xs [T, n, 2] (levels) or
    ys [T, n] (pH), refs [T, n], actions [T, n], rewards [T, n], totals [T, n] (running sum), integrators [T, n] or None.
    """
    wt = _is_wt(vec)
    if setpoints is None:
        setpoints = (WT_INTEGRATOR_SETPOINTS if vec.obs_mode == "integrator" else WT_SETPOINTS) if wt else PH_SETPOINTS
    # stacking observation: set_state / set_r do not touch the frame history (nonlinear_watertank.py:205-212 are inherited
    # unchanged), so the first num_stack observations of a segment still show the frames of the reset -- reproduced
    steps = int(steps or (vec.cfg.max_step if wt else vec.cfg.max_episode_steps))
    K = np.asarray(K, dtype=np.float64).reshape(-1)
    S = vec.state_dim
    obs_l, act_l, rew_l = [], [], []
    # episodes never terminate inside a segment (the reference ignores `done` here): lift the step limit
    limit_field = "max_step" if wt else "max_episode_steps"
    saved_limit = getattr(vec.cfg, limit_field)
    setattr(vec.cfg, limit_field, 2 ** 30)
    try:
        for k, r in enumerate(setpoints):
            _begin_segment(vec, r, resample_params, k == 0, start)
            if policy == "agent":
                out = vec.rollout(steps, -K[:S], actor=actor, deterministic=True, replay=True, want_actions=True)
                obs_l.append(out["buf_state"])
                act_l.append(out["env_action"].float())
                rew_l.append(out["buf_other"][..., 0])
            elif policy == "linear":
                o_seg = torch.empty((steps, vec.n, S), dtype=torch.float32, device=vec.device)
                a_seg = torch.empty((steps, vec.n), dtype=torch.float32, device=vec.device)
                r_seg = torch.empty((steps, vec.n), dtype=torch.float32, device=vec.device)
                obs = vec.observe()
                for t in range(steps):
                    a = vec.prior_action(obs, K[:S], clip=wt)
                    o_seg[t] = obs.t()
                    a_seg[t] = a
                    obs, rew, _ = vec.step(a)
                    r_seg[t] = rew
                obs_l.append(o_seg); act_l.append(a_seg); rew_l.append(r_seg)
            else:
                raise ValueError("policy must be 'agent' or 'linear'")
    finally:
        setattr(vec.cfg, limit_field, saved_limit)
    obs = torch.cat(obs_l)
    rewards = torch.cat(rew_l)
    res = {"obs": obs, "actions": torch.cat(act_l), "rewards": rewards, "totals": torch.cumsum(rewards.double(), 0)}
    if wt and vec.obs_mode == "stacking":   # env.state = the newest frame (update_state_P :1151-1153)
        res.update(xs=obs[..., -3:-1], refs=obs[..., -1], integrators=None)
    elif wt:
        res.update(xs=obs[..., :2], refs=obs[..., 2], integrators=obs[..., 3] if vec.obs_mode == "integrator" else None)
    else:
        res.update(ys=obs[..., 0], refs=obs[..., 1], integrators=obs[..., 2] if S == 3 else None)
    return res


ROBUST_TESTS = ((0.0024, 0.0019, 0.12), (0.0024, 0.0015, 0.12), (0.0024, 0.0015, 0.07))   # utils/robust_test.py:12-37


def robust_sweep(K, actor: Optional[V.ActorPack] = None, params=ROBUST_TESTS, obs_mode="integrator", max_step=500,
                 dtype=torch.float32, device="cuda", policy="agent", setpoints=None, **cfg):
    """robust_test_nonlinear_watertank for any list / grid of (a1, a2, Kp): one env per parameter set, all sets in the
    same launches.  Returns (staircase result dict, params array [n, 3])."""
    p = np.asarray(params, dtype=np.float64).reshape(-1, 3)
    env = V.WaterTankVec(p.shape[0], dtype=dtype, device=device, obs_mode=obs_mode, max_step=max_step, **cfg)
    env.reset()
    env.reset_changable_parameters(torch.as_tensor(p[:, 0]), torch.as_tensor(p[:, 1]), torch.as_tensor(p[:, 2]))
    res = staircase(env, policy, K, actor=actor, steps=max_step, resample_params=False, setpoints=setpoints)
    return res, p


def parameter_grid(a1, a2, Kp):
    """Cartesian product of the three ensemble-parameter axes -> [n, 3] (for robust_sweep)."""
    g = np.stack(np.meshgrid(np.asarray(a1, float), np.asarray(a2, float), np.asarray(Kp, float), indexing="ij"), -1)
    return g.reshape(-1, 3)
