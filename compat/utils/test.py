"""utils/test.py of the reference: the evaluation renders train.py calls every ``test_render_times`` steps and at the end
(train.py:135-158,166-200).  The reference walks the staircase set-point scenarios one env step at a time through the gym
API and plots with matplotlib; here every scenario is a handful of fused launches (pime_b200.scenarios) on a private copy
of the env, the arrays the plots are drawn from are written to ``<log_path>/*.npz``, and the same ``.png`` files are drawn
when matplotlib is importable.

  test_watertank       utils/test.py:1056-1110  agent (deterministic) vs env.get_linear_action on set-points 2, 6, 9, 4, 1
                                                 (test_policy_uniform :70-207)
  test_ph_integrator   utils/test.py:1480-1573  set-points 10, 6, 3, 8, 5 (:1369-1407) + the 9 (qww_V, qc_V) robust sets
  test_ph              utils/test.py:1409-1478  the same for the pH env without integrator
  test_policy_uniform / test_ph_policy_uniform[_integrator]    the per-step loops themselves, for foreign policies

REFERENCE_SIDE_EFFECTS: the reference's pH renders work on the training env itself (no deepcopy, :1412-1413,:1482-1483) and
leave it with ``if_reset_all = False`` and the last robust (qww_V, qc_V) written by ``set_params`` (which does not refresh
the discretised system).  True (default) reproduces those side effects on the env that is passed in; False
leaves it untouched.
"""
from __future__ import annotations

import os
from copy import deepcopy

import numpy as np

from pime_b200 import scenarios as SC

try:  # optional, like every plot of this module
    import matplotlib
    matplotlib.use("Agg") if hasattr(matplotlib, "use") else None
    import matplotlib.pyplot as plt
    _HAVE_PLT = hasattr(plt, "savefig") and not getattr(plt, "_PIME_SHIM", False)
except Exception:  # noqa: BLE001
    plt, _HAVE_PLT = None, False

REFERENCE_SIDE_EFFECTS = True

# utils/test.py:1225-1240 -- the robust (qww_V, qc_V) sets of the 1-state pH plant
params_ph = {1: [[0.005, 0.0025], [0.005, 0.0015], [0.015, 0.0025], [0.015, 0.0015], [0.001, 0.002], [0.001, 0.0022],
                 [0.001, 0.0018], [0.0007, 0.002], [0.0013, 0.002]],
             2: []}


# ------------------------------------------------------------------------------------------------------ helpers
def _base(env):
    return getattr(env, "env", env) if not hasattr(env, "vec") else env


def _agent_pack(agent):
    """The agent's actor as a kernel image (None: prior-only policy)."""
    return agent._pack("act") if agent is not None and getattr(agent, "act", None) is not None else None


def _wt_K(base):
    return np.asarray(base.K, dtype=np.float64).reshape(-1)


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _plot(path, series, ylim=None):
    if not _HAVE_PLT:
        return
    plt.clf()
    for label, y in series:
        plt.plot(y, label=label)
    if ylim:
        plt.ylim(*ylim)
    plt.legend()
    plt.savefig(path, dpi=300)


def _save_run(log_path, agent_res, linear_res=None, water_tank=True):
    """One scenario -> arrays (npz) + the reference's figures."""
    os.makedirs(log_path, exist_ok=True)
    out = {}
    for tag, res in (("agent", agent_res), ("linear", linear_res)):
        if res is None:
            continue
        for k, v in res.items():
            if v is not None:
                out[f"{tag}.{k}"] = _np(v) if hasattr(v, "detach") else np.asarray(v)
    np.savez(os.path.join(log_path, "staircase.npz"), **out)
    a = {k.split(".", 1)[1]: v for k, v in out.items() if k.startswith("agent.")}
    l = {k.split(".", 1)[1]: v for k, v in out.items() if k.startswith("linear.")}
    if water_tank:
        both = lambda key, col=None: [(n, (d[key][..., col] if col is not None else d[key])) for n, d in (("agent", a), ("linear", l)) if key in d]
        _plot(os.path.join(log_path, "first_tank.png"), both("xs", 0) + ([("ref", l["refs"])] if l else []))
        _plot(os.path.join(log_path, "second_tank.png"), both("xs", 1) + ([("ref", l["refs"])] if l else []))
        _plot(os.path.join(log_path, "actions.png"), both("actions"))
        _plot(os.path.join(log_path, "total_rewards.png"), both("totals"))
    else:
        _plot(os.path.join(log_path, "state.png"), [("agent", a["ys"]), ("ref", a["refs"])])
        if a.get("integrators") is not None:
            _plot(os.path.join(log_path, "integrators.png"), [("agent", a["integrators"])])
        _plot(os.path.join(log_path, "actions.png"), [("agent", a["actions"])], ylim=(-1.1, 1.1))
        _plot(os.path.join(log_path, "total_rewards.png"), [("agent", a["totals"])])
    return out


def _squeeze(res):
    """[T, 1, ...] results of a one-env scenario -> [T, ...] like the reference's lists."""
    return {k: (v[:, 0] if v is not None else None) for k, v in res.items()}


# ------------------------------------------------------------------------------------------------------ renders
def test_watertank(env_eval_original, agent, log_path, if_uniform=False):
    """utils/test.py:1056-1110.  Returns (agent result, linear result) dicts of arrays over the 5 x max_step steps."""
    if not if_uniform:
        raise NotImplementedError("fixed-goal water-tank envs (test_policy) are not registered ids of the PIME path")
    base = _base(env_eval_original)
    env_eval, env_eval2 = deepcopy(base), deepcopy(base)          # :1058-1059
    for e in (env_eval, env_eval2):
        if hasattr(base, "integral_punish"):
            e.integral_punish = 0.0
            e.vec.cfg.integral_punish = 0.0
    K = _wt_K(base)
    res_a = _squeeze(SC.staircase(env_eval.vec, "agent", K, actor=_agent_pack(agent), setpoints=SC.WT_SETPOINTS,
                                  steps=int(env_eval.max_step), resample_params=env_eval.if_reset_all))
    res_l = _squeeze(SC.staircase(env_eval2.vec, "linear", K, setpoints=SC.WT_SETPOINTS, steps=int(env_eval2.max_step),
                                  resample_params=env_eval2.if_reset_all))
    _save_run(log_path, res_a, res_l, water_tank=True)
    return res_a, res_l


def _ph_render(env_eval_original, agent, log_path, if_uniform):
    if not if_uniform:
        raise NotImplementedError("fixed-goal pH envs are not registered ids of the PIME path")
    os.makedirs(log_path, exist_ok=True)
    base = _base(env_eval_original)
    env_eval = base if REFERENCE_SIDE_EFFECTS else deepcopy(base)   # the reference does not copy (:1412-1413, :1482-1483)
    if hasattr(base, "integral_punish"):
        env_eval.integral_punish = 0.0
        env_eval.vec.cfg.integral_punish = 0.0
    K = np.asarray(base.K, dtype=np.float64).reshape(-1)
    pack = _agent_pack(agent)
    steps = int(env_eval.max_episode_steps)
    run = lambda: _squeeze(SC.staircase(env_eval.vec, "agent", K, actor=pack, setpoints=SC.PH_SETPOINTS, steps=steps,
                                        resample_params=env_eval.if_reset_all, start=(0.0,)))
    results = [_save_run(log_path, run(), None, water_tank=False)]
    print("==================== Roobust Test ====================")
    env_eval.set_reset_all(False)                                   # :1449 / :1524
    for i, param in enumerate(params_ph[env_eval.dim]):
        robust_path = os.path.join(log_path, f"robust{i}/")
        env_eval.set_params(*param)                                 # does not refresh dsys (ph.py:263-265)
        results.append(_save_run(robust_path, run(), None, water_tank=False))
        np.savetxt(os.path.join(robust_path, "params.txt"), param)
    return results


def test_ph_integrator(env_eval_original, agent, log_path, if_uniform=False):
    """utils/test.py:1480-1573."""
    return _ph_render(env_eval_original, agent, log_path, if_uniform)


def test_ph(env_eval_original, agent, log_path, if_uniform=False):
    """utils/test.py:1409-1478."""
    return _ph_render(env_eval_original, agent, log_path, if_uniform)


# ------------------------------------------------------------------------------------------------------ per-step loops
def test_policy_uniform(env_eval, policy, if_lqr=False, setpoints=SC.WT_SETPOINTS):
    """utils/test.py:70-207 for an arbitrary ``policy(state) -> (action, ...)``, one env step at a time through the gym API
    (the renders above use the fused launches instead).  Returns xs [T,2], refs [T], actions [T], totals."""
    xs, refs, actions, totals, total = [], [], [], [], 0.0
    env_eval.reset()
    env_eval.set_state(0.0, 0.0)
    for k, r in enumerate(setpoints):
        if k:
            h1, h2 = env_eval.h1, env_eval.h2
            env_eval.reset()
            env_eval.set_state(h1, h2)
        state = env_eval.set_r(r)
        for _ in range(env_eval.max_step):
            action = policy(state)[0]
            actions.append(action)
            xs.append((env_eval.h1, env_eval.h2))
            refs.append(env_eval.r)
            state, reward, _, _ = env_eval.step(action)
            total += float(reward)
            totals.append(total)
    return np.array(xs), np.array(refs), np.array(actions), totals


def test_ph_policy_uniform_integrator(env_eval, policy, setpoints=SC.PH_SETPOINTS):
    """utils/test.py:1369-1407."""
    ys, refs, integrators, actions, totals, total = [], [], [], [], [], 0.0
    last_state = np.zeros(env_eval.dim)
    for r in setpoints:
        env_eval.reset()
        env_eval.set_state(last_state)
        state = env_eval.set_r(r)
        for _ in range(env_eval.max_episode_steps):
            action = policy(state)[0]
            actions.append(action)
            refs.append(env_eval.r)
            ys.append(env_eval.y)
            integrators.append(getattr(env_eval, "integrator", 0.0))
            state, reward, _, _ = env_eval.step(action)
            total += float(reward)
            totals.append(total)
        last_state = env_eval.state
    return np.array(ys), np.array(refs), np.array(integrators), np.array(actions), totals, []


test_ph_policy_uniform = test_ph_policy_uniform_integrator


def _not_on_path(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name}: this env family (physical tank / observer / quadcopter / reacher) is outside the PIME "
                                  "hot path (DESIGN.md section 7)")
    f.__name__ = name
    return f


for _n in ("test_realwatertank", "test_realwatertank_integrator", "test_realwatertankobserver", "test_watertankobserver",
           "test_quadcopter", "test_reacher", "test_watertanklqr"):
    globals()[_n] = _not_on_path(_n)
del _n
