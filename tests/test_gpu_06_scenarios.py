"""GPU tests of the batched evaluation scenarios (pime_b200.scenarios): the reference's staircase set-point tests and
robust-test sweeps (utils/test.py:70-347,1369-1407; utils/robust_test.py) against the CPU oracle, segment by segment."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

K_WT = np.array([0.0, 0.4, -0.4, 0.0])


def _oracle_staircase(O, params, setpoints, steps, policy):
    n = params.shape[0]
    cfg = O.wt_cfg(reward_type="square_distance")
    cfg.max_step = 2 ** 30
    h1, h2 = np.zeros(n), np.zeros(n)
    a1, a2, Kp = params[:, 0].copy(), params[:, 1].copy(), params[:, 2].copy()
    xs, acts, rews = [], [], []
    for r_ in setpoints:
        r, I, t = np.full(n, float(r_)), np.zeros(n), np.zeros(n, np.int32)
        if policy == "agent":
            o = O.wt_rollout(cfg, None, None, -0.5, -K_WT, 1, 0, True, steps, h1, h2, r, I, t, a1, a2, Kp, want_actions=True)
            xs.append(o["buf_state"]); acts.append(o["env_action"]); rews.append(o["buf_other"][..., 0])
        else:
            for _ in range(steps):
                obs = np.stack([h1, h2, r, I], 1)
                a = np.clip(-(obs.astype(np.float64) @ K_WT), -1.0, 1.0)
                xs.append(obs[None].astype(np.float32)); acts.append(a[None].copy())
                rews.append(O.wt_step(cfg, h1, h2, r, I, t, a1, a2, Kp, a)[0][None].copy())
    return np.concatenate(xs), np.concatenate(acts), np.concatenate(rews)


@pytest.mark.parametrize("policy", ["agent", "linear"])
def test_robust_sweep_matches_oracle(oracle, policy):
    import pime_b200.scenarios as SC
    steps = 60
    res, params = SC.robust_sweep(K_WT, actor=None, max_step=steps, dtype=torch.float64, policy=policy, noise_scale=0.0)
    assert params.shape == (3, 3) and np.allclose(params[0], [0.0024, 0.0019, 0.12])
    obs, acts, rews = _oracle_staircase(oracle, params, SC.WT_INTEGRATOR_SETPOINTS, steps, policy)
    T = steps * 5
    assert res["obs"].shape == (T, 3, 4) and res["xs"].shape == (T, 3, 2) and res["totals"].shape == (T, 3)
    # 'agent' evaluates obs32 @ priorK in float32 like the reference's act(states) (agent.py:584): the summation order of a
    # float32 dot product is not pinned (FMA chain here, separate products in the oracle) -> 1e-4; 'linear' is fp64 throughout
    tol = 1e-4 if policy == "agent" else 1e-6
    np.testing.assert_allclose(res["obs"].cpu().numpy(), obs, rtol=tol, atol=tol)
    np.testing.assert_allclose(res["actions"].cpu().numpy(), acts, rtol=tol, atol=tol if policy == "agent" else 1e-9)
    np.testing.assert_allclose(res["rewards"].cpu().numpy(), rews, rtol=10 * tol, atol=10 * tol)
    np.testing.assert_allclose(res["totals"].cpu().numpy()[-1], rews.astype(np.float64).sum(0), rtol=10 * tol)
    refs = res["refs"].cpu().numpy()
    assert np.all(refs[:steps] == 3.0) and np.all(refs[-steps:] == 2.0)           # 3, 6, 9, 4, 2 (utils/test.py:220-319)
    assert np.all(res["integrators"].cpu().numpy()[steps] == 0.0)                 # reset() clears the integrator


def test_grid_sweep_with_an_agent_and_ph_staircase():
    """A 4x4x4 parameter grid with a residual actor in one batch; the pH staircase (10, 6, 3, 8, 5) on 64 ensemble members."""
    import pime_b200.rl as R
    import pime_b200.scenarios as SC
    import pime_b200.vec as V
    torch.manual_seed(0)
    agent = R.AgentResidualIntegratorModularPPO()
    agent.init(64, 4, 1, 1)
    agent.init_residual({"init_K": K_WT.reshape(-1, 1)})
    grid = SC.parameter_grid(np.linspace(0.0015, 0.0024, 4), np.linspace(0.0015, 0.0024, 4), np.linspace(0.07, 0.17, 4))
    res, p = SC.robust_sweep(K_WT, actor=agent._pack("act"), params=grid, max_step=200)
    lin, _ = SC.robust_sweep(K_WT, params=grid, max_step=200, policy="linear")
    assert p.shape == (64, 3) and res["xs"].shape == (1000, 64, 2)
    # zero-initialised residual: the agent IS the (unclipped) prior policy; the linear policy clips (utils/test.py:1066-1067)
    pri, _ = SC.robust_sweep(K_WT, actor=None, params=grid, max_step=200)
    assert torch.equal(res["actions"], pri["actions"]) and torch.equal(res["obs"], pri["obs"])
    a0, b0 = res["actions"][0].cpu().numpy(), lin["actions"][0].cpu().numpy()
    assert np.allclose(a0, 1.2, atol=1e-6) and np.all(b0 == 1.0)                  # 0.4 * (3 - 0) = 1.2 -> clipped to 1
    assert float(res["totals"][-1].mean()) < 0 and float(lin["totals"][-1].mean()) < 0
    # pH
    env = V.PHVec(64, dtype=torch.float32, seed=3)
    K_PH = np.array([-0.02, 0.02, 0.035])
    r = SC.staircase(env, "agent", K_PH, actor=None)
    assert r["ys"].shape == (250, 64) and torch.isfinite(r["ys"]).all()
    refs = r["refs"].cpu().numpy()
    assert np.all(refs[:50] == 10.0) and np.all(refs[200:] == 5.0)
    err_end = (r["ys"][49] - 10.0).abs().median()
    assert float(err_end) < 2.0                                                    # the PI prior approaches the set-point


def test_stacking_staircase_equals_the_per_step_gym_loop():
    """The staircase on the Stacking observation (the env of run_watertank_changing.sh): the fused launches reproduce what the
    reference's per-step protocol (utils/test.py:70-207: reset, set_state, set_r, step ...) gives through the gym API,
    including its quirk that set_state / set_r leave the frame history alone (the first frames of a segment are stale)."""
    import copy
    import pime_b200.gym_api as G
    import pime_b200.scenarios as SC
    env = G.make("NonLinearWaterTankChangingParamUniformGoalStacking4-SquareDistance-v2", noise_scale=0.0, seed=5)
    env.reset()
    env.set_reset_all(False)                       # keep the ensemble member, as utils/robust_test.py does
    K = np.asarray(env.K, np.float64)
    steps = 40
    fused = SC.staircase(copy.deepcopy(env).vec, "agent", K, actor=None, setpoints=SC.WT_SETPOINTS, steps=steps, resample_params=False)
    e2 = copy.deepcopy(env)
    obs_l, act_l = [], []
    e2.reset()
    e2.set_state(0.0, 0.0)
    for k, r in enumerate(SC.WT_SETPOINTS):
        if k:
            h1, h2 = e2.h1, e2.h2
            e2.reset()
            e2.set_state(h1, h2)
        state = e2.set_r(r)
        for _ in range(steps):
            a = float(np.float32(state).astype(np.float64) @ (-K))      # the zero-residual agent: obs32 . priorK, unclipped
            obs_l.append(state.copy()); act_l.append(a)
            state, _, _, _ = e2.step(np.array([a]))
    want_obs = np.asarray(obs_l, np.float32)
    got_obs = fused["obs"][:, 0].cpu().numpy()
    # reset() draws (h, r) from Philox keyed by the episode counter: both paths performed the same number of resets
    np.testing.assert_allclose(got_obs, want_obs, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(fused["actions"][:, 0].cpu().numpy(), np.asarray(act_l), rtol=1e-6, atol=1e-6)
    assert fused["xs"].shape == (5 * steps, 1, 2) and torch.equal(fused["refs"][:, 0], fused["obs"][:, 0, -1])
