"""GPU tests of the gym-style drop-in surface (pime_b200.gym_api): the reference's env ids, gym API, ensemble and test
API driven by the CUDA kernels, pinned on the KATs of SURVEY.md section 8c (values printed by the reference), plus
size-independent properties at the BASELINE size (2^20 envs)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WT_INT = "NonLinearWaterTankChangingParamUniformGoalIntegrator-SquareDistance-v2"
PH_INT = "PH1DChangingParamUniformGoalIntegrator-SqaureDistance-v35"


@pytest.fixture(scope="module")
def G():
    import pime_b200.gym_api as G
    return G


def test_all_registered_ids_construct_with_reference_shapes(G):
    want = {"PH1DChangingParamUniformGoalIntegrator-SqaureDistance-v35": (3, 50),
            "PH1DChangingParamUniformGoalIntegrator-SqaureDistance-NoIB-v35": (3, 50),
            "NonLinearWaterTankChangingParamUniformGoal-SquareDistance-v2": (3, 200),
            "NonLinearWaterTankChangingParamUniformGoalIntegrator-SquareDistance-v2": (4, 200),
            "NonLinearWaterTankChangingParamUniformGoalStacking4-SquareDistance-v2": (12, 200),
            "NonLinearWaterTankChangingParamUniformGoalStacking10-SquareDistance-v2": (30, 200),
            "NonLinearWaterTankChangingParamUniformGoalStacking1-SquareDistance-v2": (3, 200)}
    assert set(G.REGISTRY) == set(want)                                  # gym_control/__init__.py:3-142
    import pime_b200.rl as R
    for env_id, (S, T) in want.items():
        env = G.make(env_id)
        assert env.unwrapped.spec.id == env_id
        obs = env.reset()
        assert obs.shape == (S,) and obs.dtype == np.float64 and env.observation_space.shape == (S,)
        assert env.action_space.shape == (1,) and float(env.action_space.high[0]) == 1.0
        o2, rew, done, info = env.step(np.array([0.1]))
        assert o2.shape == (S,) and isinstance(rew, float) and isinstance(done, bool) and info == {}
        # the reference registers a 4-entry K for the 3-observation id (gym_control/__init__.py:46) -- kept as is
        assert env.K.shape == ((4,) if env_id == "NonLinearWaterTankChangingParamUniformGoal-SquareDistance-v2" else (S,))
        assert env.seed(5) == [5]
        assert R.PreprocessEnv(env).max_step == T                        # elegantrl/env.py:221-231


def test_kat1_water_tank_through_the_gym_api(G):
    """bit-identical to the reference's printed trajectory."""
    env = G.make(WT_INT, reward_type="distance", noise_scale=0.0)
    env.reset()
    env.reset_changable_parameters(0.0019, 0.0019, 0.12)
    env.set_state(0.0, 0.0)
    env.set_r(3.0)
    env.integrator = 0.0
    obs = env._get_observe()
    rows = [(1.2000000000000002, 2.4692863612658598, 0.1371267442911145, 2.8628732557088856, -2.8628732557088856),
            (1.1451493022835542, 4.727828693028896, 0.3706586062224364, 5.492214649486449, -2.6293413937775636),
            (1.0517365575110256, 6.788364667539474, 0.6534380242159421, 7.838776625270507, -2.346561975784058),
            (0.9386247903136232, 8.648875526100547, 0.9689474740993569, 9.86982915117115, -2.031052525900643)]
    priorK = -env.K.reshape(-1, 1)
    for action, h1, h2, I, rew in rows:
        a = obs @ priorK
        assert float(a[0]) == action
        obs, r, done, _ = env.step(a)
        assert (obs[0], obs[1], obs[3], r) == (h1, h2, I, rew) and not done
    assert env.get_changable_parameters() == (0.0019, 0.0019, 0.12)
    assert (env.h1, env.h2, env.r, env.integrator) == (rows[-1][1], rows[-1][2], 3.0, rows[-1][3])


def test_kat4_ph_through_the_gym_api(G):
    """SURVEY 8c KAT-4 (x' = A x + B u bit-identical; y through the device-built table: 1e-12)."""
    env = G.make(PH_INT)
    env.reset()
    env.set_params(0.01, 0.002)
    env.update_system()                       # the reference's set_params does NOT refresh dsys
    env.set_state(0.0)
    env.set_r(7.0)
    env.integrator = 0.0
    obs = env._get_observe()
    rows = [(0.09404057651683974, 14.873693355550358, 2.0108812493180586, 4.989118750681941, -24.89130590840613),
            (-0.2744015312875068, 22.042201761386956, 1.6183247937618077, 10.370793956920133, -28.96242802543889),
            (-0.47061129261696855, 25.243770371322867, 1.5158308962973068, 15.854963060622826, -30.0761107580072),
            (-0.6646070891958528, 25.227582653974668, 1.5162583986482367, 21.33870466197459, -30.071421950396),
            (-0.856529495196146, 22.605107021724443, 1.5984100713155363, 25.0, -29.177173757665432),
            (-0.9830317985736894, 18.73818227736746, 1.7574230444634973, 25.0, -27.484613134722387)]
    priorK = -env.K.reshape(-1, 1)
    for action, x, y, I, rew in rows:
        a = obs @ priorK
        assert abs(float(a[0]) - action) <= 1e-11
        obs, r, done, _ = env.step(a)
        assert abs(env.state - x) <= 1e-9 * x
        np.testing.assert_allclose([obs[0], obs[2], r], [y, I, rew], rtol=1e-10, atol=1e-10)


def test_kat5_stacking_frames(G):
    env = G.make("NonLinearWaterTankChangingParamUniformGoalStacking4-SquareDistance-v2", noise_scale=0.0)
    obs = env.reset()
    assert np.array_equal(obs, np.tile(obs[:3], 4)) and env.m == 12
    assert np.array_equal(env.K, np.array([0.0] * 9 + [0.0, 0.4, -0.4]))
    o2, _, _, _ = env.step(np.array([0.3]))
    assert np.array_equal(o2[:9], obs[:9]) and not np.array_equal(o2[9:11], obs[9:11]) and o2[11] == obs[11]


def test_reference_error_behaviour(G):
    env = G.make(WT_INT)
    with pytest.raises(AssertionError):       # "Please reset the env first" (nonlinear_watertank.py:794)
        env.step(np.array([0.0]))
    ph = G.make(PH_INT)
    ph.reset()
    with pytest.raises(IndexError):           # table overflow (ph.py:188)
        ph.set_state(1e9)
    with pytest.raises(NotImplementedError):
        G.make(WT_INT, controller_type="LQR")


def test_deepcopy_gives_an_independent_env(G):
    """utils/test.py:1058-1059 and utils/robust_test.py:5 deep-copy the env."""
    env = G.make(WT_INT, noise_scale=0.0)
    env.reset()
    twin = copy.deepcopy(env)
    a = np.array([0.25])
    o1, r1, _, _ = env.step(a)
    o2, r2, _, _ = twin.step(a)
    assert np.array_equal(o1, o2) and r1 == r2
    env.step(a)
    assert twin._episode_steps == 1 and env._episode_steps == 2


def test_ensemble_resampling_and_reset_r(G):
    env = G.make(WT_INT)
    env.seed(11)
    env.reset()
    p0 = env.get_changable_parameters()
    env.reset()
    p1 = env.get_changable_parameters()
    assert p0 != p1 and all(lo <= v <= hi for v, (lo, hi) in zip(p1, ([0.0015, 0.0024], [0.0015, 0.0024], [0.07, 0.17])))
    env.set_reset_all(False)                  # README.md:27-33: keep the ensemble member, new state / set-point only
    env.reset()
    assert env.get_changable_parameters() == p1 and env.integrator == 0.0 and env._episode_steps == 0


def test_full_size_properties_at_2_pow_20_envs():
    """Size-independent properties at the BASELINE size: (a) zero-initialised actor == prior-only policy, bit for bit;
    (b) any split of the env-id range reproduces the unsplit run (Philox keyed by global env id); (c) done exactly at
    t = 200 and in-kernel auto-reset restarts every env."""
    import pime_b200.vec as V
    n, T = 1 << 20, 200
    K = np.array([0.0, 0.4, -0.4, 0.0])
    H = 256
    sd = {}
    rng = np.random.default_rng(0)
    for name, o, i in [("other_net.0", H, 3), ("other_net.2", H // 2, H), ("integrator_net.0", H, 1),
                       ("integrator_net.2", H // 2, H), ("net.0", H, H), ("net.2", 1, H)]:
        b = 1.0 / np.sqrt(i)
        sd[name + ".weight"] = rng.uniform(-b, b, (o, i)).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-b, b, o).astype(np.float32)
    sd["net.2.weight"][:] = 0.0
    sd["net.2.bias"][:] = 0.0                                              # init_actor_zero (agent_residual.py:45-50)
    actor = V.ActorPack("modular", 4, H, 1).update(sd)

    def run(n_, off, use_actor):
        env = V.WaterTankVec(n_, dtype=torch.float32, seed=7, env_offset=off, noise_scale=0.01)
        env.reset()
        stats = torch.zeros(8, dtype=torch.float64, device="cuda")
        env.rollout(T + 3, -K, actor=actor if use_actor else None, deterministic=True, auto_reset=True, stats=stats)
        return env, stats.cpu().numpy()

    whole, st = run(n, 0, True)
    prior, st_p = run(n, 0, False)
    assert torch.equal(whole.h2, prior.h2) and torch.equal(whole.ep_return, prior.ep_return)
    np.testing.assert_allclose(st, st_p, rtol=1e-12)                        # per-env data bit-equal; the sums differ by atomicAdd order
    assert st[2] == n and st[5] == n * (T + 3) and int(whole.t.min()) == 3 == int(whole.t.max())   # one episode each, then reset
    assert int(whole.episode.min()) == 2                                                            # reset() + in-kernel reset
    lo, _ = run(n // 2, 0, True)
    hi, _ = run(n // 2, n // 2, True)
    assert torch.equal(torch.cat([lo.h2, hi.h2]), whole.h2) and torch.equal(torch.cat([lo.ep_return, hi.ep_return]), whole.ep_return)
    mean_ret = st[0] / st[2]
    assert -4000 < mean_ret < -500                                          # the P prior on random set-points (SquareDistance)


def test_full_size_properties_at_2_pow_23_ph_envs():
    """BASELINE configs[3] at its full size (2^23 pH envs, T = 50, Modular-128): zero-initialised actor == prior-only policy
    bit for bit (x is fp64 in both), any split of the ensemble-member range reproduces the unsplit sweep, every env is done
    exactly at the TimeLimit and restarted by the in-kernel reset, no table fault, pH inside the table's range."""
    import pime_b200.vec as V
    n, T = 1 << 23, 50
    K = np.array([-0.02, 0.02, 0.035])                                     # PH1D...Integrator K (ph.py), priorK = -K
    H = 128
    sd = {}
    rng = np.random.default_rng(1)
    for name, o, i in [("other_net.0", H, 2), ("other_net.2", H // 2, H), ("integrator_net.0", H, 1),
                       ("integrator_net.2", H // 2, H), ("net.0", H, H), ("net.2", 1, H)]:
        b = 1.0 / np.sqrt(i)
        sd[name + ".weight"] = rng.uniform(-b, b, (o, i)).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-b, b, o).astype(np.float32)
    sd["net.2.weight"][:] = 0.0
    sd["net.2.bias"][:] = 0.0
    actor = V.ActorPack("modular", 3, H, 1).update(sd)

    def run(n_, off, use_actor):
        env = V.PHVec(n_, dtype=torch.float32, seed=3, env_offset=off)
        env.reset()
        stats = torch.zeros(8, dtype=torch.float64, device="cuda")
        env.rollout(T + 2, -K, actor=actor if use_actor else None, deterministic=True, auto_reset=True, stats=stats)
        env.check_status()
        return env, stats.cpu().numpy()

    whole, st = run(n, 0, True)
    prior, st_p = run(n, 0, False)
    assert torch.equal(whole.x, prior.x) and torch.equal(whole.y, prior.y) and torch.equal(whole.ep_return, prior.ep_return)
    np.testing.assert_allclose(st[:6], st_p[:6], rtol=1e-12)
    assert st[2] == n and st[5] == n * (T + 2) and int(whole.t.min()) == 2 == int(whole.t.max())
    assert int(whole.episode.min()) == 2
    assert 0.0 < float(whole.y.min()) and float(whole.y.max()) < 11.8         # pH[0] = 11.70 is the table's largest entry (KAT-3)
    keep = {k: getattr(whole, k).clone() for k in ("x", "y", "I", "ep_return", "qww_V")}
    del whole, prior
    torch.cuda.empty_cache()
    lo, _ = run(n // 2, 0, True)
    hi, _ = run(n // 2, n // 2, True)
    for k, v in keep.items():
        assert torch.equal(torch.cat([getattr(lo, k), getattr(hi, k)]), v), k


def test_reset_from_last_state_follows_the_reference_script(G, golden):
    """reset_from_last_state=True (nonlinear_watertank.py:904-910, :819-821; ph.py:417-420, :345-346) against the
    scripted run of the reference in tests/golden/last_state.npz (oracle/gen_golden.py:gen_last_state): the ensemble
    member and set-point drawn by each reset are replayed through the public setters, the levels are the kernels'."""
    d = golden("last_state")
    env = G.make(WT_INT, reset_from_last_state=True, max_step=6, noise_scale=0.0, reward_type="distance")
    assert env.last_h1 is None and env.last_h2 is None
    first = True
    for (kind, arg), row in zip(d["wt_script"], d["wt_rows"]):
        h1, h2, r, I, t, l1, l2, a1, a2, Kp = row
        if kind == 0:
            env.reset()
            if first:                                   # no finished episode yet: |N(0,1)| * 0.1 (:905-907)
                assert 0.0 <= env.h1 < 1.0 and 0.0 <= env.h2 < 1.0 and env.last_h1 is None
                env.set_state(h1, h2)
                first = False
            env.reset_changable_parameters(a1, a2, Kp)
            env.set_r(r)
        else:
            _, _, done, _ = env.step(np.array([arg]))
            assert done == (t == 6)
        assert (env.h1, env.h2, env.integrator, env._episode_steps) == (h1, h2, I, int(t))
        assert (env.last_h1, env.last_h2) == ((None, None) if np.isnan(l1) else (l1, l2))

    ph = G.make(PH_INT, reset_from_last_state=True)
    assert ph.last_state is None
    for (kind, arg), row in zip(d["ph_script"], d["ph_rows"]):
        x, y, r, I, t, last, qww, qc, done_ref = row
        if kind == 0:
            ph.reset()
            if np.isnan(last):
                assert 0.0 <= ph.state < 50.0          # ph.py:420
                ph.set_state(x)
            else:
                assert ph.state == ph.last_state        # kept across the reset, whatever the new ensemble member is
                np.testing.assert_allclose(ph.state, last, rtol=1e-10)   # (B differs from scipy's c2d in the 16th digit)
            ph.set_params(qww, qc)
            ph.update_system()
            ph.set_state(ph.state)                      # y = observe_state(x) with the replayed C (:422)
            ph.set_r(r)
        else:
            _, _, done, _ = ph.step(np.array([arg]))
            assert done == bool(done_ref)
        np.testing.assert_allclose([ph.state, ph.y, ph.integrator], [x, y, I], rtol=1e-10, atol=1e-10)
        assert ph._episode_steps == int(t)
        if np.isnan(last):
            assert ph.last_state is None
        else:
            np.testing.assert_allclose(ph.last_state, last, rtol=1e-10)


def test_reset_from_last_state_in_the_fused_rollout(oracle):
    """The in-kernel auto-reset keeps the levels (every in-kernel reset follows a done) and records them; the first
    reset of a fresh env draws |N(0,1)| * 0.1 from the Box-Muller pair of the (h1, h2) uniforms."""
    import pime_b200.vec as V
    n, T = 4096, 7
    K = np.array([0.0, 0.4, -0.4, 0.0])
    env = V.WaterTankVec(n, dtype=torch.float64, seed=3, reset_from_last_state=True, max_step=T, noise_scale=0.0)
    env.reset()
    u = np.stack([oracle.reset_uniforms(3, i, 0) for i in range(64)])
    rad = np.sqrt(-2.0 * np.log1p(-u[:, 3]))
    np.testing.assert_allclose(env.h1[:64].cpu().numpy(), np.abs(rad * np.cos(2 * np.pi * u[:, 4])) * 0.1, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(env.h2[:64].cpu().numpy(), np.abs(rad * np.sin(2 * np.pi * u[:, 4])) * 0.1, rtol=1e-12, atol=1e-15)
    assert torch.isnan(env.last_h1).all()
    twin = V.WaterTankVec(n, dtype=torch.float64, seed=3, reset_from_last_state=True, max_step=T, noise_scale=0.0)
    twin.reset()
    env.rollout(T, -K, deterministic=True, auto_reset=True)            # ends exactly on the done step, then resets
    twin.rollout(T, -K, deterministic=True, auto_reset=False)
    assert torch.equal(env.h1, twin.h1) and torch.equal(env.h2, twin.h2)            # levels survive the reset ...
    assert torch.equal(env.last_h1, twin.h1) and torch.equal(env.last_h2, twin.h2)  # ... and are what `done` recorded
    assert int(env.t.max()) == 0 and int(twin.t.min()) == T and not torch.equal(env.r, twin.r)
    ph = V.PHVec(n, dtype=torch.float64, seed=5, reset_from_last_state=True, max_episode_steps=T)
    ph.reset()
    ph2 = V.PHVec(n, dtype=torch.float64, seed=5, reset_from_last_state=True, max_episode_steps=T)
    ph2.reset()
    Kp = np.array([-0.02, 0.02, 0.035])
    ph.rollout(T, -Kp, deterministic=True, auto_reset=True)
    ph2.rollout(T, -Kp, deterministic=True, auto_reset=False)
    assert torch.equal(ph.x, ph2.x) and torch.equal(ph.last_x, ph2.x) and not torch.equal(ph.C, ph2.C)


def test_attribute_mirrors_are_assignable_like_the_reference(G):
    """utils/robust_test.py:12-19 writes test_env.a1 / a2 / Kp / max_step / if_reset_all on a deep copy of the env."""
    env = G.make(WT_INT, noise_scale=0.0)
    env.reset()
    test_env = copy.deepcopy(env)
    test_env.a1, test_env.a2, test_env.Kp = 0.0024, 0.0019, 0.12
    test_env.if_reset_all = False
    test_env.max_step = 500
    assert (test_env.a1, test_env.a2, test_env.Kp) == (0.0024, 0.0019, 0.12) and test_env.max_step == 500
    assert env.max_step == 200 and env.a1 != 0.0024                      # the original is untouched
    test_env.reset()                                                     # reset_r(): the parameters stay (:935-939)
    assert test_env.get_changable_parameters() == (0.0024, 0.0019, 0.12)
    done = False
    for _ in range(500):
        _, _, done, _ = test_env.step(np.array([0.0]))
    assert done and test_env._episode_steps == 500                       # the new step limit reached the kernels
    test_env.h1 = 3.5
    assert test_env.h1 == 3.5 and float(test_env.vec.h1[0]) == 3.5


def test_stacking_set_state_leaves_the_frame_history_like_the_reference(G):
    """nonlinear_watertank.py:205-212 are inherited unchanged by the stacking env: set_state / set_r change h1, h2, r but the
    observation (the deque of frames, :1164-1166) only changes in reset() and step()."""
    env = G.make("NonLinearWaterTankChangingParamUniformGoalStacking4-SquareDistance-v2", noise_scale=0.0)
    obs0 = env.reset()
    assert np.array_equal(obs0.reshape(4, 3), np.tile(obs0[:3], (4, 1)))   # KAT-5: every frame = the reset state
    obs1 = env.set_state(0.0, 0.0)
    obs1 = env.set_r(2.0)
    assert np.array_equal(obs1, obs0) and env.h1 == 0.0 and env.r == 2.0
    obs2, _, _, _ = env.step(np.array([0.5]))
    assert np.array_equal(obs2[:9], obs0[3:]) and obs2[11] == 2.0 and obs2[9] > 0.0   # the new frame comes last


def test_in_kernel_reset_honours_if_reset_all_false(G):
    """rollout(auto_reset=True, resample_params=False) = reset_r() at every episode boundary (advisor finding, round 1)."""
    import pime_b200.vec as V
    env = V.WaterTankVec(512, dtype=torch.float32, seed=3)
    env.reset()
    p0 = [t.clone() for t in env.get_changable_parameters()]
    r0 = env.r.clone()
    env.rollout(450, np.array([0.0, -0.4, 0.4, 0.0]), auto_reset=True, resample_params=False)
    assert all(torch.equal(a, b) for a, b in zip(p0, env.get_changable_parameters())) and not torch.equal(r0, env.r)
    assert int(env.episode.min()) == 3
    env.rollout(200, np.array([0.0, -0.4, 0.4, 0.0]), auto_reset=True)                 # default: reset_all()
    assert not torch.equal(p0[0], env.a1)
    ph = V.PHVec(512, dtype=torch.float32, seed=3)
    ph.reset()
    q0, A0 = ph.qww_V.clone(), ph.A.clone()
    ph.rollout(120, -np.array([-0.02, 0.02, 0.035]), auto_reset=True, resample_params=False)
    assert torch.equal(q0, ph.qww_V) and torch.equal(A0, ph.A)


def test_device_fault_flag_accumulates_until_checked():
    """A table overflow in an early launch must still be reported after later, clean launches (advisor finding, round 1)."""
    import pime_b200.vec as V
    ph = V.PHVec(64, dtype=torch.float64, seed=1)
    ph.reset()
    x0 = ph.x.clone()
    ph.x.fill_(1e6); ph.A.fill_(1.0); ph.B.fill_(0.0); ph.C.fill_(0.0025)
    K = -np.array([-0.02, 0.02, 0.035])
    ph.rollout(2, K)                         # runs past the table
    ph.x.copy_(x0); ph.reset()
    ph.rollout(2, K)                         # a clean launch afterwards
    with pytest.raises(IndexError):
        ph.check_status()
    ph.check_status()                        # reading the flag cleared it
