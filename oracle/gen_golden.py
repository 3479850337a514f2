"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the dev container (needs /root/reference):   python oracle/gen_golden.py
The fixtures travel to the GPU box; this script and the reference do not need to.

Every fixture is produced by calling the reference's own classes (gym_control.envs.*, elegantrl.*) through
oracle/ref_loader.py.  Randomness is injected from the outside only (env attributes are set through the
reference's public test API -- set_state / set_r / reset_changable_parameters / set_params -- and
``get_noise`` is replaced by a recorded sequence), the arithmetic is the reference's.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

WT_INT = "NonLinearWaterTankChangingParamUniformGoalIntegrator-SquareDistance-v2"
WT_S3 = "NonLinearWaterTankChangingParamUniformGoal-SquareDistance-v2"
WT_STACK = "NonLinearWaterTankChangingParamUniformGoalStacking{}-SquareDistance-v2"
PH_INT = "PH1DChangingParamUniformGoalIntegrator-SqaureDistance-v35"
PH_NOIB = "PH1DChangingParamUniformGoalIntegrator-SqaureDistance-NoIB-v35"


def scal(v):
    return float(np.asarray(v, dtype=np.float64).reshape(-1)[0])


class NoiseTape:
    """Replacement for env.get_noise (nonlinear_watertank.py:271-272) that plays back recorded draws."""

    def __init__(self, values):
        self.values = list(values)
        self.i = 0

    def __call__(self):
        v = self.values[self.i]
        self.i += 1
        return v


def gen_wt_step(ref, rng):
    out = {}
    n = 192
    a1 = rng.uniform(0.0015, 0.0024, n); a2 = rng.uniform(0.0015, 0.0024, n); Kp = rng.uniform(0.07, 0.17, n)
    h1 = rng.uniform(0, 12, n); h2 = rng.uniform(0, 12, n); r = rng.uniform(0, 10, n)
    I = rng.uniform(-25, 25, n); t = rng.integers(0, 199, n).astype(np.int32)
    act = rng.uniform(-1.5, 1.5, n)
    nz1 = rng.normal(0, 0.01, n); nz2 = rng.normal(0, 0.01, n)
    # edge cases: empty tanks, integrator at the clip, last step of the episode, strongly negative action, zero noise
    h1[:8] = 0.0; h2[4:12] = 0.0; h1[12] = 1e-9; h2[13] = 1e-12
    I[16:20] = 24.9; I[20:24] = -24.9; t[24:32] = 198; t[32] = 199; t[33] = 250
    act[34:40] = -3.0; act[40:44] = 3.0; nz1[44:64] = 0.0; nz2[44:64] = 0.0; nz1[64:68] = -0.5; nz2[64:68] = -0.5
    out.update(a1=a1, a2=a2, Kp=Kp, h1=h1, h2=h2, r=r, I=I, t=t, action=act, noise1=nz1, noise2=nz2)
    for rt in ("distance", "square_distance", "sparse"):
        for env_id, tag in ((WT_INT, "int"), (WT_S3, "s3")):
            env = ref.gym.make(env_id, reward_type=rt)
            res = {k: [] for k in ("h1", "h2", "I", "reward", "done", "obs")}
            for i in range(n):
                env.reset()
                env.reset_changable_parameters(float(a1[i]), float(a2[i]), float(Kp[i]))
                env.set_state(float(h1[i]), float(h2[i])); env.set_r(float(r[i]))
                if tag == "int":
                    env.integrator = float(I[i])
                env._episode_steps = int(t[i])
                env.get_noise = NoiseTape([float(nz1[i]), float(nz2[i])])
                obs, rew, done, _ = env.step(np.array([act[i]]))
                res["h1"].append(float(env.h1)); res["h2"].append(float(env.h2))
                res["I"].append(float(env.integrator) if tag == "int" else 0.0)
                res["reward"].append(scal(rew)); res["done"].append(bool(done)); res["obs"].append(np.asarray(obs, np.float64))
            for k, v in res.items():
                out[f"{rt}.{tag}.{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "wt_step.npz"), **out)


def gen_wt_traj(ref, rng):
    """Prior-only closed loop (action = obs64 @ priorK), staircase set-points of utils/test.py:78-106, 200 steps."""
    out = {}
    cases = [(0.0019, 0.0019, 0.12), (0.0015, 0.0024, 0.07), (0.0024, 0.0015, 0.17), (0.002, 0.0022, 0.1)]
    T = 200
    for ci, (a1, a2, Kp) in enumerate(cases):
        for noisy in (0, 1):
            env = ref.gym.make(WT_INT, reward_type="distance", noise_scale=0.01 if noisy else 0.0)
            env.reset()
            env.reset_changable_parameters(a1, a2, Kp)
            env.set_state(0.0, 0.0); env.set_r(3.0 if ci == 0 else 2.0); env.integrator = 0.0
            tape = rng.normal(0, 0.01, 2 * T) if noisy else np.zeros(2 * T)
            env.get_noise = NoiseTape(tape)
            priorK = -env.K.reshape(-1, 1)
            obs = env._get_observe()
            rows = []
            for s in range(T):
                if ci > 0 and s in (50, 100, 150):
                    obs = np.append(env.set_r(float(env.r) + 2.0)[:3], env.integrator)  # staircase
                a = obs @ priorK
                obs, rew, done, _ = env.step(a)
                rows.append([scal(a), float(env.h1), float(env.h2), float(env.r), float(env.integrator), scal(rew), float(done)])
            out[f"case{ci}.noisy{noisy}"] = np.asarray(rows)
            out[f"case{ci}.noisy{noisy}.tape"] = np.asarray(tape)
        out[f"case{ci}.params"] = np.asarray([a1, a2, Kp])
    np.savez_compressed(os.path.join(OUT, "wt_traj.npz"), **out)


def gen_wt_stack(ref, rng):
    out = {}
    for k in (1, 4, 10):
        env = ref.gym.make(WT_STACK.format(k), reward_type="distance", noise_scale=0.0)
        np.random.seed(100 + k)
        obs0 = env.reset()
        a1, a2, Kp = env.get_changable_parameters()
        out[f"k{k}.params"] = np.asarray([a1, a2, Kp, env.h1, env.h2, env.r])
        out[f"k{k}.K"] = np.asarray(env.K, np.float64)
        priorK = -env.K.reshape(-1, 1)
        obs_rows, rows = [np.asarray(obs0, np.float64)], []
        obs = obs0
        for s in range(30):
            a = obs.astype(np.float32) @ priorK + 0.3 * np.sin(0.37 * s)
            obs, rew, done, _ = env.step(a)
            obs_rows.append(np.asarray(obs, np.float64))
            rows.append([scal(a), float(env.h1), float(env.h2), scal(rew), float(done)])
        out[f"k{k}.obs"] = np.asarray(obs_rows)
        out[f"k{k}.rows"] = np.asarray(rows)
    np.savez_compressed(os.path.join(OUT, "wt_stack.npz"), **out)


def gen_ph(ref, rng):
    out = {}
    env = ref.gym.make(PH_INT)
    u = env.unwrapped
    table = np.asarray(u.pH, np.float64)
    mh = np.asarray(u.MHCl, np.float64)
    # facts the index shortcut rests on
    k = np.arange(mh.shape[0], dtype=np.float64)
    out["mhcl_is_k_times_step"] = np.asarray(bool(np.array_equal(mh, k * 1e-5)))
    out["mhcl_ge_k_over_1e5"] = np.asarray(bool(np.all(mh >= k / 1e5)))
    out["mhcl_prev_lt"] = np.asarray(bool(np.all(mh[:-1] < (k[1:] / 1e5))))
    sel = np.unique(np.concatenate([np.arange(0, 100000, 61), np.arange(1700, 2300), np.arange(900, 1100, 3),
                                    [99999, 99998, 50000, 10000, 5000, 2500]]))
    out["table_idx"] = sel.astype(np.int64)
    out["table_val"] = table[sel]
    out["table_sha256"] = np.frombuffer(hashlib.sha256(table.tobytes()).digest(), dtype=np.uint8)
    out["table_sum"] = np.asarray(table.sum())

    # ---- update_system through the (stubbed) control.tf2ss/c2d == scipy path
    nq = 64
    qww = rng.uniform(0.005, 0.015, nq); qc = rng.uniform(0.0015, 0.0025, nq)
    qww[:3] = [0.01, 0.005, 0.015]; qc[:3] = [0.002, 0.0015, 0.0025]
    ABC = []
    for i in range(nq):
        u.set_params(float(qww[i]), float(qc[i])); u.update_system()
        ABC.append([u.dsys.A.item(), u.dsys.B.item(), u.dsys.C.item()])
    out["sys.qww_V"], out["sys.qc_V"], out["sys.ABC"] = qww, qc, np.asarray(ABC)

    # ---- single steps
    n = 256
    qi = rng.integers(0, nq, n)
    x = rng.uniform(0, 120, n); r = rng.uniform(3, 11, n); I = rng.uniform(-25, 25, n)
    t = rng.integers(0, 49, n).astype(np.int32); act = rng.uniform(-1.3, 1.3, n)
    x[:6] = 0.0; x[6:10] = 300.0; I[10:14] = 24.99; I[14:18] = -24.99; t[18:22] = 49; act[22:26] = -5.0; act[26:30] = 5.0
    out.update({"step.qi": qi, "step.x": x, "step.r": r, "step.I": I, "step.t": t, "step.action": act})
    for env_id, tag in ((PH_INT, "int"), (PH_NOIB, "noib")):
        for rt in ("square_distance", "distance", "sparse"):
            e = ref.gym.make(env_id, reward_type=rt)
            uu = e.unwrapped
            uu.pH = table  # same object contents; avoids rebuilding 6x (the constructor already built an equal table)
            res = {k_: [] for k_ in ("x", "y", "I", "reward", "done", "idx")}
            for i in range(n):
                e.reset()
                uu.set_params(float(qww[qi[i]]), float(qc[qi[i]])); uu.update_system()
                uu.set_state(float(x[i])); uu.set_r(float(r[i])); uu.integrator = float(I[i])
                uu._episode_steps = int(t[i]); e._elapsed_steps = int(t[i]); uu.last_action = 0.0
                obs, rew, done, _ = e.step(np.array([act[i]]))
                res["x"].append(scal(uu.state)); res["y"].append(float(uu.y)); res["I"].append(float(uu.integrator))
                res["reward"].append(scal(rew)); res["done"].append(bool(done))
                res["idx"].append(int(np.argwhere(uu.MHCl >= np.around(uu.dsys.C.item() * uu.state, 5))[0][0]))
            for k_, v in res.items():
                out[f"step.{tag}.{rt}.{k_}"] = np.asarray(v)

    # ---- index shortcut: argwhere scan vs rint on many C*x values (incl. exact half-integers)
    cx = np.concatenate([rng.uniform(0, 0.75, 4000), (np.arange(0, 2000) + 0.5) * 1e-5, np.arange(0, 2000) * 1e-5])
    idx_ref = np.array([int(np.argwhere(mh >= np.around(v, 5))[0][0]) for v in cx])
    out["index.cx"], out["index.ref"] = cx, idx_ref

    # ---- closed-loop trajectories, prior only (KAT-4 and utils/test.py:1385-1388 set-points)
    cases = [(0.01, 0.002, 0.0, 7.0), (0.005, 0.0025, 10.0, 10.0), (0.015, 0.0015, 30.0, 3.0), (0.012, 0.0021, 5.0, 8.0)]
    for ci, (q1, q2, x0, r0) in enumerate(cases):
        e = ref.gym.make(PH_INT)
        uu = e.unwrapped; uu.pH = table
        e.reset()
        uu.set_params(q1, q2); uu.update_system()
        uu.set_state(x0); uu.set_r(r0); uu.integrator = 0.0
        pk = -uu.K.reshape(-1, 1)
        obs = uu._get_observe()
        rows = []
        for s in range(50):
            a = obs @ pk
            obs, rew, done, _ = e.step(a)
            rows.append([scal(a), scal(uu.state), float(uu.y), float(uu.integrator), scal(rew), float(done)])
        out[f"traj{ci}"] = np.asarray(rows)
        out[f"traj{ci}.setup"] = np.asarray([q1, q2, x0, r0, uu.dsys.A.item(), uu.dsys.B.item(), uu.dsys.C.item()])
    np.savez_compressed(os.path.join(OUT, "ph.npz"), **out)
    return table


def _sd_np(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def gen_actor(ref, rng):
    import torch

    out = {}
    specs = [("modular", 32, 4, 1), ("modular", 64, 3, 1), ("plain", 32, 30, 0), ("plain", 32, 3, 0)]
    for kind, H, S, D in specs:
        torch.manual_seed(1234 + H + S)
        if kind == "modular":
            act = ref.net_residual.ActorResidualIntegratorModularPPO(H, S, 1, D)
        else:
            act = ref.net_residual.ActorResidualPPO(H, S, 1)
        tag = f"{kind}.H{H}.S{S}"
        for k, v in _sd_np(act).items():
            out[f"{tag}.sd.{k}"] = v
        obs = rng.uniform(-3, 12, (64, S)).astype(np.float32)
        with torch.no_grad():
            tobs = torch.as_tensor(obs)
            if kind == "modular":
                a_avg = act.net(torch.cat([act.other_net(tobs[:, :act.other_dim]), act.integrator_net(tobs[:, act.other_dim:])], -1))
            else:
                a_avg = act.net(tobs)
            det = act(tobs)
        out[f"{tag}.obs"], out[f"{tag}.a_avg"], out[f"{tag}.det"] = obs, a_avg.numpy()[:, 0], det.numpy()[:, 0]
    # full-size nets: outputs + weight hash only; the weights are re-created from the seed by the product's
    # net_residual module (same constructor order => same init), which the test verifies through the hash.
    for kind, H, S, D in [("modular", 256, 4, 1), ("modular", 128, 3, 1), ("plain", 256, 30, 0), ("plain", 256, 3, 0)]:
        torch.manual_seed(0)
        act = (ref.net_residual.ActorResidualIntegratorModularPPO(H, S, 1, D) if kind == "modular"
               else ref.net_residual.ActorResidualPPO(H, S, 1))
        tag = f"{kind}.H{H}.S{S}"
        sd = _sd_np(act)
        h = hashlib.sha256()
        for k in sorted(sd):
            h.update(k.encode()); h.update(np.ascontiguousarray(sd[k]).tobytes())
        out[f"{tag}.sha256"] = np.frombuffer(h.digest(), dtype=np.uint8)
        out[f"{tag}.nparam"] = np.asarray(sum(p.numel() for p in act.parameters()))
        obs = rng.uniform(-3, 12, (32, S)).astype(np.float32)
        with torch.no_grad():
            tobs = torch.as_tensor(obs)
            if kind == "modular":
                a_avg = act.net(torch.cat([act.other_net(tobs[:, :act.other_dim]), act.integrator_net(tobs[:, act.other_dim:])], -1))
            else:
                a_avg = act.net(tobs)
            det = act(tobs)
        out[f"{tag}.obs"], out[f"{tag}.a_avg"], out[f"{tag}.det"] = obs, a_avg.numpy()[:, 0], det.numpy()[:, 0]
    np.savez_compressed(os.path.join(OUT, "actor.npz"), **out)


def gen_explore(ref, rng, table):
    """The reference's own rollout loops (agent_residual.py:52-69, run.py:600-619) on both plants."""
    import torch

    out = {}
    for plant in ("wt", "ph", "wtstack"):
        torch.manual_seed(7)
        np.random.seed(11)
        if plant == "wt":
            env = ref.gym.make(WT_INT, reward_type="distance")
            agent = ref.agent_residual.AgentResidualIntegratorModularPPO()
            H = 32
        elif plant == "wtstack":
            env = ref.gym.make(WT_STACK.format(4), reward_type="distance")
            agent = ref.agent_residual.AgentResidualPPO()
            H = 32
        else:
            env = ref.gym.make(PH_INT)
            env.unwrapped.pH = table
            agent = ref.agent_residual.AgentResidualIntegratorModularPPO()
            H = 32
        env.seed(3)
        penv = ref.env.PreprocessEnv(env, if_print=False)
        if plant == "wtstack":
            agent.init(H, penv.state_dim, penv.action_dim)
        else:
            agent.init(H, penv.state_dim, penv.action_dim, env.unwrapped.n_integrator)
        agent.init_residual({"init_K": env.unwrapped.K.reshape(-1, 1)})
        agent.fix_K()
        # make the residual non-trivial (the zero-initialised last layer would hide actor mismatches)
        with torch.no_grad():
            agent.act.net[-1].weight.normal_(0.0, 0.1)
            agent.act.net[-1].bias.fill_(0.05)
        for k, v in _sd_np(agent.act).items():
            out[f"{plant}.sd.{k}"] = v
        out[f"{plant}.priorK"] = np.asarray(agent.priorK, np.float64).reshape(-1)
        out[f"{plant}.H"] = np.asarray(H)

        u = env.unwrapped
        resets, pnoise = [], []
        orig_reset = u.reset

        def rec_reset(u=u, orig_reset=orig_reset, resets=resets, plant=plant):
            o = orig_reset()
            if plant == "ph":
                resets.append([u.qww_V, u.qc_V, scal(u.state), u.r, u.dsys.A.item(), u.dsys.B.item(), u.dsys.C.item()])
            else:
                resets.append([u.a1, u.a2, u.Kp, u.h1, u.h2, u.r])
            return o

        u.reset = rec_reset
        if plant != "ph":
            orig_noise = u.get_noise

            def rec_noise(orig_noise=orig_noise, pnoise=pnoise):
                v = orig_noise()
                pnoise.append(v)
                return v

            u.get_noise = rec_noise
        max_step = penv.max_step
        target = 2 * max_step
        buf = ref.replay.ReplayBuffer(max_len=target + max_step, state_dim=penv.state_dim, action_dim=1,
                                      if_on_policy=True, if_per=False, if_gpu=False)
        steps = agent.explore_env(penv, buf, target, 1.0, 0.99)
        buf.update_now_len_before_sample()
        out[f"{plant}.explore.steps"] = np.asarray(steps)
        out[f"{plant}.explore.max_step"] = np.asarray(max_step)
        out[f"{plant}.explore.buf_state"] = buf.buf_state[:buf.now_len].copy()
        out[f"{plant}.explore.buf_other"] = buf.buf_other[:buf.now_len].copy()
        out[f"{plant}.explore.resets"] = np.asarray(resets, np.float64)
        out[f"{plant}.explore.pnoise"] = np.asarray(pnoise, np.float64)
        # deterministic evaluation episodes
        resets.clear(); pnoise.clear()
        rets = []
        with torch.no_grad():
            for _ in range(2):
                ret, nstep = ref.run.get_episode_return(penv, agent.act, agent.device)
                rets.append([scal(ret), nstep])
        out[f"{plant}.eval.returns"] = np.asarray(rets, np.float64)
        out[f"{plant}.eval.resets"] = np.asarray(resets, np.float64)
        out[f"{plant}.eval.pnoise"] = np.asarray(pnoise, np.float64)
    np.savez_compressed(os.path.join(OUT, "explore.npz"), **out)


def gen_ppo(ref, rng):
    """PPO learner pins (agent.py:611-708): values, old log-probabilities, reward-to-go / GAE / plain advantage, and
    the losses, gradients and parameters after ONE Adam step of the reference on a fixed index set."""
    import torch

    out = {}
    for tag, make, S, H, D in [("modular", lambda: ref.agent_residual.AgentResidualIntegratorModularPPO(), 4, 32, 1),
                               ("plain", lambda: ref.agent_residual.AgentResidualPPO(), 3, 32, 0)]:
        torch.manual_seed(77 + S)
        agent = make()
        agent.lambda_gae_adv, agent.ratio_clip, agent.lambda_entropy = 0.95, 0.25, 0.02
        if D:
            agent.init(H, S, 1, D)
        else:
            agent.init(H, S, 1)
        agent.device = torch.device("cpu")
        agent.act.to("cpu"); agent.cri.to("cpu")
        K = rng.uniform(-0.4, 0.4, (S, 1))
        agent.init_residual({"init_K": K})
        with torch.no_grad():   # a non-trivial last layer (init_residual zeroes it)
            agent.act.net[-1].weight.copy_(torch.as_tensor(rng.normal(0, 0.1, (1, H)), dtype=torch.float32))
            agent.act.net[-1].bias.fill_(0.03)
        agent.optimizer = torch.optim.Adam([{"params": agent.act.parameters(), "lr": 3e-4}, {"params": agent.cri.parameters(), "lr": 3e-4}])
        for k, v in _sd_np(agent.act).items():
            out[f"{tag}.act0.{k}"] = v
        for k, v in _sd_np(agent.cri).items():
            out[f"{tag}.cri0.{k}"] = v
        # synthetic on-policy buffer: n episodes of T steps, episode after episode (the reference's order)
        n, T, gamma = 6, 20, 0.98
        L = n * T
        state = rng.uniform(0, 10, (L, S)).astype(np.float32)
        state[:, -1] = rng.uniform(-25, 25, L)
        reward = (-rng.uniform(0, 30, L)).astype(np.float32)
        mask = np.full(L, gamma, np.float32)
        mask[T - 1::T] = 0.0
        noise = rng.standard_normal((L, 1)).astype(np.float32)
        ts = torch.as_tensor(state)
        with torch.no_grad():
            a_avg = agent.act.net(ts) if not D else agent.act.net(torch.cat([agent.act.other_net(ts[:, :S - D]), agent.act.integrator_net(ts[:, S - D:])], -1))
            action = (a_avg + torch.as_tensor(noise) * agent.act.a_std_log.exp()).numpy()
            buf_value = agent.cri(ts)
            buf_logprob = -(torch.as_tensor(noise).pow(2).__mul__(0.5) + agent.act.a_std_log + agent.act.sqrt_2pi_log).sum(1)
            r_sum, adv_gae = agent.compute_reward_gae(L, torch.as_tensor(reward), torch.as_tensor(mask), buf_value)
            r_sum2, adv_plain = agent.compute_reward_adv(L, torch.as_tensor(reward), torch.as_tensor(mask), buf_value)
        out[f"{tag}.state"], out[f"{tag}.reward"], out[f"{tag}.mask"] = state, reward, mask
        out[f"{tag}.noise"], out[f"{tag}.action"] = noise, action
        out[f"{tag}.value"], out[f"{tag}.logprob"] = buf_value.numpy()[:, 0], buf_logprob.numpy()
        out[f"{tag}.r_sum"], out[f"{tag}.adv_gae"], out[f"{tag}.adv_plain"] = r_sum.numpy(), adv_gae.numpy(), adv_plain.numpy()
        assert np.array_equal(r_sum.numpy(), r_sum2.numpy())
        # one minibatch of the reference loop (agent.py:632-657) on fixed indices
        idx = rng.integers(0, L, 48)
        ti = torch.as_tensor(idx)
        st, ac, rs, lp, ad = ts[ti], torch.as_tensor(action)[ti], r_sum[ti], buf_logprob[ti], adv_gae[ti]
        new_logprob = agent.act.compute_logprob(st, ac)
        ratio = (new_logprob - lp).exp()
        obj_surrogate = -torch.min(ad * ratio, ad * ratio.clamp(1 - agent.ratio_clip, 1 + agent.ratio_clip)).mean()
        obj_entropy = (new_logprob.exp() * new_logprob).mean()
        obj_actor = obj_surrogate + obj_entropy * agent.lambda_entropy
        value = agent.cri(st).squeeze(1)
        obj_critic = agent.criterion(value, rs)
        obj_united = obj_actor + obj_critic / (rs.std() + 1e-5)
        agent.optimizer.zero_grad()
        obj_united.backward()
        out[f"{tag}.idx"] = idx
        out[f"{tag}.losses"] = np.array([obj_actor.item(), obj_critic.item(), obj_united.item(), obj_entropy.item()])
        for name, p in list(agent.act.named_parameters()) + [("cri." + k, v) for k, v in agent.cri.named_parameters()]:
            if p.grad is not None:
                out[f"{tag}.grad.{name}"] = p.grad.detach().numpy().copy()
        agent.optimizer.step()
        for k, v in _sd_np(agent.act).items():
            out[f"{tag}.act1.{k}"] = v
        for k, v in _sd_np(agent.cri).items():
            out[f"{tag}.cri1.{k}"] = v
        out[f"{tag}.K"] = K
    np.savez_compressed(os.path.join(OUT, "ppo.npz"), **out)


def gen_last_state(ref):
    """reset_from_last_state=True (nonlinear_watertank.py:904-910, :819-821; ph.py:417-420, :345-346): a scripted
    sequence of reset / step calls on the reference; every row records the env after the call."""
    out = {}
    np.random.seed(4242)
    env = ref.gym.make(WT_INT, reset_from_last_state=True, max_step=6, noise_scale=0.0, reward_type="distance").unwrapped
    rows, script = [], []

    def rec(kind, arg=0.0):
        script.append((kind, arg))
        rows.append([env.h1, env.h2, env.r, env.integrator, env._episode_steps,
                     np.nan if env.last_h1 is None else env.last_h1, np.nan if env.last_h2 is None else env.last_h2,
                     env.a1, env.a2, env.Kp])

    env.reset(); rec(0)                                   # last is None -> |randn| * 0.1
    first = (env.h1, env.h2)
    assert 0 <= first[0] < 1 and 0 <= first[1] < 1
    acts = [0.9, 0.5, -0.2, 0.7, 1.0, 0.1, -0.6, 0.3]
    for ep in range(2):                                   # two full episodes: done at t = 6 records the levels
        for k in range(6):
            _, _, done, _ = env.step(np.array([acts[(k + ep) % 8]])); rec(1, acts[(k + ep) % 8])
        assert done
        env.reset(); rec(0)                               # restarts from the recorded levels
    for k in range(3):                                    # reset in mid-episode goes back to the levels at the last done
        env.step(np.array([acts[k]])); rec(1, acts[k])
    env.reset(); rec(0)
    env.if_reset_all = False                              # README.md:27-33
    for k in range(6):
        env.step(np.array([0.4])); rec(1, 0.4)
    env.reset(); rec(0)                                   # reset_r path (:920-926)
    out["wt_rows"] = np.asarray(rows, np.float64)
    out["wt_script"] = np.asarray(script, np.float64)

    np.random.seed(4243)                                  # sample_parameters draws from numpy's global stream
    env = ref.gym.make(PH_INT, reset_from_last_state=True)
    env.seed(4243)                                        # x0 / r draws come from the env's own np_random (ph.py:124,420-424)
    u = env.unwrapped
    rows, script = [], []

    def recp(kind, arg=0.0, done=False):
        script.append((kind, arg))
        rows.append([scal(u.state), scal(u.y), u.r, scal(u.integrator), u._episode_steps,
                     np.nan if u.last_state is None else scal(u.last_state), u.qww_V, u.qc_V, float(done)])

    env.reset(); recp(0)
    for ep in range(2):
        for k in range(50):
            a = 0.8 * np.sin(0.37 * k + ep)
            _, _, done, _ = env.step(np.array([a])); recp(1, a, done)
        assert done
        env.reset(); recp(0)
    for k in range(5):
        env.step(np.array([0.25])); recp(1, 0.25)
    env.reset(); recp(0)
    out["ph_rows"] = np.asarray(rows, np.float64)
    out["ph_script"] = np.asarray(script, np.float64)
    np.savez_compressed(os.path.join(OUT, "last_state.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    rng = np.random.default_rng(20261018)
    gen_wt_step(ref, rng)
    gen_wt_traj(ref, rng)
    gen_wt_stack(ref, rng)
    table = gen_ph(ref, rng)
    gen_actor(ref, rng)
    gen_explore(ref, rng, table)
    gen_ppo(ref, np.random.default_rng(20261019))
    gen_last_state(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
