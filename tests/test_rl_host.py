"""Host-side logic of pime_b200.rl (no GPU): network definitions against the reference's initialisation and state-dict
keys, the PPO objective / Adam step against fixtures produced by the reference's own update code
(tests/golden/ppo.npz, oracle/gen_golden.py:gen_ppo), the buffer, the registry, the sharding helpers and the
world_size-2 gradient all-reduce (gloo)."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch

import pime_b200.rl as R


def _sha(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


@pytest.mark.parametrize("kind,H,S,D", [("modular", 256, 4, 1), ("modular", 128, 3, 1), ("plain", 256, 30, 0), ("plain", 256, 3, 0)])
def test_actor_init_matches_reference_bit_for_bit(golden, kind, H, S, D):
    """Same constructor order as net_residual.py => torch.manual_seed(0) gives the reference's weights (sha256 pinned)."""
    g = golden("actor")
    torch.manual_seed(0)
    act = R.ActorResidualIntegratorModularPPO(H, S, 1, D) if kind == "modular" else R.ActorResidualPPO(H, S, 1)
    tag = f"{kind}.H{H}.S{S}"
    assert sum(p.numel() for p in act.parameters()) == int(g[f"{tag}.nparam"])
    assert np.array_equal(_sha(act.state_dict()), g[f"{tag}.sha256"])
    with torch.no_grad():
        obs = torch.as_tensor(g[f"{tag}.obs"])
        np.testing.assert_allclose(act.a_avg(obs).numpy()[:, 0], g[f"{tag}.a_avg"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(act(obs).numpy()[:, 0], g[f"{tag}.det"], rtol=0, atol=1e-5)


def test_param_counts_of_survey_8a():
    assert sum(p.numel() for p in R.ActorResidualIntegratorModularPPO(256, 4, 1, 1).parameters()) == 133382
    assert sum(p.numel() for p in R.ActorResidualIntegratorModularPPO(128, 3, 1, 1).parameters()) == 33797
    assert sum(p.numel() for p in R.CriticAdv(4, 256).parameters()) == 133121


def _load_agent(g, tag):
    if tag == "modular":
        agent = R.AgentResidualIntegratorModularPPO()
        agent.init(32, 4, 1, 1)
    else:
        agent = R.AgentResidualPPO()
        agent.init(32, 3, 1)
    agent.device = torch.device("cpu")
    agent.act.to("cpu"); agent.cri.to("cpu")
    agent.lambda_gae_adv, agent.ratio_clip, agent.lambda_entropy = 0.95, 0.25, 0.02
    agent.init_residual({"init_K": g[f"{tag}.K"]})
    agent.act.load_state_dict({k[len(tag) + 6:]: torch.as_tensor(g[k]) for k in g.files if k.startswith(f"{tag}.act0.")})
    agent.cri.load_state_dict({k[len(tag) + 6:]: torch.as_tensor(g[k]) for k in g.files if k.startswith(f"{tag}.cri0.")})
    agent.optimizer = torch.optim.Adam([{"params": agent.act.parameters(), "lr": 3e-4}, {"params": agent.cri.parameters(), "lr": 3e-4}])
    return agent


@pytest.mark.parametrize("tag", ["modular", "plain"])
def test_state_dict_keys_and_ppo_step_match_reference(golden, tag):
    g = golden("ppo")
    agent = _load_agent(g, tag)
    ref_act_keys = sorted(k[len(tag) + 6:] for k in g.files if k.startswith(f"{tag}.act0."))
    assert sorted(agent.act.state_dict().keys()) == ref_act_keys           # actor.pth interchangeable
    assert sorted(agent.cri.state_dict().keys()) == sorted(k[len(tag) + 6:] for k in g.files if k.startswith(f"{tag}.cri0."))
    state, action = torch.as_tensor(g[f"{tag}.state"]), torch.as_tensor(g[f"{tag}.action"])
    with torch.no_grad():
        np.testing.assert_allclose(agent.cri(state).numpy()[:, 0], g[f"{tag}.value"], rtol=1e-6, atol=1e-6)
        noise = torch.as_tensor(g[f"{tag}.noise"])
        lp = -(noise.pow(2) * 0.5 + agent.act.a_std_log + agent.act.sqrt_2pi_log).sum(1)
        np.testing.assert_allclose(lp.numpy(), g[f"{tag}.logprob"], rtol=1e-6, atol=1e-6)
    idx = torch.as_tensor(g[f"{tag}.idx"])
    oa, oc, ou, oe = agent.ppo_objectives(state[idx], action[idx], torch.as_tensor(g[f"{tag}.r_sum"])[idx],
                                          torch.as_tensor(g[f"{tag}.logprob"])[idx], torch.as_tensor(g[f"{tag}.adv_gae"])[idx])
    np.testing.assert_allclose([oa.item(), oc.item(), ou.item(), oe.item()], g[f"{tag}.losses"], rtol=2e-5, atol=1e-6)
    agent.optimizer.zero_grad()
    ou.backward()
    for name, p in list(agent.act.named_parameters()) + [("cri." + k, v) for k, v in agent.cri.named_parameters()]:
        key = f"{tag}.grad.{name}"
        if key in g.files:
            np.testing.assert_allclose(p.grad.numpy(), g[key], rtol=1e-4, atol=1e-6, err_msg=name)
        else:
            assert p.grad is None or not p.requires_grad or float(p.grad.abs().max()) == 0.0
    agent.optimizer.step()
    for k in g.files:
        if k.startswith(f"{tag}.act1."):
            np.testing.assert_allclose(agent.act.state_dict()[k[len(tag) + 6:]].numpy(), g[k], rtol=1e-5, atol=1e-6, err_msg=k)
        if k.startswith(f"{tag}.cri1."):
            np.testing.assert_allclose(agent.cri.state_dict()[k[len(tag) + 6:]].numpy(), g[k], rtol=1e-5, atol=1e-6, err_msg=k)


def test_advantage_normalisation_is_torch_std():
    x = torch.randn(1000) * 3 + 1
    np.testing.assert_allclose(R.AgentPPO._normalise(x).numpy(), ((x - x.mean()) / (x.std() + 1e-5)).numpy(), rtol=1e-5, atol=1e-6)


def test_init_actor_zero_makes_policy_equal_prior():
    agent = R.AgentResidualIntegratorModularPPO()
    agent.init(32, 4, 1, 1)
    K = np.array([0.0, 0.4, -0.4, 0.0])
    agent.init_residual({"init_K": K.reshape(-1, 1)})
    assert agent.act.priorK.requires_grad is False and np.allclose(agent.priorK[:, 0], -K)
    obs = np.array([3.0, 2.0, 5.0, 1.0], np.float32)
    a, _ = agent.select_action(obs, if_deterministic=True)
    assert abs(float(a[0]) - float(obs @ -K)) < 1e-6                        # residual is exactly zero at step 0
    a_raw, nz = agent.select_action(obs)
    assert abs(float(a_raw[0]) - float(nz[0]) * float(np.exp(-0.5))) < 1e-6


class _ToyEnv:
    """Foreign gym env: exercises the reference's sequential explore loop (no device env involved)."""
    max_step = 5

    class _Box:
        def __init__(self, k):
            self.shape, self.high = (k,), np.ones(k)
    observation_space, action_space = _Box(3), _Box(1)

    def reset(self):
        self.t = 0
        return np.array([1.0, 2.0, 3.0])

    def step(self, a):
        self.t += 1
        return np.array([1.0, 2.0, 3.0]) + self.t, -float(np.abs(a).sum()), self.t >= 5, {}


def test_foreign_env_explore_and_buffer_staging():
    env = R.PreprocessEnv(_ToyEnv())
    assert (env.state_dim, env.action_dim, env.max_step, env.if_discrete) == (3, 1, 5, False)
    agent = R.AgentResidualPPO()
    agent.init(32, 3, 1)
    agent.init_residual({"init_K": np.array([[0.1], [0.0], [-0.1]])})
    buf = R.ReplayBuffer(64, 3, 1, if_on_policy=True, if_per=False, if_gpu=True)
    steps = agent.explore_env(env, buf, 12, reward_scale=2.0, gamma=0.9)
    assert steps == 15                                                      # whole episodes: 3 x 5
    buf.update_now_len_before_sample()
    r, m, a, nz, s = buf.sample_all()
    assert buf.now_len == 15 and s.shape == (15, 3) and a.shape == (15, 1)
    assert np.allclose(m.cpu().numpy().reshape(3, 5), [[0.9, 0.9, 0.9, 0.9, 0.0]] * 3)
    assert s.dtype == torch.float32 and np.allclose(s[0].cpu().numpy(), [1, 2, 3])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            agent.update_net(buf, 12, 8, 1)                                 # the learner's value / GAE passes are CUDA-only


def test_registry_and_arguments():
    assert set(R.MODELS) == {"ppo", "residualppo", "residualintegratormodularppo"}
    assert all(R.IF_ONPOLICY[k] for k in R.MODELS)
    args = R.Arguments(if_on_policy=True)
    with pytest.raises(RuntimeError):
        args.init_before_training()
    args.agent = R.AgentPPO
    with pytest.raises(RuntimeError):
        args.init_before_training()


def test_shard_range_partitions_envs():
    for n, w in [(1 << 23, 8), (10, 3), (7, 8), (1, 1)]:
        spans = [R.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Linear(4, 3)
    x = torch.full((5, 4), float(rank + 1))
    net(x).sum().backward()
    params = list(net.parameters())
    R.allreduce_mean_grads(params)
    adv = torch.arange(6, dtype=torch.float32) + 10 * rank          # global moments for the advantage normalisation
    q.put((rank, params[0].grad.clone().numpy(), R.AgentPPO._normalise(adv).numpy(), R.shard_range(11, rank, world)))
    dist.destroy_process_group()


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = R.shard_range(5, rank, world)                           # 5 envs over 2 ranks: 3 + 2 (uneven on purpose)
    n, T, S = hi - lo, 4, 3
    buf = R.ReplayBuffer(T * n, S, 1, True, False, True, num_envs=n, device="cpu")
    bs, bo = buf.rollout_views(T)
    t = torch.arange(T, dtype=torch.float32).view(T, 1, 1)
    e = torch.arange(lo, hi, dtype=torch.float32).view(1, n, 1)
    bs.copy_((100 * t + e).expand(T, n, S) + torch.arange(S, dtype=torch.float32) * 0.1)      # row (t, global env) is recognisable
    bo.copy_((1000 + 100 * t + e).expand(T, n, 4))
    buf.commit_rollout(T)
    full = buf.gather()
    q.put((rank, full.num_envs, full.now_len, full.buf_state.clone().numpy(), full.buf_other.clone().numpy()))
    dist.destroy_process_group()


def test_world_size_2_replay_gather_is_time_major_over_all_envs():
    """ReplayBuffer.gather (single-learner mode): both ranks end up with the [T, 5, .] rows one process owning all five envs
    would hold, although the ranks own 3 and 2 envs."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    T, N, S = 4, 5, 3
    t = np.arange(T, dtype=np.float32).reshape(T, 1, 1)
    e = np.arange(N, dtype=np.float32).reshape(1, N, 1)
    want_s = (np.broadcast_to(100 * t + e, (T, N, S)) + np.arange(S, dtype=np.float32) * 0.1).reshape(T * N, S).astype(np.float32)
    want_o = np.broadcast_to(1000 + 100 * t + e, (T, N, 4)).reshape(T * N, 4).astype(np.float32)
    for rank, n_envs, now_len, bs, bo in res:
        assert n_envs == N and now_len == T * N
        assert np.array_equal(bs, want_s) and np.array_equal(bo, want_o)


def test_replay_gather_without_a_process_group_is_the_buffer_itself():
    buf = R.ReplayBuffer(8, 3, 1, True, False, True, num_envs=2, device="cpu")
    assert buf.gather() is buf


def test_world_size_2_grad_allreduce_and_global_advantage_moments():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    g0, g1 = res[0][1], res[1][1]
    assert np.allclose(g0, g1) and np.allclose(g0, 5 * 1.5)            # mean of the per-rank grads 5*1 and 5*2
    allv = np.concatenate([np.arange(6.0), np.arange(6.0) + 10])
    want = (allv - allv.mean()) / (allv.std(ddof=1) + 1e-5)
    assert np.allclose(np.concatenate([res[0][2], res[1][2]]), want, atol=1e-5)
    assert res[0][3] == (0, 6) and res[1][3] == (6, 11)
