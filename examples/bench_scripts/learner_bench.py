import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import pime_b200.rl as R
import pime_b200.gym_api as G
WT = "NonLinearWaterTankChangingParamUniformGoalIntegrator-SquareDistance-v2"
for H, B in ((256, 256), (128, 128), (256, 512), (256, 2048), (256, 4096)):
    for mode in ("fused", "graph", "eager"):
        torch.manual_seed(0)
        n = 64
        env = R.PreprocessEnv(G.make(WT, num_envs=n, dtype=torch.float32))
        agent = R.AgentResidualIntegratorModularPPO()
        agent.init(H, env.state_dim, env.action_dim, env.n_integrator)
        agent.use_fused_learner = mode == "fused"
        agent.use_cuda_graph = mode == "graph"
        agent.init_residual({"init_K": env.K.reshape(-1, 1)})
        buf = R.ReplayBuffer(n * env.max_step, env.state_dim, 1, True, False, True, num_envs=n)
        steps = agent.explore_env(env, buf, n * env.max_step, 1.0, 0.99)
        rt = max(1, int(200 * B / steps))
        agent.update_net(buf, steps, B, rt)      # warm-up (graph capture etc.)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        agent.update_net(buf, steps, B, rt)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        iters = int(rt * steps / B)
        print(f"H={H} B={B} {mode:6s} {iters} steps  {dt / iters * 1e6:8.1f} us/minibatch", flush=True)
