import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import pime_b200.rl as R
import pime_b200.gym_api as G
WT = "NonLinearWaterTankChangingParamUniformGoalIntegrator-SquareDistance-v2"
H, B = int(sys.argv[1]), int(sys.argv[2])
n = 64
env = R.PreprocessEnv(G.make(WT, num_envs=n, dtype=torch.float32))
agent = R.AgentResidualIntegratorModularPPO()
agent.init(H, env.state_dim, env.action_dim, env.n_integrator)
agent.init_residual({"init_K": env.K.reshape(-1, 1)})
buf = R.ReplayBuffer(n * env.max_step, env.state_dim, 1, True, False, True, num_envs=n)
steps = agent.explore_env(env, buf, n * env.max_step, 1.0, 0.99)
agent.update_net(buf, steps, B, 1)
torch.cuda.synchronize()
agent.update_net(buf, steps, B, 1)
torch.cuda.synchronize()
