"""Times the UNMODIFIED reference on this machine's CPU: the Python explore_env loop of
agent_residual.py:52-69 through oracle/ref_loader (single process, as the reference runs), then the env-only loop in P
independent processes (one env object each).  Needs /root/reference, so it runs in the dev container only; the result is
committed under profiles/ and quoted beside the C-port baseline that bench.py measures on the GPU box.

    python oracle/time_reference.py [seconds per measurement]
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
WT_INT = "NonLinearWaterTankChangingParamUniformGoalIntegrator-SquareDistance-v2"
PH_INT = "PH1DChangingParamUniformGoalIntegrator-SqaureDistance-v35"


def _make(plant, H):
    import torch
    from gen_golden import load_reference
    ref = load_reference()
    torch.manual_seed(0)
    np.random.seed(0)
    env = ref.gym.make(WT_INT, reward_type="distance") if plant == "wt" else ref.gym.make(PH_INT)
    env.seed(0)
    penv = ref.env.PreprocessEnv(env, if_print=False)
    agent = ref.agent_residual.AgentResidualIntegratorModularPPO()
    agent.init(H, penv.state_dim, penv.action_dim, env.unwrapped.n_integrator)
    agent.init_residual({"init_K": env.unwrapped.K.reshape(-1, 1)})
    agent.fix_K()
    buf = ref.replay.ReplayBuffer(max_len=4096 + penv.max_step, state_dim=penv.state_dim, action_dim=1, if_on_policy=True, if_per=False, if_gpu=False)
    return ref, penv, agent, buf


def explore(plant, H, seconds):
    """reference explore_env (actor forward on torch CPU + env step), median of 3 measurements of >= `seconds`."""
    import torch
    torch.set_num_threads(8)                      # elegantrl/run.py:40
    ref, penv, agent, buf = _make(plant, H)
    target = 2000 if plant == "wt" else 1000      # run_*_changing.sh target_step
    agent.explore_env(penv, buf, penv.max_step, 1.0, 0.99)   # warm-up: one episode
    rates = []
    for _ in range(3):
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            n += agent.explore_env(penv, buf, target, 1.0, 0.99)
        rates.append(n / (time.perf_counter() - t0))
    return float(np.median(rates))


def _env_only(args):
    plant, seconds, seed = args
    ref, penv, agent, buf = _make(plant, 32)
    env = penv
    K = np.asarray(agent.priorK, np.float64)
    np.random.seed(seed)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        s = env.reset()
        for _ in range(env.max_step):
            s, r, d, _ = env.step(np.asarray(s @ K).reshape(-1))   # prior-only policy: env step + P/PI prior
            n += 1
    return n / (time.perf_counter() - t0)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
    cores = os.cpu_count() or 1
    out = {"machine": "dev container (no GPU)", "cores": cores, "seconds_per_measurement": seconds}
    out["wt_explore_env_modular256_1proc"] = explore("wt", 256, seconds)
    out["ph_explore_env_modular128_1proc"] = explore("ph", 128, seconds)
    for plant in ("wt", "ph"):
        with mp.get_context("spawn").Pool(cores) as pool:
            rates = pool.map(_env_only, [(plant, seconds, 100 + i) for i in range(cores)])
        out[f"{plant}_env_plus_prior_{cores}proc_total"] = float(np.sum(rates))
        out[f"{plant}_env_plus_prior_1proc"] = float(np.median(rates))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
