#!/usr/bin/env python
"""train.py of the reference (ruoqizzz/PIME..., train.py:22-272) on the B200 stack: same command line, same agents,
same checkpoints -- the env is a batch of ``--num_envs`` ensemble members stepped by the fused CUDA rollout kernel.

    python examples/train.py --fix_K --algo ResidualIntegratorModularPPO --net_dim 128 \
        --env PH1DChangingParamUniformGoalIntegrator-SqaureDistance-v35 --target_step 1000 --batch_size 128 \
        --repeat_times 8 --lambda_gae_adv 0.99 --ratio_clip 0.2 --gamma 0.98 --break_step 200000          # run_ph_changing.sh
    python examples/train.py --fix_K --algo ResidualPPO --net_dim 256 --robust_test \
        --env NonLinearWaterTankChangingParamUniformGoalStacking10-SquareDistance-v2 --target_step 2000 \
        --batch_size 256 --repeat_times 10 --break_step 400000                                            # run_watertank_changing.sh

Differences from the reference script: no matplotlib figures (the arrays they plot come from pime_b200.scenarios and
are saved as .npz next to the checkpoints), ``--num_envs`` / ``--fp32`` are new, td3 / sac are not available.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pime_b200.gym_api as gym  # noqa: E402
import pime_b200.scenarios as scenarios  # noqa: E402
from pime_b200.rl import IF_ONPOLICY, MODELS, Arguments, PreprocessEnv, configure_logger, train_and_evaluate  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--algo", default="PPO", type=str)
    p.add_argument("--env", default="NonLinearWaterTankChangingParamUniformGoalIntegrator-SquareDistance-v2", type=str)
    p.add_argument("--reward_type", default="distance", type=str, choices=["distance", "sparse", "square_distance"])
    p.add_argument("--fix_K", action="store_true")
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--target_return", default=1e6, type=float)
    p.add_argument("--target_step", default=2 ** 3, type=int)
    p.add_argument("--reward_scale", default=1.0, type=float)
    p.add_argument("--break_step", default=2 ** 20, type=int)
    p.add_argument("--gamma", default=0.995, type=float)
    p.add_argument("--batch_size", default=2 ** 6, type=int)
    p.add_argument("--learning_rate", default=3e-4, type=float)
    p.add_argument("--net_dim", default=2 ** 4, type=int)
    p.add_argument("--verbose", default=0, type=int)
    p.add_argument("--tensorboard_log", default="tensorboard", type=str)
    p.add_argument("--env_zero_noise", action="store_true")
    p.add_argument("--eval_times1", default=2 ** 3, type=int)
    p.add_argument("--eval_times2", default=2 ** 4, type=int)
    p.add_argument("--eval_gap", default=5, type=int)
    p.add_argument("--robust_test", action="store_true")
    p.add_argument("--goal", default=4.0, type=float)
    p.add_argument("--repeat_times", default=2 ** 4, type=int)
    p.add_argument("--lambda_gae_adv", default=0.97, type=float)
    p.add_argument("--lambda_entropy", default=0.02, type=float)
    p.add_argument("--ratio_clip", default=0.2, type=float)
    p.add_argument("--test_render_times", default=10000, type=int)
    p.add_argument("--load", default="None", type=str)
    p.add_argument("--frozen_modular_integrator", action="store_true")
    p.add_argument("--frozen_transfer", action="store_true")
    # new
    p.add_argument("--num_envs", default=64, type=int, help="ensemble members stepped in parallel by one launch")
    p.add_argument("--fp64", action="store_true", help="plant arithmetic in float64 (the reference's), default float32")
    p.add_argument("--out", default=None, help="log root (default log_<break_step>)")
    return p.parse_args()


def make_env(args, num_envs):
    kw = dict(num_envs=num_envs, dtype=torch.float64 if args.fp64 else torch.float32)
    if "NonLinearWaterTank" in args.env:   # train.py:84-103
        kw.update(reward_type=args.reward_type, r=args.goal)
    if args.env_zero_noise:
        kw.update(noise_scale=0.0)
    env = gym.make(args.env, **kw)
    env.seed(args.seed)
    env.target_return = args.target_return
    return env


def save_staircase(env, agent, save_dir):
    """The data behind the reference's test_watertank / test_ph_integrator figures (utils/test.py:1056-1113,1480-1573)."""
    os.makedirs(save_dir, exist_ok=True)
    vec = env.env._clone_vec()
    K = env.K
    out = {}
    for policy, actor in (("agent", agent._pack("act")), ("linear", None)):
        res = scenarios.staircase(vec, policy, K, actor=actor, resample_params=env.if_reset_all)
        for k, v in res.items():
            if v is not None:
                out[f"{policy}.{k}"] = v.cpu().numpy()
    np.savez_compressed(os.path.join(save_dir, "staircase.npz"), **out)


def main():
    args = parse()
    algo = args.algo.lower()
    assert algo in MODELS, f"--algo must be one of {sorted(MODELS)} (off-policy agents are outside this package)"
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    env = make_env(args, args.num_envs)
    kargs = Arguments(if_on_policy=IF_ONPOLICY[algo])
    kargs.env = PreprocessEnv(env)
    kargs.env_eval = PreprocessEnv(make_env(args, max(args.eval_times2, 1)))
    kargs.env_eval.env.vec.seed = args.seed + 104729      # distinct Philox streams: evaluation must not replay the training draws
    kargs.agent = MODELS[algo]()
    kargs.agent.learning_rate = args.learning_rate
    kargs.agent.lambda_entropy, kargs.agent.ratio_clip, kargs.agent.lambda_gae_adv = args.lambda_entropy, args.ratio_clip, args.lambda_gae_adv
    kargs.random_seed, kargs.gamma, kargs.reward_scale = args.seed, args.gamma, args.reward_scale
    kargs.net_dim, kargs.batch_size, kargs.repeat_times = args.net_dim, args.batch_size, args.repeat_times
    kargs.target_step, kargs.max_memo, kargs.break_step = args.target_step, args.target_step, args.break_step
    kargs.eval_times1, kargs.eval_times2, kargs.eval_gap = args.eval_times1, args.eval_times2, args.eval_gap
    kargs.fix_K, kargs.frozen_modular_integrator, kargs.frozen_transfer = args.fix_K, args.frozen_modular_integrator, args.frozen_transfer
    kargs.load, kargs.if_remove, kargs.if_residual = args.load, False, True
    if "residual" in algo and hasattr(env, "K"):
        kargs.residual_kwargs = {"init_K": env.K.reshape(-1, 1)}           # train.py:119-123
    if "modular" in algo and hasattr(env, "n_integrator"):
        kargs.Modular_kwargs = {"integrator_dim": env.n_integrator}        # train.py:124-125
    root = args.out or f"log_{args.break_step}"
    tag = f"{args.algo}-{args.net_dim}{'-fixK' if args.fix_K else ''}"
    now = time.strftime("%Y-%m-%d-%H_%M_%S")
    kargs.cwd = os.path.join(root, f"{args.env}{'-zero' if args.env_zero_noise else ''}", tag, f"seed{args.seed}", now)
    configure_logger(args.verbose, os.path.join(root, f"{args.tensorboard_log}_{args.env}"), tag, True)
    print(f"Agent: {args.algo}, Env: {args.env} x {args.num_envs}, Seed: {args.seed}")
    agent, _ = train_and_evaluate(kargs)
    final = os.path.join(kargs.cwd, "final_model")
    os.makedirs(final, exist_ok=True)
    agent.save_load_model(final, if_save=True)                             # train.py:166-169
    if "Stacking" not in args.env:
        save_staircase(kargs.env, agent, final)
        if args.robust_test and "NonLinearWaterTank" in args.env:          # train.py:203-205
            cfg = env._cfg_kwargs()      # the sweep runs under the training env's reward type / noise setting
            cfg.pop("max_step", None)
            res, p = scenarios.robust_sweep(env.K, actor=agent._pack("act"), obs_mode=env._obs_mode, max_step=500,
                                            dtype=env._dtype, seed=args.seed + 7919, **cfg)
            np.savez_compressed(os.path.join(kargs.cwd, "robust_test.npz"), params=p, **{k: v.cpu().numpy() for k, v in res.items() if v is not None})
    with open(os.path.join(kargs.cwd, "args.txt"), "w") as f:
        f.write(str(args))
    print(f"Finish Training and Saved in {kargs.cwd}")


if __name__ == "__main__":
    main()
