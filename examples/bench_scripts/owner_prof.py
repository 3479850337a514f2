import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
import pime_b200.vec as V
for wl in sys.argv[1:]:
    args = type("A", (), dict(workload=wl, net_dim=0, T=0, envs=0))()
    w = bench.resolve(args)
    n, T, S, H = 1 << 17, w["T"], w["S"], w["H"]
    sd = bench.actor_state_dict(H, S, kind=w["kind"])
    actor = V.ActorPack(w["kind"], S, H, 1).update(sd)
    if wl == "wts10":
        env = V.WaterTankVec(n, dtype=torch.float32, obs_mode="stacking", num_stack=10, noise_scale=0.01)
    elif wl == "wts1":
        env = V.WaterTankVec(n, dtype=torch.float32, obs_mode="stacking", num_stack=1, noise_scale=0.01)
    elif wl == "wt":
        env = V.WaterTankVec(n, dtype=torch.float32, noise_scale=0.01)
    else:
        env = V.PHVec(n, dtype=torch.float32)
    env.reset()
    bs = torch.empty((T, n, S), dtype=torch.float32, device="cuda"); bo = torch.empty((T, n, 4), dtype=torch.float32, device="cuda")
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    K = np.array(w["K"])
    for it in range(3):
        stats.zero_()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        env.rollout(T, -K, actor=actor, auto_reset=True, replay=(bs, bo), stats=stats)
        ev1.record(); torch.cuda.synchronize()
    st = stats.cpu().numpy()
    ms = ev0.elapsed_time(ev1)
    print(wl, "ms", round(ms, 3), "env-steps/s %.3e" % (n * T / ms * 1e3), "owner wait clk/step-pair", st[6] / T, "work clk/step-pair", st[7] / T, flush=True)
