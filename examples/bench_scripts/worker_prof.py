import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
import pime_b200.vec as V, pime_b200._lib as L
wl = sys.argv[1] if len(sys.argv) > 1 else "wt"           # wt | ph
args = type("A", (), dict(workload=wl, net_dim=0, T=0, envs=0))()
w = bench.resolve(args)
n, T, S, H = 1 << 18, w["T"], w["S"], w["H"]
sd = bench.actor_state_dict(H, S, kind=w["kind"])
actor = V.ActorPack(w["kind"], S, H, 1).update(sd)
env = V.WaterTankVec(n, dtype=torch.float32, noise_scale=0.01) if wl == "wt" else V.PHVec(n, dtype=torch.float32)
env.reset()
bs = torch.empty((T, n, S), dtype=torch.float32, device="cuda"); bo = torch.empty((T, n, 4), dtype=torch.float32, device="cuda")
stats = torch.zeros(8, dtype=torch.float64, device="cuda")
K = np.array(w["K"])
for it in range(3):
    env.rollout(T, -K, actor=actor, auto_reset=True, replay=(bs, bo), stats=stats)
torch.cuda.synchronize()
out = (C.c_double * 16)()
print("rc", (L.lib().pime_debug_worker_prof if wl == "wt" else L.lib().pime_debug_worker_prof_ph)(out))
names = ["wait o_rdy", "E1a l1i", "wait d (P3 prev)", "dot", "E1b l1i", "wait l1b", "E2", "wait h_rdy", "E3a", "wait d (P2)", "E3b"]
tot = sum(out[:11])
for k, nme in enumerate(names): print(f"{nme:18s} {out[k]:9.1f} clk/pass  {100*out[k]/tot:5.1f}%")
print("total", tot)
