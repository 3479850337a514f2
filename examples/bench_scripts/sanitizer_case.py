"""Smallest cases for compute-sanitizer (racecheck / memcheck): the fused rollout (tcgen05 engine, H = 32) of both plants,
the fidelity-mode rollout, the 4-env step kernel and one tensor-core learner step (H = 128, 300 rows)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pime_b200.vec as V
import pime_b200.rl as R

rng = np.random.default_rng(0)


def sd_mod(H, S):
    out = {}
    for name, o, i in [("other_net.0", H, S - 1), ("other_net.2", H // 2, H), ("integrator_net.0", H, 1), ("integrator_net.2", H // 2, H),
                       ("net.0", H, H), ("net.2", 1, H)]:
        b = 1 / np.sqrt(i)
        out[name + ".weight"] = rng.uniform(-b, b, (o, i)).astype(np.float32)
        out[name + ".bias"] = rng.uniform(-b, b, o).astype(np.float32)
    return out


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "rollout"):
    env = V.WaterTankVec(300, dtype=torch.float32, seed=1); env.reset()
    env.rollout(6, -np.array([0., .4, -.4, 0.]), actor=V.ActorPack("modular", 4, 32, 1).update(sd_mod(32, 4)), auto_reset=True, replay=True)
    ph = V.PHVec(300, dtype=torch.float32, seed=1); ph.reset()
    ph.rollout(6, -np.array([-.02, .02, .035]), actor=V.ActorPack("modular", 3, 32, 1).update(sd_mod(32, 3)), auto_reset=True, replay=True)
    env.rollout(3, -np.array([0., .4, -.4, 0.]), actor=V.ActorPack("modular", 4, 32, 1, precision="fp32").update(sd_mod(32, 4)), replay=True)
    env.step(torch.zeros(300, device="cuda"))
if which in ("all", "learner"):
    torch.manual_seed(0)
    a = R.AgentResidualIntegratorModularPPO(); a.init(128, 4, 1, 1)
    L = 1000
    data = (torch.rand(L, 4, device="cuda"), torch.randn(L, device="cuda"), torch.randn(L, device="cuda"), -torch.rand(L, device="cuda"), torch.randn(L, device="cuda"))
    f = R.FusedLearner(a.act, a.cri, 4, 128, a.device); f.load(a.act, a.cri)
    f.step_tc(data, torch.randint(L, size=(300,), device="cuda"), a)
    f.step(data, torch.randint(L, size=(64,), device="cuda"), a)
torch.cuda.synchronize()
print("sanitizer case ok")
