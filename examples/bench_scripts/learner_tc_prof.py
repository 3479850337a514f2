"""Three tensor-core learner steps at B = 131072 (Modular-256 + CriticAdv-256) for ncu: launch list / --set full captures."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pime_b200.rl as R

S, H, L, B = 4, 256, 1 << 21, int(os.environ.get("PIME_TC_B", 1 << 17))
torch.manual_seed(0)
a = R.AgentResidualIntegratorModularPPO(); a.init(H, S, 1, 1)
with torch.no_grad():
    a.act.net[-1].weight.normal_(0, 0.1)
state = torch.rand(L, S, device="cuda") * 10
action = torch.randn(L, device="cuda"); r_sum = torch.randn(L, device="cuda") * 30 - 50
logprob = -(torch.randn(L, device="cuda").pow(2) * 0.5 + a.act.a_std_log.item() + a.act.sqrt_2pi_log); adv = torch.randn(L, device="cuda")
data = (state, action, r_sum, logprob, adv)
f = R.FusedLearner(a.act, a.cri, S, H, a.device); f.load(a.act, a.cri)
for it in range(3):
    idx = torch.randint(L, size=(B,), device="cuda")
    f.step_tc(data, idx, a)
torch.cuda.synchronize()
print("ok", float(f.losses(2, 1)[0, 0]))
