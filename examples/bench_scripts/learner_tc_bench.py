"""Per-minibatch time of the three PPO learner steps (Modular-256 actor + CriticAdv-256, fp32 semantics):
pime_ppo_step (rows + weight-gradient kernels), pime_ppo_grad_tc + pime_ppo_apply_grad (tcgen05), torch autograd (cuBLAS fp32)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pime_b200.rl as R

torch.backends.cuda.matmul.allow_tf32 = False
S, H, L = 4, 256, 1 << 21
torch.manual_seed(0)
a = R.AgentResidualIntegratorModularPPO(); a.init(H, S, 1, 1)
with torch.no_grad():
    a.act.net[-1].weight.normal_(0, 0.1)
state = torch.rand(L, S, device="cuda") * 10
action = torch.randn(L, device="cuda"); r_sum = torch.randn(L, device="cuda") * 30 - 50
logprob = -(torch.randn(L, device="cuda").pow(2) * 0.5 + a.act.a_std_log.item() + a.act.sqrt_2pi_log); adv = torch.randn(L, device="cuda")
data = (state, action, r_sum, logprob, adv)
f = R.FusedLearner(a.act, a.cri, S, H, a.device); f.load(a.act, a.cri)
flops_row = 2 * 3 * (132352 + 133121)      # forward + data gradient + weight gradient, 2 FLOP per weight each (first layers included)


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B in (1024, 2048, 4096, 16384, 65536, 131072, 524288):
    idx = torch.randint(L, size=(B,), device="cuda")
    row = {"B": B}
    if B <= 4096:
        row["simt_ms"] = timeit(lambda: f.step(data, idx, a), 20)
    row["tc_ms"] = timeit(lambda: f.step_tc(data, idx, a), 20 if B <= 65536 else 5)

    def autograd():
        oa, oc, ou, oe = a.ppo_objectives(state[idx], action[idx].unsqueeze(1), r_sum[idx], logprob[idx], adv[idx])
        a.optimizer.zero_grad(set_to_none=False); ou.backward(); a.optimizer.step()
    row["autograd_fp32_ms"] = timeit(autograd, 10 if B <= 65536 else 3)
    row["tc_tflops_algorithmic"] = B * flops_row / (row["tc_ms"] * 1e-3) / 1e12
    row["tc_tflops_issued_x3"] = 3 * row["tc_tflops_algorithmic"]
    print(row, flush=True)
