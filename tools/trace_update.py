"""Kernel timeline of one policy step and one PPO optimizer step (torch.profiler / CUPTI): start offset and duration
of every kernel, in launch order, for the warm steady state.  Shows where the time between kernels goes, which the
serialised ncu launch list cannot.

    python tools/trace_update.py [hidden] [minibatch] > profiles/rNN_timeline.txt
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200.ppo.agent import PPOAgent  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
S, A, n = 60, 2, 4 * B
torch.manual_seed(0)
agent = PPOAgent(S, A, lr=3e-4, hidden_dim=H, batch_size=B, epochs=2, device="cuda:0")
dev = "cuda:0"
x = torch.randn(n, S, device=dev) * 0.5
out = agent.act(x[:B])
flat = {"states": x, "pre_tanh": torch.randn(n, A, device=dev), "log_prob": -torch.rand(n, device=dev) - 1,
        "adv": torch.randn(n, device=dev), "ret": torch.rand(n, device=dev)}
perm = torch.randperm(n, device=dev)


def show(prof, title):
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print(title, ": no device events")
        return
    t0 = evs[0].time_range.start
    print(f"== {title}: {len(evs)} device activities, span {evs[-1].time_range.end - t0:.1f} us")
    prev_end = t0
    for e in evs:
        gap = e.time_range.start - prev_end
        print(f"  +{e.time_range.start - t0:8.1f} us  dur {e.time_range.end - e.time_range.start:7.1f}  gap {gap:6.1f}  {e.name[:90]}")
        prev_end = max(prev_end, e.time_range.end)


for _ in range(3):
    agent.act(x[:B], out=out)
    agent._minibatch_step(flat, perm[:B], B, 1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        agent.act(x[:B], out=out)
    torch.cuda.synchronize()
show(prof, "3 x policy forward (hrp_ppo_act_sample), eager")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    agent._minibatch_step(flat, perm[:B], B, 1)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.replay()
    g.replay()
    torch.cuda.synchronize()
show(prof, "2 x optimizer step (CUDA graph replay)")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) * 1e3 / 50:.1f} us per optimizer step")
e0.record()
for _ in range(50):
    agent.act(x[:B], out=out)
e1.record()
torch.cuda.synchronize()
print(f"policy forward: {e0.elapsed_time(e1) * 1e3 / 50:.1f} us per call")
