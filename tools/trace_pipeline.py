"""Device timeline of the host-buffer pipeline (training/host_pipeline.py) in steady state, G groups (torch.profiler /
CUPTI): start offset, duration and stream of every device activity of a few group-steps.

    python tools/trace_pipeline.py [groups] [envs] > profiles/rNN_timeline_pipeline_gG.txt
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG  # noqa: E402
from highway_rope_ppo_b200.experiments.config import Condition  # noqa: E402
from highway_rope_ppo_b200.experiments.wrappers import make_vec_env  # noqa: E402
from highway_rope_ppo_b200.ppo.agent import PPOAgent  # noqa: E402
from highway_rope_ppo_b200.training.host_pipeline import HostBufferPipeline  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
Eg = E // G
over = {"observation": {"order": "shuffled"}}
agent = PPOAgent(60, 2, lr=3e-4, hidden_dim=256, batch_size=4096, epochs=8, device="cuda:0")
envs = [make_vec_env(Condition.SHUFFLED_ROPE, HIGHWAY_CONFIG, 4, over, num_envs=Eg, seed=42, env_id_base=g * Eg)
        for g in range(G)]
pipe = HostBufferPipeline(agent, envs)
pipe.reset(42)
for _ in range(30):
    for g in range(G):
        pipe.launch(g)
for g in range(G):
    pipe.wait(g)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        for g in range(G):
            pipe.launch(g)
    for g in range(G):
        pipe.wait(g)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
print(f"== {G} group(s) of {Eg} envs, 4 steps of every group: {len(evs)} device activities, span "
      f"{evs[-1].time_range.end - t0:.1f} us ({(evs[-1].time_range.end - t0) / 4:.1f} us per full step)")
for e in evs:
    print(f"  +{e.time_range.start - t0:8.1f} us  dur {e.time_range.end - e.time_range.start:7.1f}  {e.name[:80]}")
