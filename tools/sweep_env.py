"""Env-step-only throughput sweep (BASELINE.json configs[4]) and the N=30 / Rank / Dist shapes of configs[2].

    python tools/sweep_env.py > profiles/rNN_env_sweep.txt

Random actions resident in HBM, autoreset on, 40 warm-up steps (one full episode), CUDA-event timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG
from highway_rope_ppo_b200.experiments.config import Condition
from highway_rope_ppo_b200.experiments.wrappers import make_vec_env


def run(cond, d, over, E, steps=120, warm=40, strict=True):
    env = make_vec_env(cond, HIGHWAY_CONFIG, d, over, num_envs=E, seed=1, strict_d_embed=strict)
    env.reset(1)
    a = torch.rand((E, 2), device="cuda") * 2 - 1
    for _ in range(warm):
        env.step(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        env.step(a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    env.close()
    return ms


print(f"{'config':46s} {'envs':>8s} {'ms/step':>9s} {'env-steps/s':>13s}")
sh = {"observation": {"order": "shuffled"}}
for E in (256, 1024, 4096, 8192, 16384, 65536, 262144):
    ms = run(Condition.SHUFFLED_ROPE, 4, sh, E, steps=120 if E <= 16384 else 30)
    print(f"{'shuffled + RoPE(4), N=15, F=4, V=51':46s} {E:8d} {ms:9.4f} {E / ms * 1e3:13.3e}")
n30 = {"observation": {"order": "shuffled", "vehicles_count": 30}}
for cond, d, name in ((Condition.SHUFFLED_RANKPE, 16, "RankPE d=16"), (Condition.SHUFFLED_DISTPE, 16, "DistPE d=16"),
                      (Condition.SHUFFLED_DISTPE, 4, "DistPE d=4"), (Condition.SORTED, None, "sorted, no embedding")):
    ms = run(cond, d, n30 if cond is not Condition.SORTED else {"observation": {"vehicles_count": 30}}, 16384,
             steps=60, strict=False)
    print(f"{name + ', N=30, F=4, V=51':46s} {16384:8d} {ms:9.4f} {16384 / ms * 1e3:13.3e}")
