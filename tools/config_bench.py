"""BASELINE.json configs[2]: Rank / Distance embedding, 30 observed vehicles, 16384 lock-step envs on one B200.
Policy + env steps/s (CUDA events, one pair per step, L2 flushed between steps) and PPO samples/s.

    python tools/config_bench.py > profiles/rNN_config2.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG
from highway_rope_ppo_b200.experiments.config import Condition
from highway_rope_ppo_b200.experiments.wrappers import make_vec_env
from highway_rope_ppo_b200.ppo.agent import PPOAgent
from highway_rope_ppo_b200.training.routine import rollout_and_update
from highway_rope_ppo_b200.utils.reproducibility import set_random_seeds

E, H, K, T = 16384, 256, 100, 8
over = {"observation": {"order": "shuffled", "vehicles_count": 30}}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print(f"{'condition':34s} {'state_dim':>9s} {'ms/step':>8s} {'policy env-steps/s':>19s} {'PPO samples/s':>14s}")
for cond, d, name in ((Condition.SHUFFLED_RANKPE, 16, "RankPE d=16"), (Condition.SHUFFLED_DISTPE, 16, "DistPE d=16"),
                      (Condition.SHUFFLED_DISTPE, 4, "DistPE d=4"), (Condition.SHUFFLED_ROPE, 4, "RoPE rotate_dim=4")):
    set_random_seeds(42)
    env = make_vec_env(cond, HIGHWAY_CONFIG, d, over, num_envs=E, seed=42, strict_d_embed=False)
    S = env.N * env.F_out
    agent = PPOAgent(S, 2, lr=3e-4, hidden_dim=H, batch_size=4096, epochs=8, device="cuda")
    obs = env.reset(42).view(E, S)
    out = {"action": torch.empty((E, 2), device="cuda"), "pre_tanh": torch.empty((E, 2), device="cuda"),
           "log_prob": torch.empty(E, device="cuda"), "value": torch.empty(E, device="cuda")}
    for _ in range(40):
        agent.act(obs, out=out); env.step(out["action"])
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_(); a.record(); agent.act(obs, out=out); env.step(out["action"]); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / K
    o = env.reset(42)
    _, o = rollout_and_update(env, agent, T, obs=o)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(2):
        m, o = rollout_and_update(env, agent, T, obs=o)
    e1.record(); torch.cuda.synchronize()
    sps = 2 * E * T / (e0.elapsed_time(e1) * 1e-3)
    print(f"{name + ', N=30, V=51':34s} {S:9d} {ms:8.4f} {E / ms * 1e3:19.3e} {sps:14.3e}")
    env.close(); agent.close()
