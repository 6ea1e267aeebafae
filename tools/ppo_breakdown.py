"""Where one PPO iteration goes: rollout, update, and the replay of one minibatch graph (CUDA events)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG
from highway_rope_ppo_b200.experiments.config import Condition
from highway_rope_ppo_b200.experiments.wrappers import make_vec_env
from highway_rope_ppo_b200.ppo.agent import PPOAgent
from highway_rope_ppo_b200.training.routine import collect_rollout

E, T, H = int(os.environ.get("E", 4096)), int(os.environ.get("T", 32)), int(os.environ.get("H", 256))
env = make_vec_env(Condition.SHUFFLED_ROPE, HIGHWAY_CONFIG, 4, {"observation": {"order": "shuffled"}}, num_envs=E, seed=42)
agent = PPOAgent(60, 2, lr=3e-4, hidden_dim=H, batch_size=4096, epochs=8, device="cuda")
obs = env.reset(42)


def timed(fn, n=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - w0) * 1e3 / n, r


for it in range(3):
    t_roll, w_roll, r = timed(lambda: collect_rollout(env, agent, T, obs))
    last_obs = r["states"][T].clone()
    _, _, last_v = agent.actor_critic.forward(last_obs)
    t_upd, w_upd, _ = timed(lambda: agent.update(last_value=last_v.view(-1)))
    obs = last_obs
    print(f"iter {it}: rollout {t_roll:.2f} ms (wall {w_roll:.2f}), update {t_upd:.2f} ms (wall {w_upd:.2f})")
st = agent._graph_state
g = st["graphs"][0]
t_g, w_g, _ = timed(g.replay, 200)
print(f"one minibatch graph replay: {t_g * 1e3:.1f} us device, {w_g * 1e3:.1f} us wall; {len(st['graphs'])} graphs")
# eager minibatch step for comparison
idx = st["perm"][:4096]
t_e, w_e, _ = timed(lambda: agent._minibatch_step(st["buf"], idx, 4096, 1), 100)
print(f"one eager minibatch step: {t_e * 1e3:.1f} us device, {w_e * 1e3:.1f} us wall")
# forward only
x = st["buf"]["states"][:4096]
t_f, _, _ = timed(lambda: agent.actor_critic.forward(x), 200)
print(f"forward (4 kernels + allocs): {t_f * 1e3:.1f} us")
out = None
t_a, _, _ = timed(lambda: agent.act(x), 200)
print(f"act: {t_a * 1e3:.1f} us")
