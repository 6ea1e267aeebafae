"""One PPOAgent.update on a synthetic [T, E] rollout (profiling aid: run under ncu for the launch list)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from highway_rope_ppo_b200.ppo.agent import PPOAgent

T, E, S, H = int(os.environ.get("T", 8)), 4096, 60, int(os.environ.get("H", 256))
epochs = int(os.environ.get("EPOCHS", 1))
torch.manual_seed(0)
np.random.seed(0)
agent = PPOAgent(S, 2, lr=3e-4, hidden_dim=H, batch_size=4096, epochs=epochs, device="cuda:0")


def fill():
    r = agent.memory.begin_rollout(T, E, S, 2)
    r["states"].normal_(0, 0.5)
    for t in range(T):
        agent.act(r["states"][t], out={"action": r["action"][t], "pre_tanh": r["pre_tanh"][t],
                                       "log_prob": r["log_prob"][t], "value": r["value"][t]})
    r["reward"].uniform_(0, 1)
    r["done"].copy_((torch.rand(T, E, device="cuda:0") < 0.03).to(torch.uint8))


for it in range(int(os.environ.get("ITERS", 2))):
    fill()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = agent.update(last_value=torch.zeros(E, device="cuda:0"))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"update {it}: {dt*1e3:.2f} ms for {T*E} samples x {epochs} epochs = {T*E*epochs/dt/1e6:.2f} M sample-passes/s loss {m['loss']:.4f}")
