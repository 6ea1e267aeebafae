"""Dump one env-step of the state-injection parity run for offline analysis: the injected state, the action, the
kernel's (fp32 and fp64 instantiation) and the oracle's outcome.

    python tools/debug_parity.py <seed> <E> <gentle|random> <t> <env> [<t> <env> ...]  ->  gpurun_out/debug_case_*.npz
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_env_gpu as T  # noqa: E402
from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG  # noqa: E402
from oracle import highway as oh  # noqa: E402

seed, E, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
cases = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(4, len(sys.argv), 2)]
fn = T._gentle_actions if kind == "gentle" else T._random_actions
cfg = copy.deepcopy(HIGHWAY_CONFIG)
oracles = [oh.OracleEnv(cfg) for _ in range(E)]
for e, o in enumerate(oracles):
    o.reset(seed, env_id=e, episode=0)
rng = np.random.default_rng(seed)
env32 = T._vec(cfg, 1, autoreset=False)
trace32 = env32.enable_trace()
env64 = T._vec(cfg, 1, autoreset=False, real64=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for t in range(max(c[0] for c in cases) + 1):
    actions = fn(rng, E, t).astype(np.float32)
    for e, o in enumerate(oracles):
        if (t, e) in cases:
            st = o.get_state()
            out = {f"st0_{k}": st[k] for k in oh.STATE_F64 + oh.STATE_I32}
            out["st0_time"], out["st0_steps"], out["action"] = st["time"], st["steps"], actions[e]
            for name, env in (("k32", env32), ("k64", env64)):
                full = {k: st[k][None] for k in oh.STATE_F64 + oh.STATE_I32}
                full["time"] = np.array([st["time"]])
                env.set_state(full)
                env.step(torch.from_numpy(actions[e][None]).cuda())
                got = env.get_state()
                for k in oh.STATE_F64 + oh.STATE_I32:
                    out[f"{name}_{k}"] = got[k][0]
                if name == "k32":
                    out["k32_trace"] = trace32[0].cpu().numpy()
            np.savez(os.path.join(ROOT, "gpurun_out", f"debug_case_{kind}_{seed}_{t}_{e}.npz"), **out)
        r, te, tr = o.step(actions[e])
        if te or tr:
            o.reset(seed, env_id=e, episode=1 + t)
print("done")
