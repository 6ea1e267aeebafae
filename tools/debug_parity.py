"""GPU-box debugging aid: rerun an injected-state parity loop and print the worst offenders with context."""
import copy
import sys
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import highway as oh
from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG
from highway_rope_ppo_b200.envs.highway_vec import HighwayVecEnv


def stack(envs):
    sts = [e.get_state() for e in envs]
    out = {k: np.stack([s[k] for s in sts]) for k in oh.STATE_F64 + oh.STATE_I32}
    out["time"] = np.array([s["time"] for s in sts])
    return out


def run(cfg, E, steps, seed, gentle, meta=False, label=""):
    env = HighwayVecEnv(cfg, E, device="cuda:0", autoreset=False)
    oracles = [oh.OracleEnv(cfg) for _ in range(E)]
    for e, o in enumerate(oracles):
        o.reset(seed, env_id=e, episode=0)
    rng = np.random.default_rng(seed)
    shown = 0
    for t in range(steps):
        st = stack(oracles)
        env.set_state(st)
        a = rng.uniform(-1, 1, (E, 2))
        if gentle:
            a[:, 1] *= 0.08
        if meta:
            a[:, 0] = rng.integers(0, 5, E); a[:, 1] = 0
        a = a.astype(np.float32)
        env.step(torch.from_numpy(a).cuda())
        got = env.get_state()
        for e, o in enumerate(oracles):
            r, te, tr = o.step(a[e])
            ref = o.get_state()
            m = o.min_margin()
            for k in ("x", "y", "speed", "heading"):
                err = np.abs(got[k][e] - ref[k])
                j = int(err.argmax())
                if err[j] > 2e-3 and shown < 12 and m >= 1e-3:
                    shown += 1
                    print(f"[{label}] t={t} e={e} field={k} veh={j} err={err[j]:.5f} margin={m:.2e} "
                          f"before: x={st['x'][e][j]:.3f} y={st['y'][e][j]:.3f} v={st['speed'][e][j]:.3f} h={st['heading'][e][j]:.4f} "
                          f"crashed={st['crashed'][e][j]} imp={st['has_impact'][e][j]} ({st['impact_x'][e][j]:.4f},{st['impact_y'][e][j]:.4f}) "
                          f"lane={st['lane'][e][j]}->{st['target_lane'][e][j]} | ref after: x={ref['x'][j]:.4f} y={ref['y'][j]:.4f} v={ref['speed'][j]:.4f} "
                          f"h={ref['heading'][j]:.5f} crashed={ref['crashed'][j]} imp={ref['has_impact'][j]} ({ref['impact_x'][j]:.4f},{ref['impact_y'][j]:.4f}) | "
                          f"gpu after: x={got['x'][e][j]:.4f} y={got['y'][e][j]:.4f} v={got['speed'][e][j]:.4f} h={got['heading'][e][j]:.5f} "
                          f"crashed={got['crashed'][e][j]} imp={got['has_impact'][e][j]} ({got['impact_x'][e][j]:.4f},{got['impact_y'][e][j]:.4f}) a={a[e]}")
                    break
            if te or tr:
                o.reset(seed, env_id=e, episode=1 + t)
    env.close()


def c(**over):
    cfg = copy.deepcopy(HIGHWAY_CONFIG)
    for k, v in over.items():
        if isinstance(v, dict):
            cfg[k].update(v)
        else:
            cfg[k] = v
    return cfg


run(c(lanes_count=2, vehicles_count=12, vehicles_density=3, observation=dict(vehicles_count=5, see_behind=True)), 32, 30, 4, False, label="dense")
m = c(); m["action"] = {"type": "DiscreteMetaAction"}
run(m, 24, 40, 5, False, meta=True, label="meta")
