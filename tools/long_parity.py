"""Long state-injection parity runs (north_star: "1000 injected-state steps"): the same harness as
tests/test_env_gpu.py::_injected_parity, at larger env x step counts, for the BASELINE configs.

    python tools/long_parity.py [scale] > profiles/rNN_long_parity.txt      (scale multiplies the env counts; default 1)
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_env_gpu as T  # noqa: E402
from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG  # noqa: E402

SCALE = int(sys.argv[1]) if len(sys.argv) > 1 else 1
base = copy.deepcopy(HIGHWAY_CONFIG)
runs = [
    ("configs[0/1]: N=15 F=4 sorted, random actions", base, 64, 250, 11, T._random_actions, True),
    ("configs[0/1]: N=15 F=4 sorted, gentle actions (full 40-step episodes)", base, 64, 250, 12, T._gentle_actions, True),
    ("configs[1]: N=15 F=4 shuffled (injected permutation)", T._cfg(base, observation=dict(order="shuffled")), 64, 160, 13,
     T._random_actions, False),
    ("configs[2]: N=30 F=4 shuffled", T._cfg(base, observation=dict(vehicles_count=30, order="shuffled")), 64, 160, 14,
     T._gentle_actions, False),
]
total = 0
for name, cfg, E, steps, seed, fn, sorted_obs in runs:
    # observation tolerance = the state tolerance of y (1e-3 m) over its normalisation half-range (16 m); the unit
    # tests' 2e-5 is tighter than the state tolerances imply and is exceeded about once in 1e4 env-steps
    E *= SCALE
    s = T._injected_parity(cfg, E=E, steps=steps, seed=seed, action_fn=fn, sorted_obs=sorted_obs, obs_tol=6.5e-5)
    n = s["agree"] + s["flipped"] + s["neither"]
    total += n
    print(f"{name}: {E} envs x {steps} injected-state steps = {n} env-steps, NONE skipped: agree with the fp64 oracle "
          f"{s['agree']}, agree with the oracle after flipping marginal (< {T.MARGIN}) decisions {s['flipped']} "
          f"(of which {s['by_frames']} pile-ups resolved frame by frame against the kernel's per-frame trace), "
          f"neither {s['neither']}")
    print("   flipped decision kinds:", dict(s["kinds"]))
    print("   max abs error of the agreeing steps:", ", ".join(f"{k} {v:.3g}" for k, v in s["worst"].items()))
    for f in s["failures"]:
        print("   FAILURE (step, env, first difference):", f)
    assert s["neither"] == 0
# the fp64 validation instantiation of the same kernels: bit-exact discrete state, no margin rule
for name, cfg, E, steps, seed, fn, sorted_obs in runs[:2]:
    E *= SCALE
    s = T._injected_parity(cfg, E=E, steps=steps // 2, seed=seed + 100, action_fn=fn, sorted_obs=sorted_obs, real64=True)
    n = s["agree"] + s["flipped"] + s["neither"]
    print(f"fp64 kernel, {name}: {n} env-steps: discrete state bit-exact and continuous state within 1e-7 on {s['agree']}; "
          f"equal to the oracle after flipping exact ties (margin < {T.MARGIN64}: resting contacts) on {s['flipped']} "
          f"{dict(s['kinds'])}; neither {s['neither']}")
    print("   max abs error:", ", ".join(f"{k} {v:.3g}" for k, v in s["worst"].items()))
    assert s["neither"] == 0
for steps, fn, nm in ((45, T._gentle_actions, "gentle"), (45, T._random_actions, "random")):
    compared, _, worst = T._free_run(base, E=64 * SCALE, steps=steps, seed=40, action_fn=fn, real64=True)
    print(f"fp64 kernel, free running ({nm} actions, one injection, in-kernel respawn, {64 * SCALE} envs x {steps} steps): "
          f"{compared} env-steps compared, every one bit-exact in the discrete state and within 1e-7 in the continuous "
          f"state; {T._free_run.ended} windows ended early at an exact tie or after a crawling / reversing vehicle "
          "(unstable in the oracle itself); max abs error:", ", ".join(f"{k} {v:.3g}" for k, v in worst.items()))
print(f"total fp32 env-steps checked with the either-branch rule: {total}; skipped: 0")
