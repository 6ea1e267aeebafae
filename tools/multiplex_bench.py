"""Runs per hour of the reference sweep on ONE GPU: R multiplexed experiments (experiments/multiplex.py) against the
same experiments run one after the other (experiments/runner.py) -- the replacement of utils/device_pool.py:45-72's
OVERSUB time-sharing.

    python tools/multiplex_bench.py [--runs 16] [--episodes 60] [--steps-per-update 512]

The experiments are the first R of the reference grid order restricted to one seed per configuration (E = 1, batch 32 /
64, hidden 128 / 256 / 384), shortened to --episodes episodes so that the comparison finishes in minutes."""
import argparse
import json
import logging
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from highway_rope_ppo_b200.config.base_config import HIGHWAY_CONFIG  # noqa: E402
from highway_rope_ppo_b200.experiments.multiplex import MultiplexedRunner  # noqa: E402
from highway_rope_ppo_b200.experiments.runner import ExperimentRunner  # noqa: E402
from highway_rope_ppo_b200.experiments.sweep import define_experiments  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=16)
    ap.add_argument("--episodes", type=int, default=60)
    ap.add_argument("--steps-per-update", type=int, default=512)
    ap.add_argument("--stride", type=int, default=7, help="take every stride-th experiment of the grid (mixes conditions)")
    ap.add_argument("--skip-sequential", action="store_true")
    a = ap.parse_args()
    logging.disable(logging.CRITICAL)
    grid = define_experiments(num_seeds=1)
    exps = grid[::a.stride][:a.runs]
    for e in exps:
        e.max_episodes = a.episodes
        e.hp.steps_per_update = a.steps_per_update
        e.extra = {"eval_interval": 20, "log_interval": 20}
    out = {"runs": len(exps), "episodes": a.episodes, "steps_per_update": a.steps_per_update,
           "conditions": sorted({e.condition.name for e in exps})}
    with tempfile.TemporaryDirectory() as tmp:
        # warm-up: kernels loaded, allocator pools grown
        ExperimentRunner(HIGHWAY_CONFIG, artifacts_dir=tmp).launch(exps[0])
        if not a.skip_sequential:
            torch.cuda.synchronize()
            t0 = time.time()
            seq = [ExperimentRunner(HIGHWAY_CONFIG, artifacts_dir=tmp).launch(e) for e in exps]
            torch.cuda.synchronize()
            dt = time.time() - t0
            assert all(r["status"] == "COMPLETED" for r in seq), [r.get("error_message") for r in seq]
            out["sequential"] = {"seconds": dt, "runs_per_hour": 3600.0 * len(exps) / dt}
        mux = MultiplexedRunner(HIGHWAY_CONFIG, artifacts_dir=tmp, max_concurrent=len(exps))
        torch.cuda.synchronize()
        t0 = time.time()
        res = mux.launch_many(exps)
        torch.cuda.synchronize()
        dt = time.time() - t0
        assert all(r["status"] == "COMPLETED" for r in res), [r.get("error_message") for r in res]
        out["multiplexed"] = {"seconds": dt, "runs_per_hour": 3600.0 * len(exps) / dt, "ticks": mux.ticks,
                              "env_launches": mux.env_launches, "env_requests": mux.env_requests}
        if not a.skip_sequential:
            out["speedup"] = out["sequential"]["seconds"] / dt
            out["identical_rewards"] = all(x["rewards"] == y["rewards"] for x, y in zip(seq, res))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
