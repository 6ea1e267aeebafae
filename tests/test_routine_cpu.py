"""The reference training loop as a request coroutine (training/routine.py): its protocol, driven by pure-Python
stand-ins for the env and the agent -- no GPU, no library.  The two drivers (``drive`` with one env / agent, the
multiplexer with R of them) answer exactly these requests, so the sequence pinned here is the contract between them.

Reference control flow: ``training/routine.py:61-297`` (initial evaluation of 5 episodes, collect until
``steps_per_update`` or ``max_episodes``, evaluation every ``eval_interval`` episodes, bootstrap value only when the
rollout was cut inside an episode, best / solved checkpoints, metrics JSON + summary CSV at the end)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Env:
    """Episodes of fixed length 7 (truncated), reward 1 per step, observation = [episode seed, t]."""

    def __init__(self):
        self.calls, self.t, self.seed = [], 0, None

    def reset(self, *, seed=None, options=None):
        self.calls.append(("reset", seed))
        self.seed, self.t = seed, 0
        return np.array([[seed, 0.0]], dtype=np.float32), {}

    def step(self, action):
        self.t += 1
        self.calls.append(("step", self.seed, self.t))
        return np.array([[self.seed, self.t]], dtype=np.float32), 1.0, False, self.t >= 7, {}


class _Memory:
    def __init__(self):
        self.rows = []

    def store(self, *row):
        self.rows.append(row)


class _AC:
    def __init__(self, log):
        self.log = log

    def forward(self, flat):
        import torch

        self.log.append(("value", float(flat[0]), float(flat[1])))
        return None, None, torch.tensor([0.25])


class _Agent:
    def __init__(self):
        self.log, self.memory = [], _Memory()
        self.actor_critic = _AC(self.log)

    def select_action(self, state, deterministic=False):
        self.log.append(("act", deterministic))
        return np.zeros(2, np.float32), np.zeros(2, np.float32), None if deterministic else -1.0, np.float32(0.5)

    def update(self, last_value=0.0):
        self.log.append(("update", last_value, len(self.memory.rows)))
        self.memory.rows = []
        return {"loss": 0.0, "policy_loss": 0.0, "value_loss": 0.0, "entropy": 0.0, "clip_fraction": 0.0, "approx_kl": 0.0,
                "explained_variance": 0.0}

    def save(self, path):
        self.log.append(("save", os.path.basename(path)))


def test_training_coroutine_protocol(tmp_path):
    from highway_rope_ppo_b200.training.routine import train_with_experiment_name

    env, agent = _Env(), _Agent()
    rewards, avg_rewards, history = train_with_experiment_name(
        env, agent, max_episodes=6, target_reward=1e9, log_interval=2, eval_interval=3, steps_per_update=10,
        experiment_name="proto", exp_seed=100, artifacts_dir=str(tmp_path))
    resets = [c[1] for c in env.calls if c[0] == "reset"]
    # initial evaluation: 5 deterministic episodes seeded exp_seed + 1000 + i; then training episodes exp_seed + n,
    # with an evaluation (same 5 seeds) after episodes 3 and 6
    ev = [1100, 1101, 1102, 1103, 1104]
    assert resets == ev + [101, 102, 103] + ev + [104, 105, 106] + ev
    acts = [a for a in agent.log if a[0] == "act"]
    assert sum(1 for a in acts if a[1]) == 3 * 5 * 7 and sum(1 for a in acts if not a[1]) == 30
    # steps_per_update = 10 cuts episodes 2, 4 and 6 after their third step (a cut episode is abandoned: the next
    # collection starts with a fresh reset, as in the reference): the bootstrap value is asked for exactly then, on the
    # last observation, and every update sees 10 stored samples
    updates = [a for a in agent.log if a[0] == "update"]
    values = [a for a in agent.log if a[0] == "value"]
    assert [u[2] for u in updates] == [10, 10, 10]
    assert [u[1] for u in updates] == [0.25, 0.25, 0.25]
    assert values == [("value", 102.0, 3.0), ("value", 104.0, 3.0), ("value", 106.0, 3.0)]
    # results: initial eval + two evals; every eval episode returns 7
    assert rewards == [7.0, 7.0, 7.0] and avg_rewards == [7.0, 7.0, 7.0]
    assert history["eval_episode_numbers"] == [0, 3, 6] and len(history["policy_updates"]) == 3
    assert history["episode_rewards"] == [7.0, 3.0, 7.0, 3.0, 7.0, 3.0]
    assert [a for a in agent.log if a[0] == "save"] == [("save", "ppo_highway_best_proto.pth")]
    data = json.load(open(tmp_path / "training_metrics_proto.json"))
    assert data["experiment_name"] == "proto" and len(data["policy_updates"]) == 3
    assert (tmp_path / "summary_proto.csv").read_text().splitlines()[1].startswith("proto,7.0000,7.0000,30,")


def test_training_coroutine_yields_only_the_six_request_kinds(tmp_path):
    from highway_rope_ppo_b200.training.routine import training_coroutine

    co = training_coroutine(_Memory(), max_episodes=2, eval_interval=1, steps_per_update=5, experiment_name="kinds",
                            exp_seed=1, artifacts_dir=str(tmp_path))
    kinds, t = set(), 0
    try:
        req = next(co)
        while True:
            kinds.add(req[0])
            if req[0] == "reset":
                resp, t = np.zeros((1, 2), np.float32), 0
            elif req[0] == "act":
                resp = (np.zeros(2, np.float32), np.zeros(2, np.float32), -1.0, np.float32(0.0))
            elif req[0] == "step":
                t += 1
                resp = (np.zeros((1, 2), np.float32), 1.0, t >= 3, False)
            elif req[0] == "value":
                resp = 0.0
            elif req[0] == "update":
                resp = {"loss": 0.0}
            else:
                resp = None
            req = co.send(resp)
    except StopIteration as done:
        rewards, avg, hist = done.value
    assert kinds <= {"reset", "act", "step", "value", "update", "save"} and {"reset", "act", "step", "update"} <= kinds
    assert len(rewards) == 3
