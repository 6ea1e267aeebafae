"""Drop-in surface on the GPU: make_env / wrappers / PPOAgent / training loop protocol."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_make_env_protocol(highway_config):
    from highway_rope_ppo_b200.envs.spaces import is_box
    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_env

    env = make_env(Condition.SORTED, highway_config)
    assert is_box(env.observation_space) and env.observation_space.shape == (15, 4)
    assert env.action_space.shape == (2,)
    obs, info = env.reset(seed=43)
    assert obs.shape == (15, 4) and obs.dtype == np.float32 and isinstance(info, dict)
    obs2, _ = env.reset(seed=43)
    assert np.array_equal(obs, obs2)  # same seed, same episode
    total, steps, done = 0.0, 0, False
    while not done:
        obs, r, term, trunc, info = env.step(np.array([0.1, 0.0], dtype=np.float32))
        assert isinstance(r, float) and isinstance(term, bool) and isinstance(trunc, bool)
        done = term or trunc
        total += r
        steps += 1
    assert 1 <= steps <= 40 and 0.0 <= total <= 40.0
    env.close()


@pytest.mark.parametrize("cond,d,shape", [("SHUFFLED_ROPE", 4, (15, 4)), ("SHUFFLED_DISTPE", 4, (15, 8)),
                                          ("SHUFFLED_RANKPE", 16, (15, 20)), ("SHUFFLED", None, (15, 4))])
def test_make_env_conditions(highway_config, cond, d, shape):
    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_env

    env = make_env(Condition[cond], highway_config, d_embed=d)
    if hasattr(env, "to"):
        env = env.to(torch.device("cuda:0"))
    assert env.observation_space.shape == shape
    obs, _ = env.reset(seed=1)
    assert obs.shape == shape and obs.dtype == np.float32
    obs, r, te, tr, _ = env.step(np.zeros(2, dtype=np.float32))
    assert obs.shape == shape
    # stock config says "sorted": setdefault keeps it
    assert env.unwrapped.config["observation"]["order"] == "sorted"
    env.close()


def test_make_env_errors(highway_config):
    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_env

    with pytest.raises(ValueError):
        make_env(Condition.SHUFFLED_ROPE, highway_config, d_embed=3)
    with pytest.raises(ValueError):
        make_env(Condition.SHUFFLED_ROPE, highway_config, d_embed=16)  # > F = 4
    with pytest.raises(ValueError):
        make_env(Condition.SHUFFLED_DISTPE, highway_config, d_embed=6)
    with pytest.raises(ValueError):
        make_env(Condition.SHUFFLED_RANKPE, highway_config, d_embed=None)
    cfg = copy.deepcopy(highway_config)
    del cfg["observation"]["order"]
    env = make_env(Condition.SHUFFLED, cfg)
    assert env.unwrapped.config["observation"]["order"] == "shuffled"
    env.close()


def test_reference_training_loop_runs(highway_config, tmp_path):
    """training/routine.py semantics end to end on a tiny budget; outputs keep the reference schemas."""
    import json

    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_env
    from highway_rope_ppo_b200.ppo.agent import PPOAgent
    from highway_rope_ppo_b200.training.routine import train_with_experiment_name
    from highway_rope_ppo_b200.utils.reproducibility import set_random_seeds

    set_random_seeds(42)
    env = make_env(Condition.SHUFFLED_ROPE, highway_config, d_embed=4)
    agent = PPOAgent(60, 2, lr=3e-4, hidden_dim=64, batch_size=32, epochs=2, device="cuda:0")
    rewards, avg, hist = train_with_experiment_name(env, agent, max_episodes=6, eval_interval=3, log_interval=2,
                                                    steps_per_update=64, experiment_name="t", exp_seed=42,
                                                    artifacts_dir=str(tmp_path))
    assert len(rewards) == 3 and len(hist["policy_updates"]) >= 1
    assert set(hist["policy_updates"][0]) >= {"loss", "policy_loss", "value_loss", "entropy", "clip_fraction",
                                              "approx_kl", "explained_variance", "episode", "steps", "time"}
    assert json.load(open(tmp_path / "training_metrics_t.json"))["experiment_name"] == "t"
    assert (tmp_path / "summary_t.csv").read_text().startswith("experiment,final_reward,max_reward,steps")
    assert os.path.exists(tmp_path / "checkpoints" / "ppo_highway_best_t.pth")
    env.close()


def test_vectorized_training_improves_reward(highway_config):
    """The batched loop learns: mean step reward after 12 iterations beats the initial policy's."""
    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_vec_env
    from highway_rope_ppo_b200.ppo.agent import PPOAgent
    from highway_rope_ppo_b200.training.routine import collect_rollout, rollout_and_update
    from highway_rope_ppo_b200.utils.reproducibility import set_random_seeds

    set_random_seeds(0)
    env = make_vec_env(Condition.SHUFFLED_ROPE, highway_config, 4, {"observation": {"order": "shuffled"}},
                       num_envs=512, seed=0)
    agent = PPOAgent(60, 2, lr=3e-4, hidden_dim=128, batch_size=1024, epochs=4, device="cuda:0")
    obs = env.reset(0)
    first = last = None
    for it in range(12):
        r = collect_rollout(env, agent, 32, obs)
        mean_r = float(r["reward"].mean())
        # survival matters more than step reward: count env-steps that did not end in a crash
        first = mean_r if first is None else first
        last = mean_r
        obs = r["states"][32].clone()
        _, _, v = agent.actor_critic.forward(obs)
        m = agent.update(last_value=v.view(-1))
        assert np.isfinite(m["loss"])
    print("mean step reward first/last", first, last)
    assert last > first
    env.close()


def test_experiment_runner_result_schema(highway_config, tmp_path):
    """experiments/runner.py:46-155 re-host: COMPLETED result with the reference's keys; failures are reported."""
    from highway_rope_ppo_b200.experiments.config import Condition, ConditionHP, Experiment
    from highway_rope_ppo_b200.experiments.runner import ExperimentRunner, experiment_name

    hp = ConditionHP(lr=3e-4, hidden_dim=64, batch_size=32, epochs=2, d_embed=4, steps_per_update=48)
    name = experiment_name("SORTED", hp, 42)
    assert name == "sorted_lr0.0003_hidden_dim64_clip_eps0.2_entropy_coef0.005_epochs2_batch_size32_d_embed4_seed42"
    runner = ExperimentRunner(highway_config, artifacts_dir=str(tmp_path))
    res = runner.launch(Experiment(name, Condition.SORTED, hp, seed=42, max_episodes=3,
                                   extra={"eval_interval": 3, "log_interval": 1}))
    assert res["status"] == "COMPLETED", res.get("error_traceback")
    assert set(res) >= {"experiment_name", "status", "rewards", "avg_rewards", "metrics_history", "duration_seconds"}
    assert len(res["rewards"]) == 2 and res["metrics_history"]["experiment_name"] == name
    # the artifacts are what the reference's offline tools read (tests/golden/result_schema.json, extracted from
    # results.py / training/routine.py by tools/gen_golden.py): name regex, metrics JSON keys, summary CSV header
    import json
    import re

    schema = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "result_schema.json")))
    data = json.load(open(tmp_path / f"training_metrics_{name}.json"))
    assert list(data) == schema["metrics_keys"]
    meta = re.match(schema["exp_rx"], data["experiment_name"]).groupdict()
    assert meta["prefix"] == "sorted" and meta["pe_type"] is None and int(meta["hidden_dim"]) == 64 and int(meta["seed"]) == 42
    assert len(data["eval_episode_numbers"]) == len(data["avg_eval_rewards"]) >= 1
    lines = (tmp_path / f"summary_{name}.csv").read_text().splitlines()
    assert lines[0] == schema["summary_header"] and lines[1].split(",")[0] == name and len(lines[1].split(",")) == 6
    assert lines[1].split(",")[4].endswith(f"ppo_highway_best_{name}.pth")
    bad = runner.launch(Experiment("bad", Condition.SHUFFLED_ROPE, ConditionHP(d_embed=16), seed=1, max_episodes=1))
    assert bad["status"] == "FAILED" and "rotate_dim" in bad["error_message"] and "error_traceback" in bad


def test_meta_action_env_through_make_env(highway_config):
    """DiscreteMetaAction ego (north_star's MDPVehicle controller) through the gymnasium-protocol env."""
    import copy

    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_env

    cfg = copy.deepcopy(highway_config)
    cfg["action"] = {"type": "DiscreteMetaAction"}
    env = make_env(Condition.SORTED, cfg)
    assert env.action_space.n == 5
    obs, _ = env.reset(seed=3)
    total = 0.0
    for t in range(40):
        obs, r, te, tr, _ = env.step(1 if t % 7 else 3)  # IDLE, occasionally FASTER
        total += r
        if te or tr:
            break
    assert obs.shape == (15, 4) and total > 0
    env.close()


def _sweep_like_experiments(n, max_episodes, steps_per_update):
    """n single-env experiments shaped like the reference sweep (main.py:42-88): every condition, hidden 128 / 256,
    batch 32 / 64, several seeds (RankPE twice with different seeds = different tables = separate env handles)."""
    from highway_rope_ppo_b200.experiments.config import Condition, ConditionHP, Experiment

    conds = [Condition.SORTED, Condition.SHUFFLED, Condition.SHUFFLED_ROPE, Condition.SHUFFLED_DISTPE,
             Condition.SHUFFLED_RANKPE]
    exps = []
    for i in range(n):
        cond = conds[i % len(conds)]
        hp = ConditionHP(lr=(1e-4, 3e-4)[i % 2], hidden_dim=(128, 256)[(i // 2) % 2], batch_size=(32, 64)[(i // 3) % 2],
                         epochs=(2, 3)[(i // 5) % 2], d_embed=4, steps_per_update=steps_per_update)
        seed = 42 + 1000 * (i % 3)
        over = {} if cond is Condition.SORTED else {"observation": {"order": "shuffled"}}
        exps.append(Experiment(Experiment.make_name(cond, hp, seed) + f"_i{i}", cond, hp, seed=seed,
                               max_episodes=max_episodes, env_config_overrides=over,
                               extra={"eval_interval": 4, "log_interval": 2}))
    return exps


def test_multiplexed_experiments_reproduce_sequential_runs(highway_config, tmp_path):
    """R = 16 experiments sharing env handles, one launch / one synchronisation per tick, updates
    overlapped on their own streams (experiments/multiplex.py) end with exactly the parameters, Adam state, rewards
    and update metrics of 16 sequential ExperimentRunner.launch calls (the replacement of utils/device_pool.py:45-72)."""
    from highway_rope_ppo_b200.experiments.multiplex import MultiplexedRunner
    from highway_rope_ppo_b200.experiments.runner import ExperimentRunner

    exps = _sweep_like_experiments(16, max_episodes=8, steps_per_update=96)

    def capture(cls, **kw):
        agents = {}

        class Capturing(cls):
            def _create_agent(self, state_dim, action_dim, hp, logger, device):
                agent = super()._create_agent(state_dim, action_dim, hp, logger, device)
                agents[logger.name] = agent
                return agent
        return Capturing(highway_config, artifacts_dir=str(tmp_path / cls.__name__), **kw), agents

    seq_runner, seq_agents = capture(ExperimentRunner)
    seq = [seq_runner.launch(e) for e in exps]
    mux_runner, mux_agents = capture(MultiplexedRunner, max_concurrent=16)
    mux = mux_runner.launch_many(exps)
    assert [r["experiment_name"] for r in mux] == [e.name for e in exps]
    for a, b in zip(seq, mux):
        assert a["status"] == b["status"] == "COMPLETED", (a.get("error_traceback"), b.get("error_traceback"))
        assert a["rewards"] == b["rewards"] and a["avg_rewards"] == b["avg_rewards"], a["experiment_name"]
        ha, hb = a["metrics_history"], b["metrics_history"]
        assert ha["episode_rewards"] == hb["episode_rewards"]
        assert len(ha["policy_updates"]) == len(hb["policy_updates"]) >= 2
        for ua, ub in zip(ha["policy_updates"], hb["policy_updates"]):
            assert {k: v for k, v in ua.items() if k != "time"} == {k: v for k, v in ub.items() if k != "time"}
    for name, agent in seq_agents.items():
        other = mux_agents[name]
        assert torch.equal(agent.actor_critic.flat, other.actor_critic.flat), name
        assert torch.equal(agent.optimizer.exp_avg, other.optimizer.exp_avg), name
        assert torch.equal(agent.optimizer.exp_avg_sq, other.optimizer.exp_avg_sq), name
    # the experiments really shared launches: far fewer simulator kernels than env interactions
    print("env launches", mux_runner.env_launches, "for", mux_runner.env_requests, "env.reset / env.step calls in",
          mux_runner.ticks, "ticks")
    assert mux_runner.env_launches < 0.8 * mux_runner.env_requests


def test_multiplexed_runner_reports_failures_per_experiment(highway_config, tmp_path):
    from highway_rope_ppo_b200.experiments.config import Condition, ConditionHP, Experiment
    from highway_rope_ppo_b200.experiments.multiplex import MultiplexedRunner

    good = _sweep_like_experiments(2, max_episodes=2, steps_per_update=32)
    bad = Experiment("bad", Condition.SHUFFLED_ROPE, ConditionHP(d_embed=16), seed=1, max_episodes=1)
    res = MultiplexedRunner(highway_config, artifacts_dir=str(tmp_path)).launch_many([good[0], bad, good[1]])
    assert [r["status"] for r in res] == ["COMPLETED", "FAILED", "COMPLETED"]
    assert "rotate_dim" in res[1]["error_message"] and "error_traceback" in res[1]


@pytest.mark.parametrize("groups,graphs,chunks", [(1, True, 1), (2, True, 1), (2, False, 1), (1, True, 4), (2, True, 2)])
def test_host_buffer_pipeline_equals_the_device_loop(highway_config, groups, graphs, chunks):
    """training/host_pipeline.py: the graphed, grouped host-buffer loop (pinned observation in, pinned results out)
    produces exactly the actions, observations, rewards and flags of the plain device-resident act + step loop."""
    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_vec_env
    from highway_rope_ppo_b200.ppo.agent import PPOAgent
    from highway_rope_ppo_b200.training.host_pipeline import HostBufferPipeline

    E, steps = 256, 6
    over = {"observation": {"order": "shuffled"}}
    mk = lambda n, base: make_vec_env(Condition.SHUFFLED_ROPE, highway_config, 4, over, num_envs=n, seed=42, env_id_base=base)
    torch.manual_seed(1)
    agent = PPOAgent(60, 2, hidden_dim=64, batch_size=E)
    torch.manual_seed(1)
    ref_agent = PPOAgent(60, 2, hidden_dim=64, batch_size=E)
    ref_env = mk(E, 0)
    obs = ref_env.reset(42).clone()
    want = []
    for _ in range(steps):
        out = ref_agent.act(obs.view(E, -1))
        o, r, te, tr = ref_env.step(out["action"])
        want.append((out["action"].cpu().numpy().copy(), o.cpu().numpy().copy(), r.cpu().numpy().copy(),
                     te.cpu().numpy().copy(), tr.cpu().numpy().copy()))
        obs = o.clone()
    Eg = E // groups
    pipe = HostBufferPipeline(agent, [mk(Eg, g * Eg) for g in range(groups)], use_graphs=graphs, chunks=chunks)
    pipe.reset(42)
    for t in range(steps):
        for g in range(groups):
            pipe.launch(g)
        for g in range(groups):
            b = pipe.wait(g)
            sl = slice(g * Eg, (g + 1) * Eg)
            a, o, r, te, tr = want[t]
            assert np.array_equal(b["action"], a[sl]), (t, g)
            assert np.array_equal(b["obs"], o[sl]) and np.array_equal(b["reward"], r[sl]), (t, g)
            assert np.array_equal(b["terminated"], te[sl]) and np.array_equal(b["truncated"], tr[sl]), (t, g)
    pipe.close()


def test_graph_captured_rollout_equals_the_eager_rollout(highway_config):
    """collect_rollout: the whole T-step rollout replayed as one CUDA graph (device-resident draw counter) fills the
    rollout buffers with exactly what the launch-by-launch loop writes, rollout after rollout, with updates between."""
    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_vec_env
    from highway_rope_ppo_b200.ppo.agent import PPOAgent
    from highway_rope_ppo_b200.training.routine import collect_rollout

    E, T = 128, 6
    over = {"observation": {"order": "shuffled"}}
    runs = []
    for use_graph in (False, True):
        torch.manual_seed(4)
        np.random.seed(4)
        env = make_vec_env(Condition.SHUFFLED_ROPE, highway_config, 4, over, num_envs=E, seed=7)
        agent = PPOAgent(60, 2, hidden_dim=64, batch_size=256, epochs=1)
        obs = env.reset(7).clone()
        got = []
        for it in range(4):
            r = collect_rollout(env, agent, T, obs, use_graph=use_graph)
            got.append({k: v.clone() for k, v in r.items()})
            obs = r["states"][T].clone()
            if it == 1:
                agent.act(obs.view(E, -1))   # an unrelated act() in between advances the draw counter
            _, _, v = agent.actor_critic.forward(obs.view(E, -1))
            agent.update(last_value=v.view(-1))
        runs.append(got)
        env.close()
    for a, b in zip(*runs):
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_fetch_host_kernel_copies_pinned_buffers_exactly():
    """hrp_fetch_host: the kernel-side host-to-device copy of a page-locked buffer (any byte count, 16-byte aligned
    pointers), and its argument checks."""
    from highway_rope_ppo_b200 import _lib

    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for n in (1, 15, 16, 17, 4096, 983040, 983040 + 7):
        src = torch.randint(0, 256, (n,), dtype=torch.uint8).pin_memory()
        dst = torch.zeros(n + 16, dtype=torch.uint8, device="cuda:0")
        assert lib.hrp_fetch_host(dst.data_ptr(), src.data_ptr(), n, st) == 0, _lib.last_error()
        torch.cuda.synchronize()
        assert torch.equal(dst[:n].cpu(), src) and int(dst[n:].sum()) == 0, n
    src = torch.zeros(64, dtype=torch.uint8).pin_memory()
    dst = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
    assert lib.hrp_fetch_host(dst.data_ptr() + 4, src.data_ptr(), 16, st) == -1          # misaligned destination
    pageable = torch.zeros(64, dtype=torch.uint8)
    assert lib.hrp_fetch_host(dst.data_ptr(), pageable.data_ptr(), 16, st) == -1          # not page-locked
    assert lib.hrp_fetch_host(dst.data_ptr(), src.data_ptr(), 0, st) == 0


def test_command_line_front_end_runs_an_array_task(tmp_path, capsys):
    """main.py re-hosted (python -m highway_rope_ppo_b200.main): one array task of the grid, multiplexed, end to end."""
    from highway_rope_ppo_b200 import main as cli

    rc = cli.main(["--array-task-id", "1", "--slurm-num-tasks", "135", "--n-jobs", "4", "--max-episodes", "4",
                   "--artifacts-dir", str(tmp_path)])          # ceil(540 / 135) = 4 experiments: grid entries 4 .. 7
    out = capsys.readouterr().out
    assert rc == 0 and "best avg_reward" in out
    names = sorted(p.name for p in tmp_path.glob("summary_*.csv"))
    assert len(names) == 4 and all(n.startswith("summary_sorted_lr0.0001_hidden_dim128") for n in names)
