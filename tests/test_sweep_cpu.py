"""Sweep front-end against the reference's own experiment grid: tests/golden/sweep_experiments.json
holds the names main.py:42-88 generates (tools/gen_golden.py:gen_sweep)."""
import json
import os

import pytest

from highway_rope_ppo_b200.experiments import sweep
from highway_rope_ppo_b200.experiments.config import Condition

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sweep_experiments.json")))


def test_experiment_grid_matches_the_reference_names_and_order():
    ex = sweep.define_experiments(GOLD["seed"], 3)
    assert len(ex) == GOLD["count"] == 540
    assert [e.name for e in ex] == GOLD["names"]
    assert sorted({e.seed for e in ex}) == GOLD["seeds"]
    for k, v in GOLD["first_hp"].items():
        assert getattr(ex[0].hp, k) == v, k
    assert ex[0].condition is Condition.SORTED and ex[-1].condition is Condition.SHUFFLED_ROPE
    assert all(not e.hp.sweep for e in ex)  # expanded entries carry no sweep of their own


def test_selection_rules():
    ex = sweep.define_experiments(42, 1)
    n = len(ex)
    # SLURM-array batches: contiguous ceil(n / tasks) slices that cover the list exactly once
    tasks = 7
    got = [e.name for t in range(tasks) for e in sweep.select_experiments(ex, array_task_id=t, num_tasks=tasks)]
    assert got == [e.name for e in ex]
    assert sweep.select_experiments(ex, array_task_id=tasks + 5, num_tasks=tasks) == []
    # exact name first, then unique prefix
    assert sweep.select_experiments(ex, single=ex[3].name) == [ex[3]]
    prefix = ex[-1].name[:-len("seed42")]
    assert sweep.select_experiments(ex, single=prefix) == [ex[-1]]
    with pytest.raises(ValueError):
        sweep.select_experiments(ex, single="sorted_")        # ambiguous
    with pytest.raises(ValueError):
        sweep.select_experiments(ex, single="no_such_run")    # missing
    # round-robin shard over ranks: a partition
    parts = [sweep.shard_for_rank(ex, r, 8) for r in range(8)]
    assert sorted(e.name for p in parts for e in p) == sorted(e.name for e in ex)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert n == 180


def test_summarize_picks_the_best_final_average_per_condition():
    res = [{"experiment_name": "sorted_lr0.0001_seed42", "avg_rewards": [1.0, 5.0]},
           {"experiment_name": "sorted_lr0.0003_seed42", "avg_rewards": [9.0, 4.0]},
           {"experiment_name": "shuffled_rope_lr0.0003_seed42", "avg_rewards": [2.0]},
           {"experiment_name": "shuffled_lr0.0003_seed42", "status": "FAILED"}]
    best = sweep.summarize(res)
    assert best["sorted"] == (5.0, "sorted_lr0.0001_seed42")
    assert best["shuffled"] == (2.0, "shuffled_rope_lr0.0003_seed42")  # the reference keys on the first name token
    assert len(best) == 2


def test_run_name_codec_round_trips_every_experiment_of_the_grid():
    from highway_rope_ppo_b200.experiments.config import ConditionHP, Experiment

    for name in GOLD["names"][::7]:
        f = Experiment.parse_name(name)
        hp = ConditionHP(**{k: f[k] for k in ("lr", "hidden_dim", "clip_eps", "entropy_coef", "epochs", "batch_size", "d_embed")})
        assert Experiment.make_name(f["condition"], hp, f["seed"]) == name
    f = Experiment.parse_name("shuffled_rope_lr0.0003_hidden_dim256_clip_eps0.2_entropy_coef0.005_epochs8_batch_size64_d_embed16_seed2042")
    assert f["condition"] is Condition.SHUFFLED_ROPE and f["hidden_dim"] == 256 and f["d_embed"] == 16 and f["seed"] == 2042
    assert isinstance(f["epochs"], int) and isinstance(f["lr"], float)
    with pytest.raises(ValueError):
        Experiment.parse_name("ppo_highway_best_sorted.pth")


def test_hyper_parameter_records_validate_and_expand():
    from highway_rope_ppo_b200.experiments.config import CommonHP, ConditionHP, expand_condition_hps

    assert [c.value for c in Condition] == [1, 2, 3, 4, 5]   # enum.auto() numbering of the reference
    with pytest.raises(ValueError):
        ConditionHP(lr=0.0)
    with pytest.raises(ValueError):
        CommonHP(gamma=1.5)
    with pytest.raises(ValueError):
        ConditionHP(sweep={"learning_rate": [1e-3]})
    hp = ConditionHP(sweep={"lr": [1e-4, 3e-4], "epochs": [6, 8, 10]})
    got = expand_condition_hps(hp)
    assert [(h.lr, h.epochs) for h in got] == [(1e-4, 6), (1e-4, 8), (1e-4, 10), (3e-4, 6), (3e-4, 8), (3e-4, 10)]
    assert all(h.sweep == {} and h.hidden_dim == hp.hidden_dim for h in got)
    plain = ConditionHP()
    assert expand_condition_hps(plain) == [plain]


def test_every_run_name_is_readable_by_the_reference_offline_tools():
    """results.py:33-44 / analysis.py:21-32 parse run names with one regular expression (fixture: result_schema.json)."""
    import re

    from highway_rope_ppo_b200.experiments.config import Experiment

    schema = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "result_schema.json")))
    rx = re.compile(schema["exp_rx"])
    pe = {"SORTED": ("sorted", None), "SHUFFLED": ("shuffled", None), "SHUFFLED_RANKPE": ("shuffled", "rankpe"),
          "SHUFFLED_DISTPE": ("shuffled", "distpe"), "SHUFFLED_ROPE": ("shuffled", "rope")}
    for e in sweep.define_experiments(42, 3):
        m = rx.match(e.name)
        assert m, e.name
        d = m.groupdict()
        assert (d["prefix"], d["pe_type"]) == pe[e.condition.name]
        mine = Experiment.parse_name(e.name)
        assert float(d["lr"]) == mine["lr"] == e.hp.lr and int(d["hidden_dim"]) == mine["hidden_dim"] == e.hp.hidden_dim
        assert int(d["epochs"]) == e.hp.epochs and int(d["batch_size"]) == e.hp.batch_size and int(d["seed"]) == e.seed
        assert int(d["d_embed"]) == e.hp.d_embed and float(d["clip_eps"]) == e.hp.clip_eps


def test_command_line_front_end_selection(capsys):
    """main.py:90-245 re-hosted: the grid size, the single-experiment rules and the array-task slices, without a GPU."""
    from highway_rope_ppo_b200 import main as cli
    from highway_rope_ppo_b200.experiments.sweep import define_experiments, select_experiments

    assert cli.main(["--get-total-experiments"]) == 0
    assert capsys.readouterr().out.strip() == "540"
    assert cli.main(["--get-total-experiments", "--num-seeds", "1"]) == 0
    assert capsys.readouterr().out.strip() == "180"
    assert cli.main(["--run-single-experiment", "no_such_experiment"]) == 1            # reference: exit(1)
    assert cli.main(["--run-single-experiment", "sorted_lr0.0001"]) == 1                # ambiguous prefix
    assert cli.main(["--generate-slurm"]) == 2
    a = cli.parse(["--array-task-id", "3", "--slurm-num-tasks", "68", "--n-jobs", "8"])
    exps = define_experiments()
    sel = select_experiments(exps, a.array_task_id, a.slurm_num_tasks, a.run_single_experiment)
    assert [e.name for e in sel] == [e.name for e in exps[24:32]]                       # ceil(540 / 68) = 8 per task
