"""The sweep front-end of the reference (``main.py:42-88`` ``define_experiments``, ``:30-40`` ``summarize``,
``:192-243`` experiment selection and launch), re-hosted for "one process per GPU".

The experiment grid, the run-name format (parsed by the reference's ``analysis.py`` / ``results.py``) and the
selection rules (``--array-task-id`` batches, ``--run-single-experiment`` exact-then-prefix match) are the
reference's.  What replaces ``joblib`` + ``DevicePool`` time-sharing is a static shard -- rank r of W runs experiments
r, r + W, r + 2W, ... of the selected list on its own GPU -- and, within a GPU, ``experiments/multiplex.py``: R runs
at a time on shared env handles instead of ``OVERSUB`` processes time-sharing the device.
"""
from __future__ import annotations

import math
import os
from collections import defaultdict
from typing import Any, Dict, Iterable, List, Optional

from ..utils.reproducibility import SEED
from .config import CommonHP, Condition, ConditionHP, Experiment, expand_condition_hps
from .runner import ExperimentRunner

# main.py:50-60 -- the full-grid sweep attached to every condition
SWEEP: Dict[str, List[Any]] = {
    "lr": [1e-4, 3e-4],
    "hidden_dim": [128, 256, 384],
    "clip_eps": [0.2],
    "entropy_coef": [0.005],
    "epochs": [6, 8, 10],
    "batch_size": [32, 64],
    "d_embed": [4],  # with 4 features per vehicle only 4 channels can rotate
}
CONDITIONS = (Condition.SORTED, Condition.SHUFFLED, Condition.SHUFFLED_RANKPE, Condition.SHUFFLED_DISTPE,
              Condition.SHUFFLED_ROPE)


def define_experiments(base_seed: int = SEED, num_seeds: int = 3) -> List[Experiment]:
    """Conditions x hyper-parameter grid x seeds (``base_seed + 1000 i``), in the reference's order and with the
    reference's names: ``<condition>_<key><value>..._seed<seed>`` with the keys in sweep order."""
    experiments: List[Experiment] = []
    for cond in CONDITIONS:
        template = ConditionHP(**vars(CommonHP()))
        template.sweep = {k: list(v) for k, v in SWEEP.items()}
        for hp in expand_condition_hps(template):
            for i in range(num_seeds):
                seed = base_seed + i * 1000
                experiments.append(Experiment(Experiment.make_name(cond, hp, seed, tuple(template.sweep)), cond, hp, seed))
    return experiments


def select_experiments(all_experiments: List[Experiment], array_task_id: Optional[int] = None,
                       num_tasks: Optional[int] = None, single: Optional[str] = None) -> List[Experiment]:
    """``main.py:196-228``: a SLURM-array batch (contiguous ceil(n / tasks) slices), one experiment by exact name or
    unique prefix (``ValueError`` when ambiguous or missing, where the reference exits 1), or everything."""
    if array_task_id is not None:
        tasks = num_tasks or int(os.getenv("SLURM_ARRAY_TASK_COUNT", 1))
        per = math.ceil(len(all_experiments) / tasks)
        return all_experiments[array_task_id * per:min((array_task_id + 1) * per, len(all_experiments))]
    if single:
        matches = [e for e in all_experiments if e.name == single] or \
                  [e for e in all_experiments if e.name.startswith(single)]
        if len(matches) != 1:
            raise ValueError(f"Experiment '{single}' selection ambiguous or not found.")
        return matches
    return list(all_experiments)


def shard_for_rank(experiments: List[Experiment], rank: int, world: int) -> List[Experiment]:
    """The experiments rank ``rank`` of ``world`` runs (round-robin, so long and short runs mix evenly)."""
    return experiments[rank::world]


def run_experiments(experiments: Iterable[Experiment], base_env_config: dict, artifacts_dir: Optional[str] = None,
                    runner: Optional[ExperimentRunner] = None, multiplex: int = 0) -> List[Dict[str, Any]]:
    """Run the experiments on this process's GPU: one after the other, or -- ``multiplex`` = R > 1 -- R at a time on
    shared env handles (``experiments/multiplex.py``; bit-identical results, several times the runs per hour).  A
    sweep over W GPUs is ``run_experiments(shard_for_rank(selected, rank, W), ..., multiplex=R)`` in each rank."""
    if multiplex and multiplex > 1 and runner is None:
        from .multiplex import MultiplexedRunner

        return MultiplexedRunner(base_env_config, artifacts_dir=artifacts_dir, max_concurrent=multiplex).launch_many(
            list(experiments))
    runner = runner or ExperimentRunner(base_env_config, artifacts_dir=artifacts_dir)
    return [runner.launch(exp) for exp in experiments]


def summarize(results: Iterable[Dict[str, Any]]) -> Dict[str, Any]:
    """Best final average reward per condition (``main.py:30-40``); returns {condition: (score, name)}.  Failed
    runs carry no rewards and are skipped."""
    best = defaultdict(lambda: (-1e9, ""))
    for r in results:
        if not r.get("avg_rewards"):
            continue
        cond = r["experiment_name"].split("_")[0]
        avg = r["avg_rewards"][-1]
        if avg > best[cond][0]:
            best[cond] = (avg, r["experiment_name"])
    return dict(best)
