"""Experiment description types (reference: ``experiments/config.py:9-70``)."""
from __future__ import annotations

import copy
import itertools
from dataclasses import dataclass, field, fields
from enum import Enum, auto
from typing import Any, Dict, List, Optional


class Condition(Enum):
    SORTED = auto()
    SHUFFLED = auto()
    SHUFFLED_RANKPE = auto()
    SHUFFLED_DISTPE = auto()
    SHUFFLED_ROPE = auto()


@dataclass
class CommonHP:
    """Hyper-parameters shared by every condition."""

    gamma: float = 0.99
    lam: float = 0.95
    value_coef: float = 0.5
    entropy_coef: float = 0.005
    max_grad_norm: float = 0.5
    steps_per_update: int = 2048


@dataclass
class ConditionHP(CommonHP):
    """Per-condition hyper-parameters; ``sweep`` maps a field name to the values to try."""

    lr: float = 1e-4
    clip_eps: float = 0.2
    epochs: int = 6
    batch_size: int = 64
    hidden_dim: int = 128
    d_embed: Optional[int] = None
    sweep: Dict[str, List[Any]] = field(default_factory=dict)


@dataclass
class Experiment:
    """One training run."""

    name: str
    condition: Condition
    hp: ConditionHP = field(default_factory=ConditionHP)
    seed: int = 42
    max_episodes: int = 1500
    target_reward: float = 130.0
    env_config_overrides: Dict[str, Any] = field(default_factory=dict)
    extra: Dict[str, Any] = field(default_factory=dict)


def expand_condition_hps(hp: ConditionHP) -> List[ConditionHP]:
    """Cartesian product of ``hp.sweep``; the expanded entries carry an empty sweep."""
    if not hp.sweep:
        return [hp]
    base = {f.name: copy.deepcopy(getattr(hp, f.name)) for f in fields(hp) if f.name != "sweep"}
    names = list(hp.sweep)
    out = []
    for combo in itertools.product(*(hp.sweep[n] for n in names)):
        out.append(ConditionHP(**{**copy.deepcopy(base), **dict(zip(names, combo))}))
    return out
