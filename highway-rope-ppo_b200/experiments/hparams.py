"""Hyper-parameter records and the experiment record of the sweep, with the run-name codec.

The reference keeps three plain dataclasses in ``experiments/config.py:9-70`` and formats run names inline in
``main.py:77-87``; its offline tools (``analysis.py``, ``results.py``, ``visualize.py``) parse those names back
with regular expressions.  Here the records validate themselves and own both directions of the name format, so
that results written by this framework stay readable by the reference's tools.
"""
from __future__ import annotations

import copy
import dataclasses
import itertools
import re
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, Iterator, List, Optional, Tuple

# member values 1..5 in this order, as ``enum.auto()`` numbers them in the reference
Condition = Enum("Condition", ["SORTED", "SHUFFLED", "SHUFFLED_RANKPE", "SHUFFLED_DISTPE", "SHUFFLED_ROPE"])
Condition.__doc__ = "Observation ordering x positional embedding of one run (reference: experiments/config.py:9-14)."

#: keys that appear in a run name, in name order (``main.py:50-60``)
NAME_KEYS: Tuple[str, ...] = ("lr", "hidden_dim", "clip_eps", "entropy_coef", "epochs", "batch_size", "d_embed")
_NAME_RX = re.compile(r"^(?P<cond>[a-z_]+?)_" + "_".join(rf"{k}(?P<{k}>[^_]+)" for k in NAME_KEYS) + r"_seed(?P<seed>\d+)$")


def _positive(name: str, value: Any, strict: bool = True) -> None:
    if value is None or (value <= 0 if strict else value < 0):
        raise ValueError(f"{name} must be {'positive' if strict else 'non-negative'}, got {value!r}")


@dataclass
class CommonHP:
    """PPO settings every condition shares (defaults: the reference's ``CommonHP``)."""

    gamma: float = 0.99
    lam: float = 0.95
    value_coef: float = 0.5
    entropy_coef: float = 0.005
    max_grad_norm: float = 0.5
    steps_per_update: int = 2048

    def __post_init__(self) -> None:
        for k in ("gamma", "lam"):
            if not 0.0 <= getattr(self, k) <= 1.0:
                raise ValueError(f"{k} must lie in [0, 1], got {getattr(self, k)!r}")
        _positive("steps_per_update", self.steps_per_update)
        _positive("max_grad_norm", self.max_grad_norm, strict=False)


@dataclass
class ConditionHP(CommonHP):
    """``CommonHP`` plus what a condition may tune; ``sweep`` maps a field name to the values to try."""

    lr: float = 1e-4
    clip_eps: float = 0.2
    epochs: int = 6
    batch_size: int = 64
    hidden_dim: int = 128
    d_embed: Optional[int] = None
    sweep: Dict[str, List[Any]] = field(default_factory=dict)

    def __post_init__(self) -> None:
        super().__post_init__()
        for k in ("lr", "epochs", "batch_size", "hidden_dim"):
            _positive(k, getattr(self, k))
        unknown = [k for k in self.sweep if k == "sweep" or k not in {f.name for f in dataclasses.fields(self)}]
        if unknown:
            raise ValueError(f"sweep names unknown hyper-parameters: {unknown}")

    def expand(self) -> Iterator["ConditionHP"]:
        """Cartesian product of ``sweep`` (insertion order of the keys, first key slowest); the expanded records
        carry an empty sweep.  Without a sweep the record itself is the only element."""
        if not self.sweep:
            yield self
            return
        fixed = {f.name: getattr(self, f.name) for f in dataclasses.fields(self) if f.name != "sweep"}
        names = tuple(self.sweep)
        for combo in itertools.product(*(self.sweep[n] for n in names)):
            yield ConditionHP(**{**copy.deepcopy(fixed), **dict(zip(names, combo))})

    def name_part(self, keys: Tuple[str, ...] = NAME_KEYS) -> str:
        return "_".join(f"{k}{getattr(self, k)}" for k in keys)


@dataclass
class Experiment:
    """One training run of the sweep."""

    name: str
    condition: Condition
    hp: ConditionHP = field(default_factory=ConditionHP)
    seed: int = 42
    max_episodes: int = 1500
    target_reward: float = 130.0
    env_config_overrides: Dict[str, Any] = field(default_factory=dict)
    extra: Dict[str, Any] = field(default_factory=dict)

    @staticmethod
    def make_name(condition: Condition, hp: ConditionHP, seed: int, keys: Tuple[str, ...] = NAME_KEYS) -> str:
        """``<condition>_<key><value>..._seed<seed>`` (``main.py:77-87``)."""
        return f"{condition.name.lower()}_{hp.name_part(keys)}_seed{seed}"

    @staticmethod
    def parse_name(name: str) -> Dict[str, Any]:
        """Inverse of :meth:`make_name`: condition, the named hyper-parameters (typed) and the seed."""
        m = _NAME_RX.match(name)
        if not m:
            raise ValueError(f"not a run name of the sweep: {name!r}")
        types = {f.name: f.type for f in dataclasses.fields(ConditionHP)}
        out: Dict[str, Any] = {"condition": Condition[m["cond"].upper()], "seed": int(m["seed"])}
        for k in NAME_KEYS:
            raw = m[k]
            out[k] = None if raw == "None" else (int(raw) if "int" in str(types[k]) else float(raw))
        return out


def expand_condition_hps(hp: ConditionHP) -> List[ConditionHP]:
    """List form of :meth:`ConditionHP.expand` (the reference's helper of the same name)."""
    return list(hp.expand())
