"""``experiments.config`` of the reference (``experiments/config.py:9-70``): the import path its callers use.
The records live in :mod:`.hparams`."""
from .hparams import NAME_KEYS, CommonHP, Condition, ConditionHP, Experiment, expand_condition_hps

__all__ = ["NAME_KEYS", "CommonHP", "Condition", "ConditionHP", "Experiment", "expand_condition_hps"]
