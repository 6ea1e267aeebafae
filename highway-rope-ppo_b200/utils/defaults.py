"""Quantities the embedding wrappers derive from the env definition (reference: ``utils/defaults.py:10-23``).
Every helper takes the config to read; the reference's zero-argument calls read ``HIGHWAY_CONFIG``."""
from typing import Any, Mapping, Optional

from ..config.base_config import HIGHWAY_CONFIG


def _observation(cfg: Optional[Mapping[str, Any]]) -> Mapping[str, Any]:
    return (HIGHWAY_CONFIG if cfg is None else cfg)["observation"]


def max_dist(cfg: Optional[Mapping[str, Any]] = None) -> float:
    """Largest |x| or |y| (metres) that survives the observation clip: the distance scale of RoPE / DistPE."""
    span = _observation(cfg)["features_range"]
    return float(max(abs(bound) for axis in "xy" for bound in span[axis]))


def max_rank(cfg: Optional[Mapping[str, Any]] = None) -> int:
    """Rows of the Kinematics observation (the ego row included)."""
    return int(_observation(cfg)["vehicles_count"])


def feature_count(cfg: Optional[Mapping[str, Any]] = None) -> int:
    """Scalar features per observed vehicle."""
    return len(_observation(cfg)["features"])
