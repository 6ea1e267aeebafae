"""Defaults derived from HIGHWAY_CONFIG (reference: ``utils/defaults.py:10-23``)."""
from ..config.base_config import HIGHWAY_CONFIG


def max_dist() -> float:
    """Largest |x| / |y| the observation normalisation admits, in metres."""
    ranges = HIGHWAY_CONFIG["observation"]["features_range"]
    return max(abs(v) for axis in ("x", "y") for v in ranges[axis])


def max_rank() -> int:
    """Rows in the Kinematics observation."""
    return HIGHWAY_CONFIG["observation"]["vehicles_count"]


def feature_count() -> int:
    """Scalar features per observed vehicle."""
    return len(HIGHWAY_CONFIG["observation"]["features"])
