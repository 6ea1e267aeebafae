"""Minimal gymnasium-protocol pieces.

gymnasium is used when importable, so that ``isinstance(space, gymnasium.spaces.Box)`` checks in
the reference runner (``experiments/runner.py:87-98``) hold; otherwise a small stand-in with the
same attributes is provided (the reference only reads ``shape``, ``low``, ``high``, ``dtype``).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    from gymnasium import spaces as _gym_spaces

    Box = _gym_spaces.Box
    Discrete = _gym_spaces.Discrete
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent: stand-ins
    HAVE_GYMNASIUM = False

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(int(s) for s in shape)
            self.dtype = np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete:  # type: ignore[no-redef]
        def __init__(self, n: int):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def contains(self, x) -> bool:
            return 0 <= int(x) < self.n

        def sample(self):
            return int(np.random.randint(self.n))

        def __repr__(self):
            return f"Discrete({self.n})"


def is_box(space) -> bool:
    """True for gymnasium's Box and for the stand-in (duck-typed on low/high/shape)."""
    return isinstance(space, Box) or all(hasattr(space, a) for a in ("low", "high", "shape", "dtype"))
