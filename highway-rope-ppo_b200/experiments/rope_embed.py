"""RoPE observation wrapper (reference: ``experiments/rope_embed.py:6-74``).

Rotates the first ``rotate_dim`` features of every observed vehicle, pair by pair, by an angle
``2*pi * d_hat * inv_freq[p]`` where ``d_hat`` is the clipped, ``max_dist``-normalised distance
of the row to the ego row.  The observation shape is unchanged.  All arithmetic is float32 as
in the reference; it runs in the step kernel's epilogue (fused) or in ``hrp_embed_apply``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .._lib import EMBED_ROPE
from ..envs.highway_vec import EmbedSpec
from ..utils.defaults import max_dist as _max_dist
from ._wrapper_base import EmbedWrapperBase


def rope_inv_freq(rotate_dim: int, base: float) -> np.ndarray:
    """One inverse frequency per rotated pair: ``base ** -(p / P)`` in float32 (rope_embed.py:37-39)."""
    pairs = rotate_dim // 2
    return (1.0 / (base ** (np.arange(pairs, dtype=np.float32) / pairs))).astype(np.float32)


class RotaryEmbedWrapper(EmbedWrapperBase):
    def __init__(self, env, rotate_dim: Optional[int] = None, max_dist: float = _max_dist(),
                 base: Optional[float] = None, ego_idx: int = 0):
        super().__init__(env)
        N, F = env.observation_space.shape
        self.rotate_dim = rotate_dim or (F - (F % 2))
        if self.rotate_dim % 2 != 0 or self.rotate_dim > F:
            raise ValueError(f"rotate_dim must be even and ≤ {F}; got {self.rotate_dim}")
        self.max_dist = float(max_dist)
        self.ego_idx = ego_idx
        self.inv_freq = rope_inv_freq(self.rotate_dim, base or self.max_dist)
        self.observation_space = env.observation_space
        self._try_fuse()

    def _spec(self) -> EmbedSpec:
        return EmbedSpec(EMBED_ROPE, self.rotate_dim, self.inv_freq, self.max_dist, True, self.ego_idx)

    def _apply_rope(self, obs: np.ndarray, dist_norm: np.ndarray) -> np.ndarray:
        """Rotation with caller-supplied normalised distances (rope_embed.py:44-62)."""
        return self._apply(obs, dist_override=dist_norm)

    def observation(self, obs: np.ndarray) -> np.ndarray:
        return self._apply(obs)
