"""Sinusoidal distance code (reference: ``experiments/dist_embed.py:8-96``).

Appends ``[sin(2*pi*d_hat*f_k), cos(2*pi*d_hat*f_k)]``, ``f_k = exp(-2k ln(base)/d_embed)``,
to every observation row; ``d_hat`` is the clipped normalised distance to the ego row
(Euclidean over the first two features, or |Δ| of the first feature).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .._lib import EMBED_DIST
from ..envs.highway_vec import EmbedSpec
from ..envs.spaces import Box
from ..utils.defaults import max_dist as _max_dist
from ._wrapper_base import EmbedWrapperBase, check_2d_box


def dist_freqs(d_embed: int, base: float) -> torch.Tensor:
    """float32 frequencies, computed with torch exactly as dist_embed.py:48-52."""
    return torch.exp(-torch.arange(0, d_embed, 2, dtype=torch.float32) * (np.log(base) / d_embed))


def extended_space(space, d_embed: int):
    """Observation space with ``d_embed`` extra columns bounded by [-1, 1]."""
    N, F = space.shape
    low = np.concatenate([space.low, -np.ones((N, d_embed))], axis=1)
    high = np.concatenate([space.high, np.ones((N, d_embed))], axis=1)
    return Box(low, high, (N, F + d_embed), np.float32)


class DistanceEmbedWrapper(EmbedWrapperBase):
    def __init__(self, env, d_embed: int = 8, max_dist: float = _max_dist(), base: Optional[float] = None,
                 use_euclidean: bool = True, ego_idx: int = 0):
        super().__init__(env)
        N, F = check_2d_box(env, "DistanceEmbedWrapper")
        self.d_embed = d_embed
        if self.d_embed % 2 != 0:
            raise ValueError(f"DistanceEmbedWrapper requires even d_embed; got {self.d_embed}")
        self.max_dist = float(max_dist)
        self.use_euclidean = use_euclidean
        self.ego_idx = ego_idx
        need = 2 if use_euclidean else 1
        if F < need:
            raise ValueError(f"DistanceEmbedWrapper requires at least {need} feature(s) for distance "
                             f"calculation (features available: {F}).")
        self.freqs = dist_freqs(d_embed, base or self.max_dist)
        self._freqs_np = self.freqs.cpu().numpy()
        self.observation_space = extended_space(env.observation_space, d_embed)
        self._try_fuse()

    def to(self, device):
        self.freqs = self.freqs.to(device)
        self._device = torch.device(device)
        if hasattr(self.env, "to"):
            self.env.to(device)
        return self

    def _spec(self) -> EmbedSpec:
        return EmbedSpec(EMBED_DIST, self.d_embed, self._freqs_np, self.max_dist, self.use_euclidean, self.ego_idx)

    def observation(self, obs: np.ndarray) -> np.ndarray:
        return self._apply(obs)
