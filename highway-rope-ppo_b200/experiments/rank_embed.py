"""Row-index tag (reference: ``experiments/rank_embed.py:9-51``).

Appends ``tanh(table[r])`` to observation row ``r``; the table is an ``nn.Embedding(N, d_embed)``
re-initialised ``U(-0.05, 0.05)`` from the torch global generator (so it is fixed by
``set_random_seeds``) and never trained.

Deviation from the reference at HEAD: ``rank_embed.py:48`` calls ``.numpy()`` on a tensor that
requires grad and raises ``RuntimeError``; the intended value, ``tanh(table).detach()``, is what
is implemented here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .._lib import EMBED_RANK
from ..envs.highway_vec import EmbedSpec
from ..utils.defaults import feature_count as _F
from ._wrapper_base import EmbedWrapperBase, check_2d_box
from .dist_embed import extended_space


class RankEmbedWrapper(EmbedWrapperBase):
    def __init__(self, env, d_embed: int = _F()):
        super().__init__(env)
        N, F = check_2d_box(env, "RankEmbedWrapper")
        self.d_embed = d_embed
        # same constructor calls as the reference, so the global RNG stream advances identically
        self.table = nn.Embedding(N, d_embed)
        self.table.weight.data.uniform_(-0.05, 0.05)
        self.observation_space = extended_space(env.observation_space, d_embed)
        self._try_fuse()

    def to(self, device):
        self._device = torch.device(device)
        if hasattr(self.env, "to"):
            self.env.to(device)
        return self

    def _spec(self) -> EmbedSpec:
        tag = torch.tanh(self.table.weight.detach().float().cpu()).numpy()
        return EmbedSpec(EMBED_RANK, self.d_embed, tag, 100.0, True, 0)

    def observation(self, obs: np.ndarray) -> np.ndarray:
        return self._apply(obs)
