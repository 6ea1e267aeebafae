"""Shared plumbing of the three observation wrappers (``gymnasium.ObservationWrapper`` protocol).

The reference wrappers post-process a numpy observation on the host after every ``env.step``
(``rope_embed.py:64``, ``dist_embed.py:76``, ``rank_embed.py:45``).  Here the arithmetic runs on
the GPU: when the wrapped env is this package's simulator, the wrapper asks it to fuse the
embedding into the step kernel's epilogue (no second pass over the observation); for any other
env, and for direct ``observation(obs)`` calls, the standalone ``hrp_embed_apply`` kernel is used.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..envs.highway_vec import EmbedSpec, embed_apply
from ..envs.spaces import is_box


def check_2d_box(env, who: str):
    """The reference's space checks (``rank_embed.py:12-17``, ``dist_embed.py:20-26``)."""
    space = env.observation_space
    if not is_box(space):
        raise TypeError(f"{who} requires Box observation space.")
    if len(space.shape) != 2:
        raise ValueError(f"{who} requires 2D Box observation space (N, F).")
    return space.shape


class EmbedWrapperBase:
    """step/reset pass-through + ``observation``; subclasses provide ``_spec()``."""

    def __init__(self, env):
        self.env = env
        self._fused = False
        self._table_dev: Optional[torch.Tensor] = None
        self._device = torch.device("cpu")

    # -- gymnasium.Wrapper surface ---------------------------------------------------------
    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    @property
    def action_space(self):
        return self.env.action_space

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return (obs if self._fused else self.observation(obs)), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return (obs if self._fused else self.observation(obs)), reward, terminated, truncated, info

    def close(self):
        return self.env.close()

    # -- embedding ----------------------------------------------------------------------------
    def _spec(self) -> EmbedSpec:  # pragma: no cover - abstract
        raise NotImplementedError

    def _try_fuse(self) -> None:
        """Ask a native simulator env to apply this wrapper inside its step kernel."""
        fuse = getattr(self.env, "_fuse_embedding", None)
        if callable(fuse) and not getattr(self.env, "_embed_fused", False):
            fuse(self._spec())
            self.env._embed_fused = True
            self._fused = True

    def _cuda_device(self) -> torch.device:
        _lib.require_device()
        if self._device.type == "cuda":
            return self._device
        return torch.device("cuda", torch.cuda.current_device())

    def _apply(self, obs: np.ndarray, dist_override: Optional[np.ndarray] = None) -> np.ndarray:
        spec = self._spec()
        dev = self._cuda_device()
        if self._table_dev is None or self._table_dev.device != dev:
            self._table_dev = torch.from_numpy(spec.table).to(dev)
        x = torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float32)).to(dev)
        d = None if dist_override is None else torch.from_numpy(
            np.ascontiguousarray(dist_override, dtype=np.float32).reshape(1, -1)).to(dev)
        out = embed_apply(spec.kind, x, self._table_dev, spec.dim, spec.max_dist, spec.use_euclidean,
                          spec.ego_idx, d)
        return out[0].cpu().numpy()
