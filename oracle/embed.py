"""oracle/embed.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy float32 restatement of the three observation wrappers of the reference:
``experiments/rope_embed.py:37-39,44-74`` (RoPE), ``experiments/dist_embed.py:48-52,76-96``
(DistPE) and ``experiments/rank_embed.py:21-22,45-51`` (RankPE, intended semantics: the HEAD
version raises).  PINNED: ``tests/test_oracle_cpu.py`` checks every function against
``tests/golden/embed_*.npz``, which ``tools/gen_golden.py`` produced by running the reference's own
wrapper classes.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.
"""
import numpy as np


def rope_inv_freq(rotate_dim: int, base: float) -> np.ndarray:
    """rope_embed.py:37-39"""
    pairs = rotate_dim // 2
    return 1.0 / (base ** (np.arange(pairs, dtype=np.float32) / pairs))


def dist_freqs(d_embed: int, base: float) -> np.ndarray:
    """dist_embed.py:48-52 (torch.exp on float32; numpy's float32 exp agrees to 1 ulp)"""
    k = np.arange(0, d_embed, 2, dtype=np.float32)
    return np.exp(-k * np.float32(np.log(base) / d_embed)).astype(np.float32)


def dist_norm(obs: np.ndarray, max_dist: float, use_euclidean: bool = True, ego_idx: int = 0) -> np.ndarray:
    """rope_embed.py:67-72 / dist_embed.py:79-88: clip(||obs[:, :2] - obs[ego, :2]|| / max_dist, 0, 1)"""
    obs = np.asarray(obs, dtype=np.float32)
    if use_euclidean:
        rel = obs[:, :2] - obs[ego_idx, :2][None, :]
        d = np.linalg.norm(rel, axis=-1)
    else:
        d = np.abs(obs[:, 0] - obs[ego_idx, 0])
    return np.clip(d / max_dist, 0.0, 1.0).astype(np.float32)


def apply_rope(obs: np.ndarray, dn: np.ndarray, inv_freq: np.ndarray, rotate_dim: int) -> np.ndarray:
    """rope_embed.py:44-62"""
    obs = np.asarray(obs, dtype=np.float32)
    N = obs.shape[0]
    pair = obs[:, :rotate_dim].reshape(N, -1, 2)
    theta = 2 * np.pi * dn[:, None] * inv_freq[None, :]
    s, c = np.sin(theta)[..., None], np.cos(theta)[..., None]
    x, y = pair[..., 0:1], pair[..., 1:2]
    rot = np.concatenate([x * c - y * s, x * s + y * c], axis=-1)
    out = obs.copy()
    out[:, :rotate_dim] = rot.reshape(N, rotate_dim)
    return out


def rope(obs, rotate_dim, max_dist=100.0, base=None, ego_idx=0):
    inv = rope_inv_freq(rotate_dim, base or max_dist)
    return apply_rope(obs, dist_norm(obs, max_dist, True, ego_idx), inv, rotate_dim)


def distpe(obs, d_embed, max_dist=100.0, base=None, use_euclidean=True, ego_idx=0, freqs=None):
    """dist_embed.py:76-96"""
    obs = np.asarray(obs, dtype=np.float32)
    f = dist_freqs(d_embed, base or max_dist) if freqs is None else freqs
    nd = dist_norm(obs, max_dist, use_euclidean, ego_idx)[:, None]
    ang = 2 * np.pi * nd * f
    return np.concatenate([obs, np.sin(ang), np.cos(ang)], axis=-1).astype(np.float32)


def rankpe(obs, tag):
    """rank_embed.py:45-51 with embed = tanh(table).detach()"""
    return np.concatenate([np.asarray(obs, dtype=np.float32), np.asarray(tag, dtype=np.float32)], axis=-1)
