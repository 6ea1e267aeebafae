"""GPU parity of the embedding kernels (fused epilogue + hrp_embed_apply) against the fixtures produced
by the reference's own wrapper classes (tests/golden/embed_*.npz) and the numpy oracle.

Tolerance: 2e-6 absolute on O(1) values -- the reference evaluates sin/cos with numpy's float32 routines, the
kernel with CUDA's sinf/cosf (<= 2 ulp); every other operation is the same float32 operation in the same order.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import embed as oe

pytestmark = pytest.mark.gpu
ATOL = 2e-6


def _apply(kind, obs, table, dim, max_dist, eu=True, ego=0, dn=None):
    from highway_rope_ppo_b200.envs.highway_vec import embed_apply

    out = embed_apply(kind, torch.from_numpy(obs).cuda(), torch.from_numpy(np.ascontiguousarray(table)).cuda(), dim,
                      max_dist, eu, ego, None if dn is None else torch.from_numpy(dn).cuda())
    return out.cpu().numpy()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "embed_rope_*.npz"))))
def test_rope_kernel_matches_reference(path):
    from highway_rope_ppo_b200._lib import EMBED_ROPE

    g = dict(np.load(path))
    rd, md = int(g["rotate_dim"]), float(g["max_dist"])
    out = _apply(EMBED_ROPE, g["obs"], g["inv_freq"], rd, md)
    np.testing.assert_allclose(out, g["out"], atol=ATOL * max(1.0, np.abs(g["obs"]).max()))
    out = _apply(EMBED_ROPE, g["obs"], g["inv_freq"], rd, md, dn=g["dist_norm"])
    np.testing.assert_allclose(out, g["out_dist_norm"], atol=ATOL * max(1.0, np.abs(g["obs"]).max()))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "embed_dist_*.npz"))))
def test_dist_kernel_matches_reference(path):
    from highway_rope_ppo_b200._lib import EMBED_DIST

    g = dict(np.load(path))
    out = _apply(EMBED_DIST, g["obs"], g["freqs"], int(g["d_embed"]), float(g["max_dist"]), bool(g["use_euclidean"]))
    assert out.shape == g["out"].shape
    F = g["obs"].shape[-1]
    assert np.array_equal(out[..., :F], g["obs"])
    np.testing.assert_allclose(out, g["out"], atol=ATOL)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "embed_rank_*.npz"))))
def test_rank_kernel_matches_reference(path):
    from highway_rope_ppo_b200._lib import EMBED_RANK

    g = dict(np.load(path))
    out = _apply(EMBED_RANK, g["obs"], g["tag"], int(g["d_embed"]), 100.0)
    assert np.array_equal(out, g["out"])


# ---- the reference's own property tests (tests/test_rope_wrapper.py:34-113), against this package's wrapper
class _DummyEnv:
    def __init__(self, shape):
        from highway_rope_ppo_b200.envs.spaces import Box

        self.observation_space = Box(-np.inf, np.inf, shape, np.float32)
        self.action_space = Box(-1.0, 1.0, (1,), np.float32)

    def reset(self, *, seed=None, options=None):
        return np.zeros(self.observation_space.shape, dtype=np.float32), {}

    def step(self, action):
        return np.zeros(self.observation_space.shape, dtype=np.float32), 0.0, True, False, {}


def _rope(shape, **kw):
    from highway_rope_ppo_b200.experiments.rope_embed import RotaryEmbedWrapper

    return RotaryEmbedWrapper(_DummyEnv(shape), **kw)


def test_shape_and_dtype_preserved():
    w = _rope((5, 4), rotate_dim=4, max_dist=10.0)
    obs = np.random.randn(5, 4).astype(np.float32)
    out = w.observation(obs)
    assert out.shape == obs.shape and out.dtype == np.float32


def test_identity_for_zero_distance():
    w = _rope((4, 4), rotate_dim=4, max_dist=1.0)
    obs = np.zeros((4, 4), dtype=np.float32)
    obs[:, 2:] = np.random.randn(4, 2).astype(np.float32)
    assert np.allclose(w.observation(obs), obs, atol=1e-6)


def test_rotation_changes_values_and_is_invertible():
    w = _rope((6, 4), rotate_dim=4, max_dist=1.0)
    obs = np.zeros((6, 4), dtype=np.float32)
    obs[:, 0] = 1.0
    wrapped = w._apply_rope(obs.copy(), np.ones(6, dtype=np.float32))
    assert not np.allclose(wrapped[:, :2], obs[:, :2])
    obs = np.random.randn(6, 4).astype(np.float32)
    dn = np.random.rand(6).astype(np.float32)
    back = w._apply_rope(w._apply_rope(obs.copy(), dn), -dn)
    assert np.allclose(back, obs, atol=1e-6)


def test_wrapper_step_and_reset_pass_through_observation():
    w = _rope((3, 4), rotate_dim=2, max_dist=1.0)
    obs, info = w.reset(seed=1)
    assert obs.shape == (3, 4) and info == {}
    obs, r, te, tr, _ = w.step(np.zeros(1, dtype=np.float32))
    assert obs.shape == (3, 4) and te and not tr


# ---- fused epilogue: make_env(...) == wrapper(observation of the plain env), same state
@pytest.mark.parametrize("cond_name,d", [("SHUFFLED_ROPE", 4), ("SHUFFLED_ROPE", 2), ("SHUFFLED_DISTPE", 4),
                                         ("SHUFFLED_RANKPE", 8)])
def test_fused_embedding_matches_oracle_on_sim_observations(highway_config, cond_name, d):
    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_vec_env

    cond = Condition[cond_name]
    over = {"observation": {"order": "shuffled"}}
    torch.manual_seed(3)
    plain = make_vec_env(Condition.SHUFFLED, highway_config, None, over, num_envs=64, seed=9)
    torch.manual_seed(3)
    fused = make_vec_env(cond, highway_config, d, over, num_envs=64, seed=9)
    a = plain.reset(9).cpu().numpy()
    b = fused.reset(9).cpu().numpy()
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for t in range(6):
        if t:
            act = torch.rand((64, 2), generator=g, device="cuda:0") * 2 - 1
            a = plain.step(act)[0].cpu().numpy()
            b = fused.step(act)[0].cpu().numpy()
        for e in range(64):
            if cond is Condition.SHUFFLED_ROPE:
                want = oe.rope(a[e], d, 100.0)
            elif cond is Condition.SHUFFLED_DISTPE:
                want = oe.distpe(a[e], d, 100.0)
            else:
                want = oe.rankpe(a[e], fused.embed.table.reshape(15, d))
            np.testing.assert_allclose(b[e], want, atol=ATOL)
    plain.close(); fused.close()


def test_dist_d16_direct_construction(highway_config):
    """DistanceEmbedWrapper built directly accepts d_embed 16 > F (visualize.py:147-152)."""
    from highway_rope_ppo_b200.experiments.config import Condition
    from highway_rope_ppo_b200.experiments.wrappers import make_vec_env

    with pytest.raises(ValueError):
        make_vec_env(Condition.SHUFFLED_DISTPE, highway_config, 16, num_envs=4)
    env = make_vec_env(Condition.SHUFFLED_DISTPE, highway_config, 16, num_envs=4, strict_d_embed=False)
    assert env.obs_shape == (15, 20)
    env.close()
