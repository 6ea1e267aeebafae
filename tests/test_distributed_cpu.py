"""world_size-2 gloo tests (CPU) of the sharding logic: the sum over ranks of the locally scaled gradients is the
global minibatch gradient, the all-reduced advantage moments give the whole-buffer mean / unbiased std of
ppo/agent.py:204, and env shards tile the global env ids."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from highway_rope_ppo_b200.ppo import distributed as D
    from oracle import ppo_ref

    torch.set_num_threads(1)
    S, A, H, B = 12, 2, 16, 64
    g = torch.Generator().manual_seed(0)                      # identical data on every rank
    flat = torch.randn(sum(int(np.prod(s)) for _, s in ppo_ref.shapes(S, A, H)), generator=g) * 0.2
    x, z = torch.randn(B, S, generator=g), torch.randn(B, A, generator=g)
    old, adv, ret = torch.randn(B, generator=g) - 2, torch.randn(B, generator=g), torch.randn(B, generator=g)
    full = ppo_ref.loss_and_grad(flat, x, z, old, adv, ret, S, A, H)
    base, per = D.shard_envs(B, D.rank(), D.world_size())
    sl = slice(base, base + per)
    local = ppo_ref.loss_and_grad(flat, x[sl], z[sl], old[sl], adv[sl], ret[sl], S, A, H)
    # the kernel weights every sample by loss_scale = 1/(B_local * world); the oracle returns the local MEAN gradient
    grad = local["grad"] * (D.loss_scale(per, D.world_size()) * per)
    D.allreduce_sum_(grad)
    loss = torch.tensor([local["loss"] * D.loss_scale(per, D.world_size()) * per], dtype=torch.float64)
    D.allreduce_sum_(loss)
    # advantage moments
    buf = torch.randn(1000, generator=g, dtype=torch.float64) * 3 + 1
    b2, p2 = D.shard_envs(1000, D.rank(), D.world_size())
    mine = buf[b2:b2 + p2]
    stats = torch.stack([mine.sum(), (mine * mine).sum(), torch.tensor(float(p2), dtype=torch.float64)])
    D.allreduce_sum_(stats)
    mean, std = D.global_mean_std(stats)
    torch.save({"grad": grad, "full_grad": full["grad"], "loss": float(loss), "full_loss": full["loss"],
                "mean": float(mean), "std": float(std), "want_mean": float(buf.mean()), "want_std": float(buf.std()),
                "shard": (base, per)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gradient_and_moment_reduction(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    for o in outs:
        np.testing.assert_allclose(o["grad"].numpy(), o["full_grad"].numpy(), atol=1e-6, rtol=1e-5)
        assert abs(o["loss"] - o["full_loss"]) < 1e-6
        assert abs(o["mean"] - o["want_mean"]) < 1e-12 and abs(o["std"] - o["want_std"]) < 1e-12
    assert [o["shard"] for o in outs] == [(0, 32), (32, 32)]
    assert torch.equal(outs[0]["grad"], outs[1]["grad"])  # every rank applies the same update


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from highway_rope_ppo_b200.ppo import distributed as D

    assert [D.shard_envs(65536, r, 8) for r in (0, 3, 7)] == [(0, 8192), (24576, 8192), (57344, 8192)]
    with pytest.raises(ValueError):
        D.shard_envs(10, 0, 3)
    assert D.world_size() == 1 and D.rank() == 0 and D.loss_scale(64, 1) == 1 / 64
    t = torch.ones(3)
    assert D.allreduce_sum_(t) is t


def test_oracle_spawn_is_shard_invariant(highway_config):
    """Global env id -> episode: two shards of 4 envs reproduce the single 8-env run (oracle side of the property
    the GPU test test_reset_matches_oracle checks for the kernels with env_id_base)."""
    from oracle import highway as oh

    e = oh.OracleEnv(highway_config)
    whole = []
    for i in range(8):
        e.reset(7, env_id=i, episode=2)
        whole.append(e.get_state()["x"].copy())
    for base in (0, 4):
        for j in range(4):
            e.reset(7, env_id=base + j, episode=2)
            assert np.array_equal(e.get_state()["x"], whole[base + j])
