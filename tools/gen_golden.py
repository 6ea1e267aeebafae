"""Generate tests/golden/*.npz by running the UNMODIFIED reference code from /root/reference.

Run once in the build container (the reference tree does not exist on the GPU box):

    python tools/gen_golden.py

* ``embed_*.npz``: outputs of the reference's RotaryEmbedWrapper / DistanceEmbedWrapper
  (``experiments/rope_embed.py``, ``experiments/dist_embed.py``) on seeded observations.  The
  wrappers import ``gymnasium``, which is absent here, so a ~25-line stand-in providing only
  ``ObservationWrapper``/``Env``/``spaces.Box`` is put on ``sys.modules`` first; the wrapper code
  itself is the reference's.  RankEmbedWrapper.observation() raises at HEAD, so its
  fixture holds ``tanh(table).detach()`` computed from the reference-constructed table.
* ``ppo_*.npz``: ``ppo/agent.py`` imported as is: initial parameters for a seed, one evaluate +
  loss + backward + clip + Adam step on a seeded minibatch, GAE on a seeded trajectory, and a
  full ``PPOAgent.update`` (all metrics + final parameters).  ``ppo_wide_*.npz`` / ``ppo_wideupdate_*.npz``: the
  same at the swept and benchmarked widths (hidden_dim 128-512, state_dim 60-600, batch 32-4096), with the
  parameter-sized vectors stored in compact form (per-tensor norms, random projections, samples).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def install_gymnasium_stub():
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            shape = tuple(shape if shape is not None else np.shape(low))
            self.shape, self.dtype = shape, np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()

    class Env:
        def __init__(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env
            self.observation_space = env.observation_space
            self.action_space = getattr(env, "action_space", None)

    class ObservationWrapper(Wrapper):
        pass

    spaces.Box = Box
    gym.spaces, gym.Env, gym.Wrapper, gym.ObservationWrapper = spaces, Env, Wrapper, ObservationWrapper
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces
    return gym


def dummy_env(gym, shape):
    e = gym.Env()
    e.observation_space = gym.spaces.Box(-np.inf, np.inf, shape, np.float32)
    e.action_space = gym.spaces.Box(-1.0, 1.0, (2,), np.float32)
    return e


def highway_like_obs(rng, N, F, n_real):
    """Observation with the statistics of the normalised Kinematics table."""
    obs = np.zeros((N, F), dtype=np.float32)
    obs[0, 0] = 1.0
    obs[0, 1] = rng.choice([0.0, 0.04, 0.08, 0.12])
    obs[0, 2:4] = [rng.uniform(0.6, 1.0), rng.uniform(-0.05, 0.05)]
    for r in range(1, n_real):
        obs[r, 0] = rng.uniform(-0.1, 1.0)
        obs[r, 1] = rng.uniform(-0.12, 0.12)
        obs[r, 2] = rng.uniform(-0.3, 0.1)
        obs[r, 3] = rng.uniform(-0.05, 0.05)
        if F > 4:
            obs[r, 4:] = rng.uniform(-1, 1, F - 4)
    perm = rng.permutation(N - 1)
    obs[1:] = obs[1:][perm]
    return obs


def gen_embed(gym):
    sys.path.insert(0, REF)
    from experiments.dist_embed import DistanceEmbedWrapper
    from experiments.rank_embed import RankEmbedWrapper
    from experiments.rope_embed import RotaryEmbedWrapper

    rng = np.random.default_rng(20261018)
    cases = {}
    # RoPE: the make_env-reachable shape (N=15, F=4, rotate 4), generic shapes incl. rotate_dim 16
    for name, (N, F, rd, md) in {"rope_n15_f4_r4": (15, 4, 4, 100.0), "rope_n15_f4_r2": (15, 4, 2, 100.0),
                                 "rope_n30_f7_r6": (30, 7, 6, 100.0), "rope_n5_f16_r16": (5, 16, 16, 10.0),
                                 "rope_n8_f20_r16": (8, 20, 16, 1.0)}.items():
        w = RotaryEmbedWrapper(dummy_env(gym, (N, F)), rotate_dim=rd, max_dist=md)
        B = 16
        if F in (4, 7):
            obs = np.stack([highway_like_obs(rng, N, F, int(rng.integers(1, N + 1))) for _ in range(B)])
        else:
            obs = rng.standard_normal((B, N, F)).astype(np.float32)
        out = np.stack([w.observation(o) for o in obs])
        dn = rng.random((B, N)).astype(np.float32)
        out_dn = np.stack([w._apply_rope(o.copy(), d) for o, d in zip(obs, dn)])
        cases[name] = dict(obs=obs, out=out, dist_norm=dn, out_dist_norm=out_dn, inv_freq=w.inv_freq,
                           rotate_dim=rd, max_dist=md)
    for name, (N, F, d, md, eu) in {"dist_n15_f4_d4": (15, 4, 4, 100.0, True), "dist_n15_f4_d16": (15, 4, 16, 100.0, True),
                                    "dist_n30_f4_d8": (30, 4, 8, 100.0, True), "dist_n15_f4_d8_abs": (15, 4, 8, 100.0, False),
                                    "dist_n6_f5_d6_md1": (6, 5, 6, 1.0, True)}.items():
        w = DistanceEmbedWrapper(dummy_env(gym, (N, F)), d_embed=d, max_dist=md, use_euclidean=eu)
        B = 16
        if md == 100.0:
            obs = np.stack([highway_like_obs(rng, N, F, int(rng.integers(1, N + 1))) for _ in range(B)])
        else:
            obs = rng.standard_normal((B, N, F)).astype(np.float32)
        out = np.stack([w.observation(o) for o in obs])
        cases[name] = dict(obs=obs, out=out, freqs=w._freqs_np, d_embed=d, max_dist=md, use_euclidean=int(eu))
    for name, (N, F, d, seed) in {"rank_n15_f4_d4": (15, 4, 4, 42), "rank_n30_f4_d16": (30, 4, 16, 1042)}.items():
        torch.manual_seed(seed)
        w = RankEmbedWrapper(dummy_env(gym, (N, F)), d_embed=d)
        tag = torch.tanh(w.table.weight).detach().numpy()  # intended value of rank_embed.py:48
        obs = np.stack([highway_like_obs(rng, N, F, int(rng.integers(1, N + 1))) for _ in range(4)])
        out = np.concatenate([obs, np.broadcast_to(tag, (4,) + tag.shape)], axis=-1).astype(np.float32)
        cases[name] = dict(obs=obs, out=out, tag=tag, d_embed=d, seed=seed)
    for name, arrs in cases.items():
        np.savez_compressed(os.path.join(OUT, f"embed_{name}.npz"), **arrs)
    print("embed fixtures:", sorted(cases))


def flat_params(ac):
    return np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in ac.parameters()])


def gen_ppo():
    sys.path.insert(0, REF)
    from ppo.agent import PPOAgent
    import torch.nn as nn
    import torch.nn.functional as Fn

    torch.set_num_threads(1)
    for name, (S, A, H, B, seed) in {"s60_h256_b64": (60, 2, 256, 64, 42), "s20_h32_b17": (20, 2, 32, 17, 7),
                                     "s300_h64_b256": (300, 2, 64, 256, 3)}.items():
        torch.manual_seed(seed)
        np.random.seed(seed)
        agent = PPOAgent(S, A, lr=3e-4, hidden_dim=H, batch_size=B, epochs=2)
        ac = agent.actor_critic
        names = [n for n, _ in ac.named_parameters()]
        p0 = flat_params(ac)
        g = torch.Generator().manual_seed(seed + 1)
        states = torch.randn(B, S, generator=g) * 0.5
        mean, std, value = ac.forward(states)
        noise = torch.randn(B, A, generator=g)
        z = (mean + std * noise).detach()
        z[0, 0] = 4.0  # saturated tanh exercises the 1e-6 guard of the log-prob correction
        logp, v, ent = ac.evaluate(states, torch.tanh(z), z)
        old_logp = (logp + 0.3 * torch.randn(B, generator=g)).detach()  # spread ratios across the clip range
        adv = torch.randn(B, generator=g)
        ret = torch.randn(B, generator=g)
        # one minibatch of PPOAgent.update (agent.py:218-252)
        new_logp, sv, entropy = ac.evaluate(states, torch.tanh(z), z)
        ratios = torch.exp(new_logp - old_logp)
        surr1 = ratios * adv
        surr2 = torch.clamp(ratios, 1 - agent.eps_clip, 1 + agent.eps_clip) * adv
        actor_loss = -torch.min(surr1, surr2).mean()
        critic_loss = Fn.mse_loss(sv.squeeze(-1), ret)
        loss = actor_loss + agent.value_coef * critic_loss - agent.entropy_coef * entropy.mean()
        agent.optimizer.zero_grad()
        loss.backward()
        grads = np.concatenate([p.grad.detach().numpy().reshape(-1) for p in ac.parameters()])
        total_norm = float(nn.utils.clip_grad_norm_(ac.parameters(), agent.max_grad_norm))
        agent.optimizer.step()
        p1 = flat_params(ac)
        # a second step (bias correction with step = 2)
        new_logp, sv, entropy = ac.evaluate(states, torch.tanh(z), z)
        ratios2 = torch.exp(new_logp - old_logp)
        loss2 = (-torch.min(ratios2 * adv, torch.clamp(ratios2, 0.8, 1.2) * adv).mean()
                 + 0.5 * Fn.mse_loss(sv.squeeze(-1), ret) - 0.005 * entropy.mean())
        agent.optimizer.zero_grad()
        loss2.backward()
        nn.utils.clip_grad_norm_(ac.parameters(), agent.max_grad_norm)
        agent.optimizer.step()
        p2 = flat_params(ac)
        kl = float(((ratios - 1) - (new_logp * 0 + torch.log(ratios))).mean())
        np.savez_compressed(
            os.path.join(OUT, f"ppo_step_{name}.npz"), names=np.array(names), params0=p0, states=states.numpy(),
            mean=mean.detach().numpy(), value=value.detach().numpy(), pre_tanh=z.numpy(), logp=logp.detach().numpy(),
            entropy=ent.detach().numpy(), old_logp=old_logp.numpy(), adv=adv.numpy(), ret=ret.numpy(),
            loss=float(loss), actor_loss=float(actor_loss), critic_loss=float(critic_loss),
            clip_fraction=float((torch.abs(ratios - 1) > agent.eps_clip).float().mean()), approx_kl=kl,
            grads=grads, total_norm=total_norm, params1=p1, params2=p2, loss2=float(loss2),
            dims=np.array([S, A, H, B, seed]))
    # GAE (agent.py:126-138) on seeded trajectories with episode boundaries
    from ppo.agent import PPOMemory
    rng = np.random.default_rng(5)
    for name, T in {"t50": 50, "t2048": 2048}.items():
        mem = PPOMemory()
        mem.rewards = [float(x) for x in rng.random(T)]
        mem.values = [np.float32(x) for x in rng.standard_normal(T)]
        mem.dones = [bool(x) for x in (rng.random(T) < 0.05)]
        last_value = 0.37
        adv, ret = mem.compute_advantages(0.99, 0.95, last_value)
        np.savez_compressed(os.path.join(OUT, f"ppo_gae_{name}.npz"), reward=np.array(mem.rewards, dtype=np.float64),
                            value=np.array(mem.values, dtype=np.float32), done=np.array(mem.dones),
                            last_value=last_value, adv=adv, ret=ret)
    # a full update(): 2048 stored transitions, bs 64, 2 epochs (the reference's loop shape)
    for name, (S, H, n, bs, epochs, seed) in {"s60_h256_n2048": (60, 256, 2048, 64, 2, 42),
                                              "s12_h16_n100": (12, 16, 100, 32, 3, 9)}.items():
        torch.manual_seed(seed)
        np.random.seed(seed)
        agent = PPOAgent(S, 2, lr=3e-4, hidden_dim=H, batch_size=bs, epochs=epochs)
        p0 = flat_params(agent.actor_critic)
        rng = np.random.default_rng(seed)
        st = (rng.standard_normal((n, S)) * 0.5).astype(np.float32)
        rew = rng.random(n)
        done = rng.random(n) < 0.03
        stored = {k: [] for k in ("action", "pre_tanh", "logp", "value")}
        for t in range(n):
            a, z, lp, v = agent.select_action(st[t])
            agent.memory.store(st[t], a, z, float(rew[t]), st[t], lp, bool(done[t]), v)
            for k, x in zip(("action", "pre_tanh", "logp", "value"), (a, z, lp, v)):
                stored[k].append(x)
        np.random.seed(seed + 100)  # fixes the minibatch permutation of get_batches()
        metrics = agent.update(last_value=0.25)
        np.savez_compressed(
            os.path.join(OUT, f"ppo_update_{name}.npz"), params0=p0, params1=flat_params(agent.actor_critic),
            states=st, reward=rew, done=done, action=np.array(stored["action"], dtype=np.float32),
            pre_tanh=np.array(stored["pre_tanh"], dtype=np.float32), logp=np.array(stored["logp"], dtype=np.float32),
            value=np.array(stored["value"], dtype=np.float32), last_value=0.25,
            metric_names=np.array(list(metrics)), metric_values=np.array([metrics[k] for k in metrics]),
            dims=np.array([S, 2, H, n, bs, epochs, seed + 100]))
    print("ppo fixtures written")


def compact(vec, sizes):
    """Compact, size-independent description of a flat parameter-shaped vector (the swept widths have up to
    1.2 M parameters; storing them in full would put tens of MB under tests/golden): the L2 norm of every
    parameter tensor, 8 seeded float64 random projections, and every 257th entry."""
    v = np.asarray(vec, dtype=np.float64)
    rng = np.random.default_rng(1234)
    proj = np.array([float(rng.standard_normal(v.size) @ v) for _ in range(8)])
    norms, off = [], 0
    for n in sizes:
        norms.append(float(np.sqrt((v[off:off + n] ** 2).sum())))
        off += n
    return {"norms": np.array(norms), "proj": proj, "samples": np.asarray(vec, dtype=np.float32)[::257].copy(),
            "sum": float(v.sum()), "sumsq": float((v ** 2).sum())}


def gen_ppo_wide():
    """ppo_wide_*.npz: the same one-minibatch fixture as ppo_step_* at the widths the reference sweeps
    (main.py:50-58: hidden_dim 128 / 256 / 384, batch 32 / 64) and BASELINE.json benchmarks (hidden_dim 512;
    state_dim 240 / 360 / 600 = 30 rows of F + d_embed), with parameters and gradients in compact() form; the
    initial parameters are regenerated by the test from the seed and checked against their compact form."""
    sys.path.insert(0, REF)
    from ppo.agent import PPOAgent
    import torch.nn as nn
    import torch.nn.functional as Fn

    torch.set_num_threads(8)
    cases = {"s60_h128_b32": (60, 128, 32, 11), "s240_h384_b64": (240, 384, 64, 12), "s600_h512_b64": (600, 512, 64, 13),
             "s60_h512_b32": (60, 512, 32, 14), "s600_h128_b64": (600, 128, 64, 15), "s360_h256_b32": (360, 256, 32, 16),
             "s240_h256_b64": (240, 256, 64, 17), "s60_h384_b64": (60, 384, 64, 18),
             "s60_h512_b4096": (60, 512, 4096, 19), "s600_h256_b512": (600, 256, 512, 20)}
    for name, (S, H, B, seed) in cases.items():
        A = 2
        torch.manual_seed(seed)
        np.random.seed(seed)
        agent = PPOAgent(S, A, lr=3e-4, hidden_dim=H, batch_size=B, epochs=2)
        ac = agent.actor_critic
        sizes = [p.numel() for p in ac.parameters()]
        p0 = flat_params(ac)
        g = torch.Generator().manual_seed(seed + 1)
        states = torch.randn(B, S, generator=g) * 0.5
        mean, std, value = ac.forward(states)
        noise = torch.randn(B, A, generator=g)
        z = (mean + std * noise).detach()
        z[0, 0] = 4.0
        logp, v, ent = ac.evaluate(states, torch.tanh(z), z)
        old_logp = (logp + 0.3 * torch.randn(B, generator=g)).detach()
        adv = torch.randn(B, generator=g)
        ret = torch.randn(B, generator=g)
        new_logp, sv, entropy = ac.evaluate(states, torch.tanh(z), z)
        ratios = torch.exp(new_logp - old_logp)
        actor_loss = -torch.min(ratios * adv, torch.clamp(ratios, 1 - agent.eps_clip, 1 + agent.eps_clip) * adv).mean()
        critic_loss = Fn.mse_loss(sv.squeeze(-1), ret)
        loss = actor_loss + agent.value_coef * critic_loss - agent.entropy_coef * entropy.mean()
        agent.optimizer.zero_grad()
        loss.backward()
        grads = np.concatenate([p.grad.detach().numpy().reshape(-1) for p in ac.parameters()])
        total_norm = float(nn.utils.clip_grad_norm_(ac.parameters(), agent.max_grad_norm))
        agent.optimizer.step()
        p1 = flat_params(ac)
        out = {"dims": np.array([S, A, H, B, seed]), "sizes": np.array(sizes), "states": states.numpy(),
               "mean": mean.detach().numpy(), "value": value.detach().numpy(), "pre_tanh": z.numpy(),
               "logp": logp.detach().numpy(), "old_logp": old_logp.numpy(), "adv": adv.numpy(), "ret": ret.numpy(),
               "loss": float(loss), "actor_loss": float(actor_loss), "critic_loss": float(critic_loss),
               "clip_fraction": float((torch.abs(ratios - 1) > agent.eps_clip).float().mean()),
               "total_norm": total_norm, "grad_absmax": float(np.abs(grads).max())}
        for tag, vec in (("params0", p0), ("grads", grads), ("step1", p1 - p0)):
            for k, x in compact(vec, sizes).items():
                out[f"{tag}_{k}"] = x
        np.savez_compressed(os.path.join(OUT, f"ppo_wide_{name}.npz"), **out)
    # a full update() at hidden_dim 512 (BASELINE configs[3]): 512 stored transitions, bs 64, 2 epochs
    S, H, n, bs, epochs, seed = 60, 512, 512, 64, 2, 77
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = PPOAgent(S, 2, lr=3e-4, hidden_dim=H, batch_size=bs, epochs=epochs)
    sizes = [p.numel() for p in agent.actor_critic.parameters()]
    p0 = flat_params(agent.actor_critic)
    rng = np.random.default_rng(seed)
    st = (rng.standard_normal((n, S)) * 0.5).astype(np.float32)
    rew = rng.random(n)
    done = rng.random(n) < 0.03
    stored = {k: [] for k in ("action", "pre_tanh", "logp", "value")}
    for t in range(n):
        a, z, lp, v = agent.select_action(st[t])
        agent.memory.store(st[t], a, z, float(rew[t]), st[t], lp, bool(done[t]), v)
        for k, x in zip(("action", "pre_tanh", "logp", "value"), (a, z, lp, v)):
            stored[k].append(x)
    np.random.seed(seed + 100)
    metrics = agent.update(last_value=0.25)
    p1 = flat_params(agent.actor_critic)
    out = {"dims": np.array([S, 2, H, n, bs, epochs, seed + 100, seed]), "sizes": np.array(sizes), "states": st, "reward": rew,
           "done": done, "action": np.array(stored["action"], dtype=np.float32),
           "pre_tanh": np.array(stored["pre_tanh"], dtype=np.float32), "logp": np.array(stored["logp"], dtype=np.float32),
           "value": np.array(stored["value"], dtype=np.float32), "last_value": 0.25,
           "metric_names": np.array(list(metrics)), "metric_values": np.array([metrics[k] for k in metrics])}
    for tag, vec in (("params0", p0), ("motion", p1 - p0)):
        for k, x in compact(vec, sizes).items():
            out[f"{tag}_{k}"] = x
    np.savez_compressed(os.path.join(OUT, "ppo_wideupdate_s60_h512_n512.npz"), **out)
    print("wide ppo fixtures written")


def gen_sweep():
    """Experiment names of the reference's full grid (main.py:42-88).  main.py builds a DevicePool and an
    ExperimentRunner (gymnasium + highway_env) when imported, so define_experiments is executed from its source
    text against the reference's own experiments.config."""
    import json
    import re

    from experiments.config import CommonHP, Condition, ConditionHP, Experiment, expand_condition_hps

    src = open(os.path.join(REF, "main.py")).read()
    fn = re.search(r"def define_experiments\(.*?\n    return experiments\n", src, re.S).group(0)
    seed = int(re.search(r"^SEED\s*=\s*(\d+)", open(os.path.join(REF, "utils", "reproducibility.py")).read(), re.M).group(1))
    ns = dict(Experiment=Experiment, Condition=Condition, ConditionHP=ConditionHP, CommonHP=CommonHP,
              expand_condition_hps=expand_condition_hps, SEED=seed)
    exec(fn, ns)
    ex = ns["define_experiments"](seed, 3)
    keys = ("lr", "hidden_dim", "clip_eps", "entropy_coef", "epochs", "batch_size", "d_embed", "gamma", "lam",
            "value_coef", "max_grad_norm", "steps_per_update")
    out = {"seed": seed, "count": len(ex), "names": [e.name for e in ex],
           "first_hp": {k: getattr(ex[0].hp, k) for k in keys}, "seeds": sorted(set(e.seed for e in ex))}
    json.dump(out, open(os.path.join(OUT, "sweep_experiments.json"), "w"))
    print("sweep fixture written:", len(ex), "experiments")


def gen_result_schema():
    """What the reference's offline tools expect of a run's artifacts, extracted from its sources: the run-name
    regular expression of results.py:33-44 (analysis.py:21-32 holds the same one), the keys of the metrics JSON
    (training/routine.py:88-97) and the header of the summary CSV (training/routine.py:285)."""
    import ast
    import json
    import re

    res = open(os.path.join(REF, "results.py")).read()
    tree = ast.parse(res)
    pattern = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "EXP_RX":
            pattern = ast.literal_eval(node.value.args[0])
    assert pattern, "EXP_RX not found in results.py"
    rt = open(os.path.join(REF, "training", "routine.py")).read()
    block = re.search(r"metrics_history = \{(.*?)\n    \}", rt, re.S).group(1)
    keys = re.findall(r'"(\w+)":', block)
    header = re.search(r'f\.write\("(experiment,[^"]*)\\n"\)', rt).group(1)
    used = sorted(set(re.findall(r'data\["(\w+)"\]', res)))
    json.dump({"exp_rx": pattern, "metrics_keys": keys, "summary_header": header, "keys_read_by_results_py": used},
              open(os.path.join(OUT, "result_schema.json"), "w"), indent=1)
    print("result schema fixture written:", keys, header, used)


def gen_checkpoint():
    """A checkpoint written by the reference's PPOAgent.save after one update (agent.py:310-318), with the network
    outputs on fixed states: pins the checkpoint wire format (model + optimizer state dicts)."""
    import torch

    from ppo.agent import PPOAgent

    S, H, n = 12, 16, 64
    torch.manual_seed(7)
    np.random.seed(7)
    agent = PPOAgent(S, 2, lr=1e-3, epochs=2, batch_size=32, hidden_dim=H, device="cpu")
    rng = np.random.default_rng(7)
    st = (rng.standard_normal((n, S)) * 0.5).astype(np.float32)
    for t in range(n):
        a, z, lp, v = agent.select_action(st[t])
        agent.memory.store(st[t], a, z, float(rng.random()), st[t], lp, bool(rng.random() < 0.05), v)
    agent.update(last_value=0.0)
    path = os.path.join(OUT, "checkpoint_ref_s12_h16.pth")
    agent.save(path)
    with torch.no_grad():
        mean, std, value = agent.actor_critic(torch.from_numpy(st))
    opt = agent.optimizer.state_dict()
    np.savez_compressed(os.path.join(OUT, "checkpoint_ref_s12_h16_outputs.npz"), states=st, mean=mean.numpy(),
                        std=std.numpy(), value=value.numpy(), step=np.array(float(opt["state"][0]["step"])),
                        exp_avg_w1=opt["state"][1]["exp_avg"].numpy(), lr=np.array(opt["param_groups"][0]["lr"]))
    print("checkpoint fixture written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--ppo-wide-only" in sys.argv:
        gen_ppo_wide()
        sys.exit(0)
    if "--extra-only" in sys.argv:
        sys.path.insert(0, REF)
        install_gymnasium_stub()
        gen_sweep()
        gen_checkpoint()
        gen_result_schema()
        sys.exit(0)
    gym = install_gymnasium_stub()
    gen_embed(gym)
    gen_ppo()
    gen_ppo_wide()
    gen_sweep()
    gen_checkpoint()
    gen_result_schema()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
