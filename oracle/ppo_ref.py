"""oracle/ppo_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain PyTorch fp32 (CPU, autograd) restatement of the reference's PPO arithmetic:
``ppo/agent.py:12-84`` (ActorCritic forward / evaluate), ``:126-138`` (GAE), ``:204`` (advantage
normalisation), ``:218-252`` (clipped surrogate + value MSE + entropy, backward, clip_grad_norm_,
Adam).  PINNED: ``tests/test_oracle_cpu.py`` checks it against ``tests/golden/ppo_*.npz`` which
``tools/gen_golden.py`` produced by running the reference's own ``ppo/agent.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.
"""
import math

import numpy as np
import torch

LOG_SQRT_2PI = 0.5 * math.log(2 * math.pi)


def shapes(S, A, H):
    return [("log_std", (A,)), ("shared.0.weight", (H, S)), ("shared.0.bias", (H,)),
            ("shared.2.weight", (H, H)), ("shared.2.bias", (H,)),
            ("actor_mean.0.weight", (H, H)), ("actor_mean.0.bias", (H,)),
            ("actor_mean.2.weight", (A, H)), ("actor_mean.2.bias", (A,)),
            ("critic.0.weight", (H, H)), ("critic.0.bias", (H,)),
            ("critic.2.weight", (1, H)), ("critic.2.bias", (1,))]


def unflatten(flat, S, A, H):
    out, off = {}, 0
    for name, shp in shapes(S, A, H):
        n = int(np.prod(shp))
        out[name] = flat[off:off + n].view(shp)
        off += n
    assert off == flat.numel()
    return out


def forward(flat, x, S, A, H):
    """agent.py:46-54"""
    p = unflatten(flat, S, A, H)
    h = torch.relu(x @ p["shared.0.weight"].T + p["shared.0.bias"])
    h = torch.relu(h @ p["shared.2.weight"].T + p["shared.2.bias"])
    a = torch.relu(h @ p["actor_mean.0.weight"].T + p["actor_mean.0.bias"])
    mean = a @ p["actor_mean.2.weight"].T + p["actor_mean.2.bias"]
    c = torch.relu(h @ p["critic.0.weight"].T + p["critic.0.bias"])
    value = c @ p["critic.2.weight"].T + p["critic.2.bias"]
    return mean, p["log_std"], value


def evaluate(flat, x, z, S, A, H):
    """agent.py:76-84: log-prob of pre-tanh z with the tanh correction, value, entropy"""
    mean, log_std, value = forward(flat, x, S, A, H)
    std = log_std.exp()
    logp = -((z - mean) ** 2) / (2 * std ** 2) - log_std - LOG_SQRT_2PI
    logp = logp - torch.log1p(-torch.tanh(z).pow(2) + 1e-6)
    ent = (0.5 + LOG_SQRT_2PI + log_std).sum().expand(x.shape[0])
    return logp.sum(-1), value, ent


def loss_and_grad(flat, x, z, old_logp, adv, ret, S, A, H, eps_clip=0.2, value_coef=0.5, entropy_coef=0.005):
    """agent.py:223-248; returns dict(loss, actor, critic, entropy, clip_fraction, approx_kl, grad)"""
    flat = flat.detach().clone().requires_grad_(True)
    logp, value, ent = evaluate(flat, x, z, S, A, H)
    ratio = torch.exp(logp - old_logp)
    s1, s2 = ratio * adv, torch.clamp(ratio, 1 - eps_clip, 1 + eps_clip) * adv
    actor = -torch.min(s1, s2).mean()
    critic = torch.nn.functional.mse_loss(value.squeeze(-1), ret)
    loss = actor + value_coef * critic - entropy_coef * ent.mean()
    loss.backward()
    with torch.no_grad():
        lr_ = logp - old_logp
        kl = ((torch.exp(lr_) - 1) - lr_).mean()
        cf = (torch.abs(ratio - 1) > eps_clip).float().mean()
    return dict(loss=float(loss.detach()), actor=float(actor.detach()), critic=float(critic.detach()),
                entropy=float(ent.mean().detach()),
                clip_fraction=float(cf), approx_kl=float(kl), grad=flat.grad.detach().clone())


def clip_adam(flat, grad, m, v, step, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8, max_norm=0.5):
    """agent.py:249-252: clip_grad_norm_(max_norm) then one torch.optim.Adam step (step is 1-based)"""
    total = torch.linalg.vector_norm(grad)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    g = grad * coef
    m = m.lerp(g, 1 - b1)
    v = v * b2 + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return flat - (lr / bc1) * (m / denom), m, v, float(total)


def gae(reward, value, done, last_value, gamma=0.99, lam=0.95):
    """agent.py:126-138: float64 arithmetic, float32 store of every advantage"""
    T = len(reward)
    values = np.concatenate([np.asarray(value, dtype=np.float64), [float(last_value)]])
    adv = np.zeros(T, dtype=np.float32)
    last = np.float32(0)
    for t in reversed(range(T)):
        nd = 1.0 - float(done[t])
        delta = float(reward[t]) + gamma * values[t + 1] * nd - values[t]
        adv[t] = delta + gamma * lam * nd * float(last)
        last = adv[t]
    return adv, adv + np.asarray(value, dtype=np.float32)


def normalize_adv(adv):
    """agent.py:204"""
    a = torch.as_tensor(adv, dtype=torch.float32)
    return (a - a.mean()) / (a.std() + 1e-8)
