"""One experiment end to end (reference: ``experiments/runner.py:19-155``).

``ExperimentRunner.launch`` keeps the reference's result dictionary (``experiment_name``, ``status``
``COMPLETED`` / ``FAILED``, ``rewards``, ``avg_rewards``, ``metrics_history``, ``error_message``,
``error_traceback``, ``duration_seconds``) and its order of operations: seed -> env -> agent -> train, with
failures captured rather than raised.  The reference's ``DevicePool`` (GPU time-sharing by
``CUDA_VISIBLE_DEVICES`` round-robin, ``utils/device_pool.py:45-72``) is replaced by "one process per GPU":
the runner uses ``cuda:LOCAL_RANK`` unless a pool-like object with an ``acquire()`` context manager is passed.
"""
from __future__ import annotations

import contextlib
import logging
import os
import time
import traceback
from typing import Any, Dict, Optional

import numpy as np
import torch

from ..envs.spaces import is_box
from ..ppo.agent import PPOAgent
from ..training.routine import train_with_experiment_name
from ..utils.reproducibility import set_random_seeds
from .config import Experiment
from .wrappers import make_env


class ExperimentRunner:
    def __init__(self, base_env_config: dict, device_pool: Optional[Any] = None, artifacts_dir: Optional[str] = None):
        self.base_config = base_env_config
        self.pool = device_pool
        self.artifacts_dir = artifacts_dir

    @contextlib.contextmanager
    def _device(self):
        if self.pool is not None:
            with self.pool.acquire() as device:
                yield device
        else:
            yield torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))

    def _create_agent(self, state_dim: int, action_dim: int, hp, logger, device) -> PPOAgent:
        return PPOAgent(state_dim=state_dim, action_dim=action_dim, lr=hp.lr, gamma=hp.gamma, lam=hp.lam,
                        eps_clip=hp.clip_eps, value_coef=hp.value_coef, entropy_coef=hp.entropy_coef,
                        max_grad_norm=hp.max_grad_norm, epochs=hp.epochs, batch_size=hp.batch_size,
                        hidden_dim=hp.hidden_dim, logger=logger, device=device)

    def launch(self, exp: Experiment) -> Dict[str, Any]:
        result: Dict[str, Any] = {"experiment_name": exp.name, "status": "FAILED"}
        t0 = time.time()
        logger = logging.getLogger(f"experiment_{exp.name}")
        try:
            with self._device() as device:
                set_random_seeds(exp.seed)
                logger.info(f"[{exp.name}] device {device} | seed {exp.seed} | condition {exp.condition.name}")
                env = None
                try:
                    env = make_env(exp.condition, self.base_config, d_embed=exp.hp.d_embed,
                                   env_overrides=exp.env_config_overrides, device=device)
                    if hasattr(env, "to") and callable(env.to):
                        env = env.to(device)
                    if not is_box(env.observation_space):
                        raise TypeError(f"Unsupported observation space: {type(env.observation_space)}")
                    if not is_box(env.action_space):
                        raise TypeError(f"Unsupported action space: {type(env.action_space)}")
                    state_dim = int(np.prod(env.observation_space.shape))
                    action_dim = env.action_space.shape[0]
                    agent = self._create_agent(state_dim, action_dim, exp.hp, logger, device)
                    rewards, avg_rewards, metrics = train_with_experiment_name(
                        env=env, agent=agent, max_episodes=exp.max_episodes, target_reward=exp.target_reward,
                        log_interval=exp.extra.get("log_interval", 20), eval_interval=exp.extra.get("eval_interval", 50),
                        steps_per_update=exp.hp.steps_per_update, experiment_name=exp.name, exp_seed=exp.seed,
                        logger=logger, artifacts_dir=self.artifacts_dir)
                    result.update(status="COMPLETED", rewards=rewards, avg_rewards=avg_rewards,
                                  metrics_history=metrics)
                except Exception as e:  # the reference reports, it does not raise (runner.py:133-146)
                    logger.error(f"[{exp.name}] Experiment execution failed!", exc_info=True)
                    result["error_message"] = str(e)
                    result["error_traceback"] = traceback.format_exc()
                finally:
                    if env is not None:
                        env.close()
        except Exception as e:
            result["error_message"] = str(e)
            result["error_traceback"] = traceback.format_exc()
        result["duration_seconds"] = time.time() - t0
        return result


def experiment_name(condition_name: str, hp, seed: int, sweep_keys=("lr", "hidden_dim", "clip_eps", "entropy_coef",
                                                                 "epochs", "batch_size", "d_embed")) -> str:
    """The reference's run-name format (``main.py:77-87``): condition, then ``<key><value>`` for every swept
    hyper-parameter in sweep order, then the seed.  Its analysis scripts parse these names."""
    parts = [condition_name.lower()] + [f"{k}{getattr(hp, k)}" for k in sweep_keys] + [f"seed{seed}"]
    return "_".join(parts)
