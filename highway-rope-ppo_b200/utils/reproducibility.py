"""Seeding helpers (reference: ``utils/reproducibility.py:10-35``).

What each seed pins on this path: torch -> the nn.Linear initialisation of the PPO network and the RankPE table
(built on the CPU in the reference's construction order) and the Philox key of the in-kernel action noise;
numpy -> the minibatch permutation of ``PPOAgent.update``; the env seed is passed to ``reset`` explicitly."""
import random
from typing import Tuple

import numpy as np
import torch

SEED = 42


def set_random_seeds(seed: int = SEED, exact_reproducibility: bool = False) -> None:
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    # cuDNN is not on this path; the flags are kept for callers that share the process with other torch code
    torch.backends.cudnn.deterministic = bool(exact_reproducibility)
    torch.backends.cudnn.benchmark = not exact_reproducibility


def get_device() -> Tuple[torch.device, str]:
    """(device, label for logs).  The kernels of this framework need CUDA; a CPU answer only serves logging."""
    if not torch.cuda.is_available():
        return torch.device("cpu"), "CPU"
    return torch.device("cuda"), f"GPU: {torch.cuda.get_device_name(0)}"
