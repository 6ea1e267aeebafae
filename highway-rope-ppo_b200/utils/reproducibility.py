"""Seeding helpers (reference: ``utils/reproducibility.py:10-35``)."""
import random

import numpy as np
import torch

SEED = 42


def set_random_seeds(seed: int = SEED, exact_reproducibility: bool = False) -> None:
    """Seed python, numpy and torch.  The torch seed fixes the PPO network init and the
    RankPE table, the numpy seed fixes the minibatch partition of ``PPOAgent.update``."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = bool(exact_reproducibility)
    torch.backends.cudnn.benchmark = not exact_reproducibility


def get_device():
    """Best available device and a label for logs; this framework needs CUDA to compute."""
    if torch.cuda.is_available():
        return torch.device("cuda"), f"GPU: {torch.cuda.get_device_name(0)}"
    return torch.device("cpu"), "CPU"
