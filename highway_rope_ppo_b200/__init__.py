"""Importable alias of the ``highway-rope-ppo_b200/`` package directory.

The product lives in ``highway-rope-ppo_b200/`` (a name Python cannot import directly);
this shim puts that directory on the package path, so
``import highway_rope_ppo_b200.experiments.wrappers`` resolves to
``highway-rope-ppo_b200/experiments/wrappers.py``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "highway-rope-ppo_b200")
__path__.insert(0, _PKG_DIR)
PKG_DIR = _PKG_DIR

__all__ = ["PKG_DIR"]
