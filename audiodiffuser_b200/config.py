"""Minimal `_target_` instantiation in the reference's Hydra config format.

The reference builds its whole object graph with `hydra.utils.instantiate(cfg.model)`
(src/train.py:57, configs/model/diffunet_complex.yaml:1-29). Hydra / OmegaConf are not part of this
image, so this module resolves the same `_target_` / `_partial_` keys with importlib; with Hydra
installed, the YAML files under configs/ work with `hydra.utils.instantiate` unchanged.
"""
import functools
import importlib
from typing import Any

import yaml


def locate(path: str):
    module, _, name = path.rpartition(".")
    if not module:
        raise ValueError(f"_target_ must be a dotted path, got {path!r}")
    return getattr(importlib.import_module(module), name)


def instantiate(cfg: Any, **overrides):
    """Recursively instantiate dicts carrying `_target_` (kwargs = remaining keys); `_partial_: true`
    returns functools.partial, like Hydra."""
    if isinstance(cfg, dict):
        if "_target_" in cfg:
            kwargs = {k: instantiate(v) for k, v in cfg.items() if k not in ("_target_", "_partial_")}
            kwargs.update(overrides)
            target = locate(cfg["_target_"])
            if cfg.get("_partial_", False):
                return functools.partial(target, **kwargs)
            return target(**kwargs)
        return {k: instantiate(v) for k, v in cfg.items()}
    if isinstance(cfg, list):
        return [instantiate(v) for v in cfg]
    if isinstance(cfg, str):
        low = cfg.lower()
        if low in ("inf", ".inf", "+inf"):
            return float("inf")
    return cfg


def load_yaml(path: str):
    with open(path) as f:
        return yaml.safe_load(f)
