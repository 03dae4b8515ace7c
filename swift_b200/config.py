"""Minimal ``_target_`` instantiation and YAML group composition.

The reference selects classes through hydra (``hydra.utils.instantiate`` on ``_target_`` strings,
models/precond.py:123-131; groups composed by configs/train.yaml and ``# @package _global_`` experiment files).
hydra/omegaconf are not installed in this image, so this module provides the small subset the forecast path
needs; when real hydra is present its ``instantiate`` is used instead and behaves identically for these configs.
"""
from __future__ import annotations

import importlib
import os
from typing import Any, Dict

_META = ("_convert_", "_recursive_", "_partial_")


def instantiate(config: Dict[str, Any], **kwargs):
    """``hydra.utils.instantiate`` for flat configs: import ``_target_`` and call it with merged kwargs."""
    try:  # pragma: no cover - hydra is absent in this image
        from hydra.utils import instantiate as _hydra_instantiate
        return _hydra_instantiate(config, **kwargs)
    except ImportError:
        pass
    cfg = dict(config)
    cfg.update(kwargs)
    for k in _META:
        cfg.pop(k, None)
    if "_target_" not in cfg:
        raise ValueError("config has no _target_")
    mod, _, name = cfg.pop("_target_").rpartition(".")
    return getattr(importlib.import_module(mod), name)(**cfg)


def _merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def load_experiment(name: str, config_dir: str | None = None) -> Dict[str, Any]:
    """Compose ``experiment/<name>.yaml``: every ``- /group: option`` of its ``defaults`` list is loaded into
    ``cfg[group]`` and the experiment's own keys are merged on top (hydra's ``# @package _global_`` semantics)."""
    import yaml

    config_dir = config_dir or os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")
    with open(os.path.join(config_dir, "experiment", name + ".yaml")) as f:
        exp = yaml.safe_load(f)
    cfg: Dict[str, Any] = {}
    for item in exp.pop("defaults", []):
        if not isinstance(item, dict):
            continue
        for group, option in item.items():
            group = group.replace("override ", "").strip().lstrip("/")
            path = os.path.join(config_dir, group, f"{option}.yaml")
            if os.path.exists(path):
                with open(path) as f:
                    node = cfg
                    parts = group.split("/")
                    for p in parts[:-1]:
                        node = node.setdefault(p, {})
                    node[parts[-1]] = _merge(node.get(parts[-1], {}) or {}, yaml.safe_load(f) or {})
    return _merge(cfg, exp)
