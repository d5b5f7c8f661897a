"""Configuration plumbing of the reference's plugin mechanism (seam 3): hydra `compose` of a YAML with
`defaults:` inheritance and `${...}` interpolation, then `_target_` / `_partial_` instantiation
(reference actionmesh/utils.py:45-74, actionmesh/pipeline.py:98-110,164-167).

With hydra-core installed `load_config` is the reference's own code path (initialize_config_dir + compose + resolve).
Without it (this image has neither hydra-core nor omegaconf) a small loader covers exactly the subset the reference's
configs use: a `defaults:` list of sibling files, `${a.b.c}` interpolation (whole-value interpolations keep their type),
dotted-key updates.  `instantiate` resolves `_target_` strings the way hydra.utils.instantiate does for these files.
"""
from __future__ import annotations

import functools
import importlib
import os
import re
from typing import Any, Optional

DEFAULT_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


class Node(dict):
    """dict with attribute access (the part of OmegaConf's DictConfig the pipeline uses: cfg.a.b, cfg.a.b = v)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = value


def _wrap(obj):
    if isinstance(obj, dict):
        return Node({k: _wrap(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [_wrap(v) for v in obj]
    return obj


def _merge(base: dict, over: dict) -> dict:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _read_yaml(config_name: str, config_dir: str) -> dict:
    import yaml

    name = config_name if config_name.endswith((".yaml", ".yml")) else config_name + ".yaml"
    with open(os.path.join(config_dir, name)) as f:
        raw = yaml.safe_load(f) or {}
    merged: dict = {}
    for d in raw.pop("defaults", []) or []:
        if isinstance(d, str) and d != "_self_":
            merged = _merge(merged, _read_yaml(d, config_dir))
    return _merge(merged, raw)


_INTERP = re.compile(r"\$\{([^${}]+)\}")


def _lookup(root: dict, dotted: str):
    cur: Any = root
    for part in dotted.strip().split("."):
        cur = cur[int(part)] if isinstance(cur, list) else cur[part]
    return cur


def _resolve(node, root, depth=0):
    if depth > 16:
        raise ValueError("interpolation cycle in config")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        m = _INTERP.fullmatch(node)
        if m:  # the whole value is one interpolation: keep the referenced value's type
            return _resolve(_lookup(root, m.group(1)), root, depth + 1)
        if _INTERP.search(node):
            return _INTERP.sub(lambda mm: str(_resolve(_lookup(root, mm.group(1)), root, depth + 1)), node)
    return node


def load_config(config_name: str, config_dir: Optional[str] = None, updates: Optional[dict] = None):
    """Reference `load_config(config_name, config_dir, updates)` (actionmesh/utils.py:45-74)."""
    config_dir = os.path.abspath(config_dir or DEFAULT_CONFIG_DIR)
    updates = updates or {}
    try:  # the reference's own path when hydra is installed
        from hydra import compose, initialize_config_dir
        from omegaconf import OmegaConf
    except ImportError:
        raw = _read_yaml(config_name, config_dir)
        for k, v in updates.items():
            cur = raw
            parts = k.split(".")
            for part in parts[:-1]:
                cur = cur.setdefault(part, {})
            cur[parts[-1]] = v
        return _wrap(_resolve(raw, raw))
    with initialize_config_dir(config_dir=config_dir, version_base="1.1", job_name="load_config"):
        cfg = compose(config_name=config_name, return_hydra_config=False,
                      overrides=["hydra.output_subdir=null", "hydra.job.chdir=false", "hydra/job_logging=none",
                                 "hydra/hydra_logging=none"])
        for k, v in updates.items():
            OmegaConf.update(cfg, k, v)
        OmegaConf.resolve(cfg)
    return cfg


def get_target(path: str):
    """'pkg.module.Class' -> the object."""
    module, _, name = path.rpartition(".")
    return getattr(importlib.import_module(module), name)


def instantiate(node, **overrides):
    """hydra.utils.instantiate for the subset in use: `_target_` (+ `_partial_: true` -> functools.partial).  Nested nodes
    with their own `_target_` are instantiated recursively; `_convert_` is accepted and ignored (plain python containers)."""
    overrides.pop("_convert_", None)
    try:
        from hydra.utils import instantiate as hydra_instantiate
        from omegaconf import DictConfig

        if isinstance(node, DictConfig):
            return hydra_instantiate(node, _convert_="partial", **overrides)
    except ImportError:
        pass
    kwargs = {}
    for k, v in dict(node).items():
        if k in ("_target_", "_partial_", "_convert_", "_recursive_"):
            continue
        kwargs[k] = instantiate(v) if isinstance(v, dict) and "_target_" in v else (list(v) if isinstance(v, list) else v)
    kwargs.update(overrides)
    target = get_target(node["_target_"])
    if node.get("_partial_", False):
        return functools.partial(target, **kwargs)
    return target(**kwargs)
