"""Inference-time configuration surface of the reference ('Config / flags').

The reference reads the frozen per-checkpoint Hydra file `ckpts/<branch>/<cat>-num_more-3/.hydra/config.yaml`
through OmegaConf (eval.py:92,97) and uses five keys of it: res, num_more, up, right, front.  Hydra and
OmegaConf are not needed for that: PyYAML reads the same file.
"""
from __future__ import annotations

import os
from typing import Optional

INFERENCE_KEYS = ("res", "num_more", "up", "right", "front")


def default_category_cfg(category: str) -> dict:
    """Values every shipped checkpoint config carries (identical axes for all categories, res = 0.01 for
    laptop only -- SURVEY.md Appendix D)."""
    return dict(res=0.01 if category == "laptop" else 0.002, num_more=3, up=[0, 1, 0], right=[1, 0, 0], front=[0, 0, 1],
                cat_name=category)


def load_cfg_file(path: str) -> dict:
    import yaml
    with open(path) as f:
        raw = yaml.safe_load(f) or {}
    cfg = {k: raw[k] for k in raw}
    cfg["res"] = float(cfg.get("res", 0.002))          # '2e-3' parses as a string in YAML 1.1
    cfg["num_more"] = int(cfg.get("num_more", 3))
    for k in ("up", "right", "front"):
        if k in cfg:
            cfg[k] = [int(v) for v in cfg[k]]
    return cfg


def load_ckpt_cfg(root: str) -> Optional[dict]:
    """`root` = ckpts/<branch>/<cat>-num_more-3; returns None when the frozen config is absent."""
    path = os.path.join(root, ".hydra", "config.yaml")
    return load_cfg_file(path) if os.path.exists(path) else None
