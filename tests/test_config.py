"""Config / checkpoint surface: cppf2_b200.config reads the reference's frozen per-checkpoint Hydra
files with PyYAML and must reproduce the five keys eval.py uses for every shipped checkpoint directory.  The fixture
(tests/golden/ckpt_cfgs.json, minted by oracle/make_ckpt_cfg_golden.py from /root/reference/ckpts) carries the 12 YAML texts
and the expected values."""
import json
import os

import pytest

from cppf2_b200 import config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_cfgs.json")
ITEMS = json.load(open(GOLDEN))


def test_fixture_covers_all_shipped_checkpoints():
    roots = {it["root"] for it in ITEMS}
    assert len(ITEMS) == 12
    for br in ("dino", "shot"):
        for cat in ("bottle", "bowl", "camera", "can", "laptop", "mug"):
            assert f"ckpts/{br}/{cat}-num_more-3" in roots


@pytest.mark.parametrize("item", ITEMS, ids=[it["root"] for it in ITEMS])
def test_load_ckpt_cfg_reproduces_inference_keys(item, tmp_path):
    root = tmp_path / item["root"]
    (root / ".hydra").mkdir(parents=True)
    (root / ".hydra" / "config.yaml").write_text(item["yaml"])
    cfg = config.load_ckpt_cfg(str(root))
    assert cfg is not None
    for key in config.INFERENCE_KEYS:
        assert cfg[key] == item["expected"][key], key
        assert type(cfg[key]) is type(item["expected"][key])
    # the shipped configs: laptop is the only category voted at 1 cm, and the defaults used when a
    # checkpoint directory is absent agree with every shipped file
    assert cfg["res"] == (0.01 if item["cat_name"] == "laptop" else 0.002)
    dflt = config.default_category_cfg(item["cat_name"])
    assert all(dflt[k] == cfg[k] for k in config.INFERENCE_KEYS)


def test_missing_ckpt_dir_returns_none(tmp_path):
    assert config.load_ckpt_cfg(str(tmp_path / "ckpts" / "shot" / "nothing")) is None


def test_build_models_reads_ckpt_cfg(tmp_path):
    """eval.py:87-101: per-category cfg comes from the checkpoint directory, not from config/category/*.yaml."""
    item = next(it for it in ITEMS if it["root"] == "ckpts/shot/laptop-num_more-3")
    root = tmp_path / item["root"]
    (root / ".hydra").mkdir(parents=True)
    (root / ".hydra" / "config.yaml").write_text(item["yaml"])
    assert config.load_ckpt_cfg(str(root))["res"] == 0.01
