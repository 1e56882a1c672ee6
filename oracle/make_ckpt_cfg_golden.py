"""Mints tests/golden/ckpt_cfgs.json: the frozen per-checkpoint Hydra configs the reference reads at inference
(ckpts/<branch>/<cat>-num_more-3/.hydra/config.yaml, eval.py:91-98) next to the five keys eval.py uses
(res, num_more, up, right, front: eval.py:172,192,210,238-240,298-299).

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference).  The expected values are produced by a
plain YAML load of the reference's own files, independently of cppf2_b200.config (which the test then checks against
them); the YAML texts travel with the fixture because /root/reference does not exist on the GPU box.
"""
import glob
import json
import os

import yaml

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ckpt_cfgs.json")

items = []
for path in sorted(glob.glob(os.path.join(REF, "ckpts", "*", "*", ".hydra", "config.yaml"))):
    text = open(path).read()
    raw = yaml.safe_load(text)
    rel = os.path.relpath(os.path.dirname(os.path.dirname(path)), REF)          # ckpts/<branch>/<cat>-num_more-3
    expected = dict(res=float(raw["res"]), num_more=int(raw["num_more"]), up=[int(v) for v in raw["up"]],
                    right=[int(v) for v in raw["right"]], front=[int(v) for v in raw["front"]])
    items.append(dict(root=rel, yaml=text, expected=expected, cat_name=raw.get("cat_name")))
json.dump(items, open(OUT, "w"), indent=1)
print(f"wrote {OUT}: {len(items)} configs")
