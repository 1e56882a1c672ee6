"""Layer inventory of the two BeyondCPPF heads and a seeded stand-in for the missing checkpoints.

Shapes follow train_shot.py:46-73 (SHOT branch) and train_dino.py:58-85 (DINO branch); parameter
names are the ones a Lightning `last.ckpt['state_dict']` of those modules carries
(`shot_encoder.0.fc1.weight` ...), so a real checkpoint drops in unchanged.  The reference mount
ships no weights (.MISSING_LARGE_BLOBS), hence `init_state_dict`: numpy-seeded U(-1/sqrt(fan_in),
1/sqrt(fan_in)) for weight and bias -- the distribution torch.nn.Linear's default init draws from --
handed identically to the reference modules (when minting golden vectors) and to this package.
"""
from __future__ import annotations

from itertools import combinations
from typing import Dict, List, Tuple

import numpy as np

NUM_BINS = 32          # logits per canonical coordinate: 64*3 outputs = 6 coords x 32 (train_shot.py:65)
SHOT_DIM = 352
DINO_DIM = 1024


def res_stack_dims(branch: str, num_more: int = 3) -> Dict[str, List[int]]:
    """Widths of every ResLayer stack, keyed by the attribute name the reference uses."""
    k = num_more + 2
    n_pairs = len(list(combinations(range(k), 2)))
    if branch == "shot":
        tuple_in = n_pairs * 4 + k * 64           # train_shot.py:56
        stacks = {"shot_encoder": [SHOT_DIM] + [128] * 5 + [64]}
    elif branch == "dino":
        tuple_in = n_pairs * 3 + 256              # train_dino.py:65
        stacks = {}
    else:
        raise ValueError(f"unknown branch {branch!r}")
    stacks["tuple_encoder"] = [tuple_in] + [128] * 5 + [256]
    stacks["logit_encoder"] = [256, 256, 256, 64 * 3]
    stacks["scale_encoder"] = [256, 128, 64, 3]
    return stacks


def linear_shapes(branch: str, num_more: int = 3) -> List[Tuple[str, int, int]]:
    """(state_dict prefix, out_features, in_features) of every nn.Linear, in module order."""
    out: List[Tuple[str, int, int]] = []
    for name, dims in res_stack_dims(branch, num_more).items():
        for i in range(len(dims) - 1):
            din, dout = dims[i], dims[i + 1]
            out.append((f"{name}.{i}.fc1", dout, din))
            out.append((f"{name}.{i}.fc2", dout, dout))
            if din != dout:
                out.append((f"{name}.{i}.fc0", dout, din))
    if branch == "dino":
        k = num_more + 2
        out.append(("desc_transform", 256, DINO_DIM))          # train_dino.py:80
        out.append(("desc_pair_transform", 256, 256 * k))      # train_dino.py:81
    return out


def init_state_dict(branch: str, seed: int, num_more: int = 3) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}
    for prefix, dout, din in linear_shapes(branch, num_more):
        bound = 1.0 / np.sqrt(din)
        sd[prefix + ".weight"] = rng.uniform(-bound, bound, (dout, din)).astype(np.float32)
        sd[prefix + ".bias"] = rng.uniform(-bound, bound, (dout,)).astype(np.float32)
    return sd


def macs_per_tuple(branch: str, num_more: int = 3) -> int:
    """Multiply-accumulates of the per-tuple stacks (870 793 for SHOT)."""
    total = 0
    for prefix, dout, din in linear_shapes(branch, num_more):
        if prefix.startswith(("shot_encoder", "desc_transform")):
            continue  # per point, not per tuple
        total += dout * din + dout
    return total
