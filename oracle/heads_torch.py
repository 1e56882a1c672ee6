"""TEST INFRASTRUCTURE ONLY.  Plain PyTorch float32 restatement of the two BeyondCPPF heads
(train_shot.py:19-122, train_dino.py:20-133); also the heads leg of the CPU baseline in bench.py.  Pinned against tests/golden/heads.npz (minted from the reference
modules themselves); used as the CPU/GPU float32 reference for the CUDA heads at sizes the golden
fixture does not cover, and -- with emulate_bf16=True -- as the bf16-rounded, fp32-accumulate
reference of the tensor-core path."""
from itertools import combinations

import torch
import torch.nn.functional as F

from cppf2_b200.heads_spec import res_stack_dims


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


class Ref:
    def __init__(self, branch, sd, num_more=3, emulate_bf16=False, device="cpu"):
        self.branch, self.k = branch, num_more + 2
        self.sd = {k: torch.as_tensor(v).to(device=device, dtype=torch.float32) for k, v in sd.items()}
        self.q = _bf16 if emulate_bf16 else (lambda x: x)
        self.stacks = res_stack_dims(branch, num_more)

    def linear(self, x, name):
        return F.linear(self.q(x), self.q(self.sd[name + ".weight"]), self.sd[name + ".bias"])

    def res_layer(self, x, prefix):
        h = F.relu(self.linear(x, prefix + ".fc1"))
        y = self.linear(h, prefix + ".fc2")
        skip = self.linear(x, prefix + ".fc0") if (prefix + ".fc0.weight") in self.sd else self.q(x)
        return y + skip

    def stack(self, x, name):
        for i in range(len(self.stacks[name]) - 1):
            x = self.res_layer(x, f"{name}.{i}")
        return x

    def coords(self, points, idx):
        return torch.cat([points[idx[:, i]] - points[idx[:, j]] for i, j in combinations(range(self.k), 2)], -1)

    def tuple_inputs_shot(self, points, idx, feat, normal):
        feats = torch.cat([feat[idx[:, i]] for i in range(self.k)], -1)
        dots = torch.cat([torch.max(torch.sum(normal[idx[:, i]] * normal[idx[:, j]], -1, keepdim=True),
                                    torch.sum(-normal[idx[:, i]] * normal[idx[:, j]], -1, keepdim=True))
                          for i, j in combinations(range(self.k), 2)], -1)
        return torch.cat([self.coords(points, idx), dots, feats], -1)

    def forward_shot(self, points, idx, shot_feat, normal):
        x = self.tuple_inputs_shot(points, idx, self.stack(shot_feat, "shot_encoder"), normal)
        feat = self.stack(x, "tuple_encoder")
        return self.stack(feat, "logit_encoder").reshape(feat.shape[0], 6, -1), self.stack(feat, "scale_encoder")

    def forward_dino(self, points, descs, idx, hoist_pair=False):
        """hoist_pair: desc_pair_transform evaluated per point and slot, W_k f(desc_n) (+ bias in block 0), each block
        rounded like the tensor-core path stores it, then gathered and summed per tuple -- the same linear map
        (train_dino.py:95-96) with the rounding points of csrc/heads_tc.cu's kActGatherSum."""
        td = self.linear(descs, "desc_transform")
        if hoist_pair:
            W, b = self.sd["desc_pair_transform.weight"], self.sd["desc_pair_transform.bias"]
            d = td.shape[1]
            blocks = [self.q(F.linear(self.q(td), self.q(W[:, k * d:(k + 1) * d]), b if k == 0 else None)) for k in range(self.k)]
            pair = sum(blocks[k][idx[:, k]] for k in range(self.k))
        else:
            pair = self.linear(torch.cat([td[idx[:, i]] for i in range(self.k)], -1), "desc_pair_transform")
        feat = self.stack(torch.cat([self.coords(points, idx), pair], -1), "tuple_encoder")
        return self.stack(feat, "logit_encoder").reshape(feat.shape[0], 6, -1), self.stack(feat, "scale_encoder")
