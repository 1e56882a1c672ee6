import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from cppf2_b200.heads import BeyondCPPFSHOT
from cppf2_b200.heads_spec import init_state_dict
from oracle.heads_torch import Ref
import test_gpu_heads as t
pc, idx, shot, normal, desc = t.make_inputs(700, 128, seed=701)
sd = init_state_dict("shot", 321)
sd0 = {k: (np.zeros_like(v) if k.endswith("weight") else v) for k, v in sd.items()}
for name, s in (("zero weights", sd0), ("random weights", sd)):
    m = BeyondCPPFSHOT(dict(num_more=3), precision=1).cuda(); m.load_state_dict(s)
    a = [torch.from_numpy(x).cuda() for x in (pc, idx, shot, normal)]
    cls, scale = m(*a); torch.cuda.synchronize()
    with torch.no_grad():
        w, ws = Ref("shot", s, emulate_bf16=True, device="cuda").forward_shot(*a)
    print(name, "max err", float((cls - w).abs().max()), "mean", float((cls - w).abs().mean()), "scale max", float((scale - ws).abs().max()), "range", float(w.max() - w.min()))
