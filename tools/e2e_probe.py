import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from cppf2_b200.estimator import Instance, PoseEstimator, build_models
torch.cuda.set_device(0)
raw = bench.build_frame(0)
cats = sorted({i["category"] for i in raw})
models, cfgs = build_models(cats, precision=1)
est = PoseEstimator(models, cfgs, num_pairs=50000, num_rots=180)
insts = []
for inst in raw:
    it = Instance(pc=torch.from_numpy(inst["pc"]).pin_memory(), category=inst["category"], desc=torch.from_numpy(inst["desc"]).pin_memory(), point_idxs=None)
    it.cells_hint = est.voter.grid_cells_on_host(inst["pc"], inst["cfg"]["res"])
    insts.append(it)
for _ in range(3): est.estimate(insts)
torch.cuda.synchronize()
# H2D bandwidth
big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty_like(big, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); print("H2D GB/s", 5 * 64 / 1024 / (time.perf_counter() - t0))
K = 20
t0 = time.perf_counter()
for _ in range(K): est.estimate(insts)
print("sync estimate ms", (time.perf_counter() - t0) / K * 1e3)
ts, tr = [], []
t0 = time.perf_counter(); pending = None
for _ in range(K):
    a = time.perf_counter(); nxt = est.submit(insts); b = time.perf_counter()
    if pending is not None: pending.result()
    c = time.perf_counter(); ts.append(b - a); tr.append(c - b); pending = nxt
pending.result()
print("pipelined ms", (time.perf_counter() - t0) / K * 1e3, "submit", np.median(ts) * 1e3, "result wait", np.median(tr) * 1e3)
# device-resident pipelined (no H2D)
dinsts = []
for inst, it in zip(raw, insts):
    d = Instance(pc=it.pc.cuda(), category=it.category, desc=it.desc.cuda(), point_idxs=None); d.cells_hint = it.cells_hint; dinsts.append(d)
t0 = time.perf_counter(); pending = None
for _ in range(K):
    nxt = est.submit(dinsts)
    if pending is not None: pending.result()
    pending = nxt
pending.result()
print("pipelined device-resident ms", (time.perf_counter() - t0) / K * 1e3)
