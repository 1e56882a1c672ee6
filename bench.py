#!/usr/bin/env python
"""Contract benchmark of the CPPF++ pose-voting hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): one synthetic REAL275-shaped 640x480 depth frame with 6 instances,
SHOT + DINO ensemble (random-init heads, seeded unit-norm stand-ins for the DINOv2 descriptors), T = 50 000
tuples x 180 rotations per (instance, branch).  The frames of a stream cycle through a pool of 7 such frames (13-19 k
points each) and are dealt round-robin to the ranks: step k of rank r is frame (k * N + r) % 7, so every rank and every N
sees the same mix.  A step is one pass of the hot path over one frame: SHOT-352
+ normals, tuple sampling, both heads, multinomial decode, centre vote, back-vote filter, rotation votes,
pose + ensemble selection -- 12 (instance, branch) votes = 600 000 tuples.  Depth back-projection and voxel
down-sampling are "next" rows of the scope table and run once, untimed.

metric / value : tuples voted per second, inputs resident in HBM when the timed region starts.
e2e            : same metric through the public call (PoseEstimator.submit()/result(), one frame in flight ahead) with
                 HOST buffers: per step the clouds and descriptors are copied from pinned host memory, the tuple indices
                 are drawn on the device and the pose records are read back; median of three K-step passes.
roofline       : the heads kernel (tensor-bound): executed flops / CUDA-event time of the frame's heads stage (4 launches:
                 per-point and per-tuple programs of both branches over all instances), events recorded by the call at the
                 stage boundaries; per-stage medians.  clocks: in-process NVML, attached before the warm-up.
--opt          : with the reference's online refinement (eval.py:319-355) in every (instance, branch).
N > 1          : one process per GPU (torchrun), frames sharded across ranks, no data-path collective (weak
                 scaling); time is the max over ranks.
sharded        : second leg, the north star's tuple sharding: ONE (instance, branch) of T = 2^22 tuples split over the N
                 ranks, heads included, five NCCL exchange steps (grid all-reduce first); {T, ms, tuples_per_sec,
                 collective_ms, parity} with bit-exact grid / kept-set flags against the unsharded chain.
--impl reference: the CPU oracle (oracle/, the port of the reference's path: PCL-semantics SHOT in C++, torch-CPU
                 float32 heads, C voting) on the host cores, one instance per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_PAIRS, NUM_ROTS, N_INSTANCES = 50000, 180, 6
FRAME_POOL = 7      # distinct synthetic frames; step k of rank r runs frame (k * world + r) % FRAME_POOL (7 is coprime with 2, 4, 8:
                    # every rank cycles through the whole pool, so the ranks' work is balanced and every N times the same mix)
METRIC, UNIT = "tuples_voted_per_sec", "tuples/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def heads_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the heads stage of one frame = its four launches (per-point and per-tuple
    programs of both branches over all instances), from the committed `ncu --set full` capture of this command
    (profiles/r02_frame_ncu_full.json, tools/ncu_summary.py); None when the summary is absent."""
    path = os.path.join(ROOT, "profiles", "r02_frame_ncu_full.json")
    try:
        for k in json.load(open(path))["kernels"]:
            if "chain_tc_kernel" in k["kernel"]:
                return int(round((k["dram_read_MB"] + k["dram_write_MB"]) * 1e6 * k["launches"]))
    except Exception:
        pass
    return None


def build_frame(frame_id: int):
    """Synthetic frame -> per-instance clouds exactly as eval.py:185-201 prepares them (host, untimed)."""
    from cppf2_b200 import synth
    from cppf2_b200.config import default_category_cfg
    from cppf2_b200.estimator import backproject_host, voxel_downsample_host
    frame = synth.synth_real275_frame(frame_id, N_INSTANCES)
    rng = np.random.default_rng(5000 + frame_id)
    instances = []
    for i, cat in enumerate(frame["cats"]):
        cfg = default_category_cfg(cat)
        pc, _ = backproject_host(frame["depth"] / 1000.0, synth.REAL275_K, frame["masks"][i])
        if pc.shape[0] < 50:
            continue
        pc = pc[voxel_downsample_host(pc, cfg["res"], rng)]
        if pc.shape[0] > 50000:                                              # eval.py:195-198
            pc = pc[rng.integers(0, pc.shape[0], 50000)]
        if ((pc.max(0) - pc.min(0)).max() / cfg["res"]) > 1000:              # eval.py:200
            continue
        desc = synth.unit_descriptors(pc.shape[0], 1024, seed=9000 + 10 * frame_id + i)
        instances.append(dict(pc=np.ascontiguousarray(pc, dtype=np.float32), category=cat, desc=desc, cfg=cfg))
    return instances


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled while a timed region runs.  In-process NVML (nvidia_ml_py) on this rank's
    device, initialised BEFORE the warm-up so that no driver initialisation lands inside a timed region; a
    `nvidia-smi` subprocess per sample is only the fallback (its start-up enumerates every GPU of the box and, on a
    fresh box, was seen to slow the launches of the process being measured)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    _nvml = None          # (module, handle) shared by every sampler of the process

    @classmethod
    def attach(cls, index: int):
        """NVML handle of CUDA device `index` (by UUID, so CUDA_VISIBLE_DEVICES remapping cannot pick a neighbour)."""
        if cls._nvml is not None:
            return cls._nvml
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            cls._nvml = (pynvml, h)
        except Exception:
            cls._nvml = (None, None)
        return cls._nvml

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nv, self.h = self.attach(index)
        self.how = "nvml" if self.nv is not None else "nvidia-smi"

    def _sample_nvml(self):
        nv, h = self.nv, self.h
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bits = [0x8, 0x40, 0x20, 0x4]        # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap
        return [str(sm), str(mx), "0"] + ["Active" if mask & b else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nv is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nv is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock query unavailable"], how=self.how)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = [n for k, n in enumerate(self.NAMES) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                    reasons=reasons, samples=len(self.rows), how=self.how)


def cpu_state_dicts(cats):
    from cppf2_b200.heads_spec import init_state_dict
    return {cat: {"dino": init_state_dict("dino", 1234 + 17 * ci), "shot": init_state_dict("shot", 1234 + 17 * ci + 1)}
            for ci, cat in enumerate(cats)}


def run_cpu_instance(inst, sds, rng, timings=None, opt=False):
    from oracle.pipeline_cpu import instance_pose_cpu
    idx = rng.integers(0, inst["pc"].shape[0], (NUM_PAIRS, 5)).astype(np.int64)
    return instance_pose_cpu(inst["pc"], idx, inst["cfg"], sds[inst["category"]], desc=inst["desc"], timings=timings,
                             sym_y_only=inst["category"] in ("can", "bottle", "bowl"), opt=opt)


def reference_shot_sweep(emit):
    """CPU leg of BASELINE config 3 (tools/shot_sweep.py drives the CUDA leg): the PCL-semantics SHOT restatement on the
    host cores, single-threaded (faithful to shot.cpp:25,82) up to 50k points and with all threads."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from shot_sweep import torus_cloud
    from oracle import cpu as oracle
    for n in (10000, 20000, 50000, 100000, 200000):
        pc = torus_cloud(n)
        line = {"impl": "reference", "config": "SHOT sweep (BASELINE configs[2]), CPU port", "points": n, "cpu_threads": os.cpu_count()}
        if n <= 50000:
            t0 = time.perf_counter()
            oracle.shot_compute(pc, 0.02, 0.02, threads=1)
            line["cpu_1thread_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        oracle.shot_compute(pc, 0.02, 0.02, threads=os.cpu_count() or 1)
        line["cpu_all_threads_s"] = time.perf_counter() - t0
        line["cpu_points_per_sec"] = n / line["cpu_all_threads_s"]
        emit(json.dumps(line))


def reference_arm(args, emit=print):
    """The reference's CPU implementation of the path (oracle port), all host threads, one instance per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if getattr(args, "shot_sweep", False):
        return reference_shot_sweep(emit)
    import torch
    from oracle import cpu as oracle
    torch.set_num_threads(os.cpu_count() or 1)
    oracle.set_num_threads(os.cpu_count() or 1)
    instances = build_frame(0)
    sds = cpu_state_dicts(sorted({i["category"] for i in instances}))
    rng = np.random.default_rng(0)
    per = 2 * NUM_PAIRS
    for w in range(args.warmup):
        run_cpu_instance(instances[w % len(instances)], sds, rng, opt=getattr(args, "opt", False))
    timings = {}
    t0 = time.perf_counter()
    for k in range(args.steps):
        run_cpu_instance(instances[k % len(instances)], sds, rng, timings, opt=getattr(args, "opt", False))
    dt = time.perf_counter() - t0
    value = per * args.steps / dt
    sample = (f"1 instance x 2 branches x {NUM_PAIRS} tuples per step (of the {N_INSTANCES}-instance frame); stage seconds "
              + ", ".join(f"{k} {v:.2f}" for k, v in timings.items()))
    emit(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": workload_config(instances_per_step=1),
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "refinement": bool(getattr(args, "opt", False)), "gpu_launches": 0}))


def workload_config(instances_per_step: int = N_INSTANCES):
    """`instances_per_step` < 6 is the CPU arm's bounded sample: one instance of the same frame per step (both branches)."""
    return {"workload": "synthetic REAL275-shaped 640x480 depth frame, 6 instances, SHOT+DINO ensemble (random-init heads, "
                        "seeded unit-norm DINO descriptors), 50000 tuples x 180 rotations per (instance, branch)",
            "instances_per_step": instances_per_step, "tuples_per_step": 2 * NUM_PAIRS * instances_per_step,
            "num_pairs": NUM_PAIRS, "num_rots": NUM_ROTS, "sphere_bins": 720,
            "frame_pool": FRAME_POOL,
            "l2": "256 MB buffer written between timed steps (L2 flush; in the e2e leg on the upload stream ahead of each step's copies)", "parallelism": "frames sharded across ranks, no collective; every stage of a frame launched once for all its instances (cppf_frame_pose)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=int, default=int(os.environ.get("CPPF_PRECISION", "-1")), help="heads: 0 fp32, 1 bf16 tcgen05, -1 best available")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="store_true", help="run the online pose refinement (eval.py:319-355) in every (instance, branch); "
                    "off by default: the CPU arm and the published tolerance are defined without it")
    ap.add_argument("--shot-sweep", action="store_true", help="with --impl reference: CPU leg of the SHOT sweep (config 3)")
    ap.add_argument("--sharded-log2", default="22", help="tuple counts (log2, comma separated) of the tuple-sharded leg; '' skips it")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: anything a library prints at C level (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        os.write(real_stdout, (line + "\n").encode())

    if args.impl == "reference":
        return reference_arm(args, emit)

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the cppf2_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from cppf2_b200 import _lib
    from cppf2_b200.estimator import Instance, PoseEstimator, build_models
    from cppf2_b200.heads_spec import macs_per_tuple
    _lib.load()
    peaks = load_peaks()

    # frame sharding: the frames of the stream are dealt round-robin to the ranks; step k of rank r is frame k * world + r of a
    # stream that cycles through FRAME_POOL synthetic frames
    pool = [build_frame(f) for f in range(FRAME_POOL)]
    raw = pool[0]

    def frame_of(k):
        return (k * world + rank) % FRAME_POOL
    cats = sorted({i["category"] for fr in pool for i in fr})
    models, cfgs = build_models(cats, precision=0)
    # heads precision per model: bf16 tcgen05 where the library carries the packed tensor-core weights for that
    # branch, else the float32 path (--precision 0 forces float32 everywhere)
    used = set()
    for cat in models:
        for m in models[cat].values():
            m._ensure(dev)
            has_tc = bool(_lib.load().cppf_heads_has_tc(m._handle))
            m.precision = 1 if (has_tc and args.precision != 0) else 0
            used.add((m.branch, m.precision))
    precision = 1 if all(p == 1 for _, p in used) else (0 if all(p == 0 for _, p in used) else 2)
    est = PoseEstimator(models, cfgs, num_pairs=NUM_PAIRS, num_rots=NUM_ROTS, seed=rank, opt=args.opt)
    n_inst_of = [len(fr) for fr in pool]
    n_pts_of = [sum(i["pc"].shape[0] for i in fr) for fr in pool]
    n_inst = n_inst_of[0]

    # ---- device-resident inputs for the kernel-side number --------------------------------------------------
    rng = np.random.default_rng(100 + rank)
    dev_frames = []
    for fr in pool:
        dev_instances = []
        for inst in fr:
            idx = torch.from_numpy(rng.integers(0, inst["pc"].shape[0], (NUM_PAIRS, 5), dtype=np.int32)).to(dev)
            di = Instance(pc=torch.from_numpy(inst["pc"]).to(dev), category=inst["category"], desc=torch.from_numpy(inst["desc"]).to(dev),
                          point_idxs=idx)
            di.cells_hint = est.voter.grid_cells_on_host(inst["pc"], inst["cfg"]["res"])
            dev_instances.append(di)
        dev_frames.append(dev_instances)
    pose_buf = torch.zeros((max(n_inst_of) * 2, est.pose_bytes), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ClockSampler.attach(local)         # NVML initialised here, ahead of the warm-up: nothing driver-side starts inside a timed region
    for k in range(max(args.warmup, 3, FRAME_POOL)):      # every frame of the pool once: all buffers reach their final size
        est.enqueue(dev_frames[k % FRAME_POOL], pose_buf)
    torch.cuda.synchronize()
    launches = est.launches

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    step_events = []
    for k in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        est.enqueue(dev_frames[frame_of(k)], pose_buf)      # cppf_frame_pose: every stage once for all instances of the frame
        b.record()
        step_events.append((a, b))
    barrier()
    clocks = sampler.summary()
    step_ms = [a.elapsed_time(b) for a, b in step_events]
    total_ms = float(sum(step_ms))
    tuples_mine = float(sum(2 * NUM_PAIRS * n_inst_of[frame_of(k)] for k in range(args.steps)))

    # per-stage durations for the roofline: the same steps, same estimator, same single stream, with CUDA events recorded by
    # cppf_frame_pose itself at the stage boundaries of the launching stream (the frame is ~25 batched launches in order:
    # tuple sampling, SHOT, heads, centre vote, back-vote filter, rotation vote, pose)
    stage_ms, serial_ms = {}, float(np.median(step_ms))
    if est.frame_call:
        per_step = []
        for k in range(args.steps):
            flush.zero_()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
            est.stage_events = evs
            est.enqueue(dev_frames[frame_of(k)], pose_buf)
            per_step.append(evs)
        est.stage_events = None
        torch.cuda.synchronize()
        for k, name in enumerate(_lib.FRAME_STAGES):          # mean over the steps (the frames of the pool differ in size)
            stage_ms[name] = float(np.mean([evs[k].elapsed_time(evs[k + 1]) for evs in per_step]))
        serial_ms = float(np.mean([evs[0].elapsed_time(evs[7]) for evs in per_step]))
        stage_ms["heads_shot"], stage_ms["heads_dino"] = stage_ms["heads"], 0.0
        stage_ms["vote_shot"] = stage_ms["center"] + stage_ms["backvote"] + stage_ms["rotation"] + stage_ms["pose"]
        stage_ms["vote_dino"] = 0.0
    else:
        est1 = PoseEstimator(models, cfgs, num_pairs=NUM_PAIRS, num_rots=NUM_ROTS, seed=rank, n_streams=1, opt=args.opt)
        stage_events = {}

        def hook(stage, begin):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            stage_events.setdefault(stage, []).append(ev)
        for k in range(FRAME_POOL):
            est1.enqueue(dev_frames[k], pose_buf)
        torch.cuda.synchronize()
        est1.timing_hook = hook
        serial_events = []
        for k in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            est1.enqueue(dev_frames[frame_of(k)], pose_buf)
            b.record()
            serial_events.append((a, b))
        torch.cuda.synchronize()
        est1.timing_hook = None
        serial_ms = float(np.median([a.elapsed_time(b) for a, b in serial_events]))
        for stage, evs in stage_events.items():
            per = len(evs) // (2 * args.steps)                 # (begin, end) pairs of this stage per step
            iv = [evs[i].elapsed_time(evs[i + 1]) for i in range(0, len(evs) - 1, 2)]
            stage_ms[stage] = float(np.median([sum(iv[k * per:(k + 1) * per]) for k in range(args.steps)])) if per else 0.0
        del est1
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        cnt = torch.tensor([tuples_mine], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        tuples_all = float(cnt.item())                # tuples of all ranks over the K steps
    else:
        tuples_all = tuples_mine
    ms_per_step = total_ms / args.steps
    value = tuples_all / (total_ms * 1e-3)

    # ---- end to end through the public call with host buffers -------------------------------------------------
    host_frames, h2d_of = [], []
    for f, fr in enumerate(pool):
        items, nbytes = [], 0
        for inst in fr:
            pc_p, desc_p = torch.from_numpy(inst["pc"]).pin_memory(), torch.from_numpy(inst["desc"]).pin_memory()
            items.append((pc_p, desc_p, inst))
            nbytes += pc_p.numel() * 4 + desc_p.numel() * 4
        host_frames.append(items)
        h2d_of.append(nbytes)
    h2d = int(round(np.mean([h2d_of[frame_of(k)] for k in range(args.steps)])))

    def e2e_instances(f):
        # the public call's inputs: pinned host clouds + descriptors; tuple indices (eval.py:207) are drawn on the device by
        # the estimator (point_idxs=None), uploads run on its copy stream ahead of the kernels
        insts = []
        for j, (pc_p, desc_p, inst) in enumerate(host_frames[f]):
            it = Instance(pc=pc_p, category=inst["category"], desc=desc_p, point_idxs=None)
            it.cells_hint = dev_frames[f][j].cells_hint
            insts.append(it)
        return insts

    def e2e_submit(k):
        return est.submit(e2e_instances(frame_of(k)))

    for k in range(2 * FRAME_POOL):                           # twice through the pool: the caching allocator has seen every size
        poses = est.estimate(e2e_instances(k % FRAME_POOL))
    import gc
    gc.collect()
    gc.disable()          # no collector pause inside the timed host loop
    barrier()
    # the public call, double-buffered: submit(frame k+1) before result(frame k); every step still uploads its clouds and
    # descriptors from pinned host memory and reads its pose records back inside the timed region.  Three passes of K
    # steps each; the MEDIAN pass is reported and all three are listed.
    e2e_passes = []
    e2e_sampler = ClockSampler(local)
    e2e_sampler.start()
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        pending = None
        for k in range(args.steps):
            with torch.cuda.stream(est.copy_stream):      # L2 flush ahead of this step's uploads, off the compute streams
                flush.zero_()
            nxt = e2e_submit(k)
            if pending is not None:
                poses = pending.result()
            pending = nxt
        poses = pending.result()
        barrier()
        e2e_passes.append((time.perf_counter() - t0) * 1e3)
    e2e_clocks = e2e_sampler.summary()
    gc.enable()
    e2e_ms = float(np.median(e2e_passes))      # median of the passes (all of them are listed in the line), not the best
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = tuples_all / (e2e_ms * 1e-3)                  # tuples of all ranks over the K steps / the pass
    d2h = int(round(np.mean([n_inst_of[frame_of(k)] for k in range(args.steps)]) * 2 * est.pose_bytes))

    # ---- tuple-sharded leg (the north star's multi-GPU design; BASELINE configs[3]) -----------------------------------
    # ONE (instance, branch) whose T tuples are split over the N ranks, heads included (eval.py:219-313): every rank runs
    # heads -> decode -> centre votes -> back-vote -> rotation votes -> loss terms on its T/N tuples, five exchange steps
    # cross NVLink (cppf2_b200/sharded.py), and rank 0 re-runs the whole T unsharded for the parity flags.
    sharded = None
    if args.sharded_log2:
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from vote_sweep import run_sharded
        recs = []
        for lg in [int(x) for x in str(args.sharded_log2).split(",") if x]:
            rec = run_sharded(1 << lg, "halfcyl", dev, rank, world, reps=5, warm=2, with_heads=True, check=world > 1)
            if rank == 0:
                recs.append({"T": rec["tuples"], "cloud": rec["cloud"], "grid_cells": rec["grid_cells"], "ms": rec["ms"],
                             "tuples_per_sec": rec["tuples_per_sec"], "collective_ms": rec["collective_ms"],
                             "collective_share": rec["collective_share"], "collectives": rec["collectives"],
                             "n_collectives": rec["n_collectives"], "heads": rec["heads"], "kept": rec["kept"],
                             "parity": rec["parity"]})
        if rank == 0 and recs:
            sharded = dict(recs[0])
            sharded["what"] = ("one (instance, branch) of T tuples sharded over the ranks, SHOT-branch heads included; "
                               "ms = CUDA events per vote, max over ranks; parity = against the unsharded chain on rank 0 "
                               "(null at 1 GPU, where the two are the same call sequence)")
            if len(recs) > 1:
                sharded["more"] = recs[1:]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    n_pts = float(np.mean([n_pts_of[frame_of(k)] for k in range(args.steps)]))          # per step, over this rank's frames
    n_inst = float(np.mean([n_inst_of[frame_of(k)] for k in range(args.steps)]))
    heads_ms = stage_ms.get("heads_shot", 0.0) + stage_ms.get("heads_dino", 0.0)
    vote_ms = stage_ms.get("vote_shot", 0.0) + stage_ms.get("vote_dino", 0.0)
    shot_ms = stage_ms.get("shot", 0.0)
    # flops EXECUTED: the tensor-core DINO head evaluates desc_pair_transform (256 x 1280 + 256 bias) per point instead of
    # per tuple (linearity, csrc/heads_tc.cu kActGatherSum), so those MACs move from the tuple count to the point count
    pair_macs = 256 * 1280 + 256
    hoisted = precision == 1
    flops = 2.0 * (macs_per_tuple("shot") + macs_per_tuple("dino") - (pair_macs if hoisted else 0)) * NUM_PAIRS * n_inst \
        + 2.0 * (258048 + 262144 + (pair_macs if hoisted else 0)) * n_pts
    vote_bytes = 2 * n_inst * (NUM_PAIRS * 24) + 2 * n_pts * 12
    shot_bytes = n_pts * 1432
    kernels = {
        "heads": {"ms": heads_ms, "share": heads_ms / serial_ms, "achieved_tflops": flops / (heads_ms * 1e-3) / 1e12 if heads_ms else None},
        "vote_chain": {"ms": vote_ms, "share": vote_ms / serial_ms, "alg_GBps": vote_bytes / (vote_ms * 1e-3) / 1e9 if vote_ms else None,
                       "stages_ms": {k: stage_ms[k] for k in ("center", "backvote", "rotation", "pose") if k in stage_ms}},
        "shot": {"ms": shot_ms, "share": shot_ms / serial_ms, "alg_GBps": shot_bytes / (shot_ms * 1e-3) / 1e9 if shot_ms else None},
    }
    if heads_ms >= max(vote_ms, shot_ms):
        peak = peaks["bf16_sustained"]
        ach = kernels["heads"]["achieved_tflops"]
        roofline = {"kernel": "heads (ResLayer chains, %s)" % {0: "fp32 CUDA cores", 1: "bf16 tcgen05", 2: "SHOT bf16 tcgen05 + DINO fp32 CUDA cores"}[precision], "bound": "tensor",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": heads_traffic(),
                    "traffic_note": "DRAM bytes of the heads stage of one frame (4 launches; ncu --set full, profiles/r02_frame_ncu_full.md); "
                                    "algorithmic: descriptors 88 MB + tuple indices 12 MB + draws/scales 11 MB + weights 8 MB per frame",
                    "peak_source": peaks["source"] + ", sustained"}
    elif vote_ms >= shot_ms:
        ach = kernels["vote_chain"]["alg_GBps"]
        roofline = {"kernel": "vote chain (decode, centre vote, back-vote, rotation, pose)", "bound": "hbm", "achieved": ach,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"]}
    else:
        ach = kernels["shot"]["alg_GBps"]
        roofline = {"kernel": "SHOT-352 (grid build, normals, LRF + histogram)", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"]}

    # ---- CPU baseline: the oracle port on this box's host cores, one full frame ----------------------------------
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only (bench contract)
        from oracle import cpu as oracle
        torch.set_num_threads(os.cpu_count() or 1)
        oracle.set_num_threads(os.cpu_count() or 1)
        sds = cpu_state_dicts(cats)
        crng = np.random.default_rng(0)
        if args.opt:      # torch.optim / autograd initialise lazily on first use (seconds): not part of the path's cost
            from oracle.refine_torch import refine_pose
            refine_pose(raw[0]["pc"][:64], np.zeros((8, 2), np.int64), np.zeros((8, 2, 3), np.float32), np.zeros(3), np.eye(3), False, iters=2)
        timings = {}
        t0 = time.perf_counter()
        n_done = 0
        for inst in raw:                                   # frame 0 of the pool
            run_cpu_instance(inst, sds, crng, timings, opt=args.opt)
            n_done += 1
            if time.perf_counter() - t0 > 30.0:
                break
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": 2 * NUM_PAIRS * n_done / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                        "sample": f"{n_done} of {len(raw)} instances of frame 0 of the pool, both branches, {dt:.1f} s; stage seconds "
                                  + ", ".join(f"{k} {v:.2f}" for k, v in timings.items())}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "bf16", 2: "bf16 (SHOT head, tcgen05) + f32 (DINO head)"}[precision], "data": "synthetic", "config": workload_config(),
            "frames_per_sec": world * 1e3 / ms_per_step, "instances": n_inst, "points": n_pts, "refinement": bool(args.opt),
            "frame_pool": {"frames": FRAME_POOL, "points_per_frame": n_pts_of, "instances_per_frame": n_inst_of,
                           "schedule": "step k of rank r runs frame (k * n_gpus + r) % frames"},
            "roofline": roofline, "kernels": kernels, "sharded": sharded,
            "kernel_timing": {"how": ("the timed configuration itself (cppf_frame_pose: every stage launched once for all instances, one stream), "
                                      "CUDA events recorded by the call at the stage boundaries, median over the steps") if est.frame_call else
                                     "same steps on one stream (instances serialised), CUDA events per stage on the launching stream",
                              "ms_per_step_staged": serial_ms, "frame_call": bool(est.frame_call), "cuda_graph": bool(est.use_graph)},
            "cpu_baseline": cpu_baseline, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "frames_per_sec": world * 1e3 / (e2e_ms / args.steps),
                    "passes_ms_per_step": [t / args.steps for t in e2e_passes], "clocks": e2e_clocks,
                    "api": "PoseEstimator.submit()/result(), one frame in flight ahead; pinned host clouds + descriptors in, pose records out; "
                           "median of three passes of K steps (all listed in passes_ms_per_step)"},
            "gpu_launches": launches * args.steps,
            "pose_check": {"finite": bool(all(p is not None and np.isfinite(p.RT).all() for p in poses)),
                           "branches": [p.branch for p in poses if p is not None]}}
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
