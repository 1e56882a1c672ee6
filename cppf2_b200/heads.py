"""Drop-in `BeyondCPPF` modules (reference train_shot.py:46-122 and train_dino.py:58-133).

    from cppf2_b200.heads import BeyondCPPFSHOT, BeyondCPPFDINO      # eval.py:17-18 import both as aliases
    model = BeyondCPPFSHOT.load_from_checkpoint(path, cfg=cfg).cuda().eval()          # eval.py:98
    preds_cls, preds_scale = model(points, point_idxs_all, shot_feat, normal)        # eval.py:223-224
    preds_cls, preds_scale = dino_model(points, point_descs, point_idxs_all)         # eval.py:221

Same constructor argument (cfg with num_more), same forward signatures and return shapes
(preds_cls [T,6,32], preds_scale [T,3]), same state_dict key names as the Lightning checkpoints, no
Lightning dependency.  The arithmetic runs in libcppf_b200.so: the whole forward is four kernel launches
with no [T,360] tuple matrix and no per-layer intermediate in HBM.  `precision` selects the float32
CUDA-core path (0, matches torch fp32 to ~1e-5) or the bf16 tensor-core path (1).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check
from .heads_spec import init_state_dict, linear_shapes
from .voting import device_index_tensor, idx_args, stream_ptr, to_device


def _cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class BeyondCPPF:
    """Inference-only head.  Not an nn.Module on purpose: there is no autograd path through the kernels."""

    branch = "shot"

    def __init__(self, cfg=None, precision: int = 0):
        self.cfg = cfg
        self.num_more = int(_cfg_get(cfg, "num_more", 3))
        self.precision = int(precision)
        self._state: Dict[str, np.ndarray] = {}
        self._handle = None
        self._device = None
        self._ws_by_stream: Dict[tuple, torch.Tensor] = {}
        self.load_state_dict(init_state_dict(self.branch, seed=0, num_more=self.num_more))

    # -- nn.Module-like surface the reference scripts touch ------------------------------------------------
    def eval(self):
        return self

    def cuda(self, device=None):
        self._device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        return self

    def to(self, device):
        return self.cuda(torch.device(device).index)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: torch.from_numpy(v.copy()) for k, v in self._state.items()}

    def load_state_dict(self, sd, strict: bool = True):
        new = {}
        for prefix, dout, din in linear_shapes(self.branch, self.num_more):
            for suffix, shape in ((".weight", (dout, din)), (".bias", (dout,))):
                key = prefix + suffix
                if key not in sd:
                    raise KeyError(f"state_dict is missing {key}")
                v = sd[key]
                v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
                if tuple(v.shape) != shape:
                    raise ValueError(f"{key}: expected shape {shape}, got {tuple(v.shape)}")
                new[key] = np.ascontiguousarray(v, dtype=np.float32)
        if strict:
            extra = set(sd) - set(new)
            if extra:
                raise KeyError(f"unexpected keys in state_dict: {sorted(extra)[:4]} ...")
        self._state = new
        self._release()
        return self

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, cfg=None, map_location=None, precision: int = 0, seed: Optional[int] = None):
        """Lightning-style constructor (eval.py:93,98).  Reads ckpt['state_dict']; when the file is absent
        (the reference mount ships no weights) a seeded random initialisation stands in and says so."""
        model = cls(cfg, precision=precision)
        path = str(checkpoint_path)
        if os.path.exists(path):
            ckpt = torch.load(path, map_location="cpu", weights_only=False)
            model.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt, strict=False)
            model.checkpoint = path
        else:
            s = seed if seed is not None else (abs(hash(os.path.basename(os.path.dirname(os.path.dirname(path))))) % (2 ** 31))
            model.load_state_dict(init_state_dict(cls.branch, seed=s, num_more=model.num_more))
            model.checkpoint = None
        return model

    # -- device object ---------------------------------------------------------------------------------------
    def _release(self):
        if self._handle is not None:
            _lib.load().cppf_heads_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure(self, device):
        if self._handle is not None and self._device == device:
            return
        self._release()
        lib = _lib.load()
        flat = np.concatenate([np.concatenate([self._state[p + ".weight"].reshape(-1), self._state[p + ".bias"]])
                               for p, _, _ in linear_shapes(self.branch, self.num_more)]).astype(np.float32)
        handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.cppf_heads_create(0 if self.branch == "shot" else 1, self.num_more,
                                        flat.ctypes.data_as(C.POINTER(C.c_float)), flat.size, C.byref(handle)), "cppf_heads_create")
        self._handle = handle
        self._device = device

    def _run(self, points, idx, feat, normal, sample=None):
        """sample = None: (logits, scale) like the reference's forward.  sample = dict(u01=..., seed=...): the decode of
        eval.py:225-229 runs as the epilogue of the logits head and (bins u8 [T,6], scale) come back instead."""
        lib = _lib.load()
        pc = to_device(points, torch.float32)
        dev = pc.device
        self._ensure(dev)
        idx = device_index_tensor(idx, dev)
        feat = to_device(feat, torch.float32, dev)
        nrm = None if normal is None else to_device(normal, torch.float32, dev)
        T, n = idx.shape[0], pc.shape[0]
        scale = torch.empty((T, 3), dtype=torch.float32, device=dev)
        need = int(lib.cppf_heads_workspace_bytes(self._handle, T, n, self.precision))
        # one scratch per (device, stream): the frame driver runs instances of the same category on several streams
        key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
        ws = self._ws_by_stream.get(key)
        if ws is None or ws.numel() < need:
            ws = self._ws_by_stream[key] = torch.empty(need, dtype=torch.uint8, device=dev)
        ip, i64, istr = idx_args(idx)
        if sample is not None:
            u01 = sample.get("u01")
            u01 = None if u01 is None else to_device(u01, torch.float32, dev)
            bins = torch.empty((T, 6), dtype=torch.uint8, device=dev)
            check(lib.cppf_heads_forward_sampled(self._handle, self.precision, pc.data_ptr(), n, ip, i64, istr, T, feat.data_ptr(),
                                                 None if nrm is None else nrm.data_ptr(), None if u01 is None else u01.data_ptr(),
                                                 int(sample.get("seed", 0)), bins.data_ptr(), scale.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), stream_ptr()), "cppf_heads_forward_sampled")
            return bins, scale
        logits = torch.empty((T, 6, 32), dtype=torch.float32, device=dev)
        check(lib.cppf_heads_forward(self._handle, self.precision, pc.data_ptr(), n, ip, i64, istr, T, feat.data_ptr(),
                                     None if nrm is None else nrm.data_ptr(), logits.data_ptr(), scale.data_ptr(),
                                     ws.data_ptr(), ws.numel(), stream_ptr()), "cppf_heads_forward")
        return logits, scale

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)


class BeyondCPPFSHOT(BeyondCPPF):
    """train_shot.py:46-122.  forward(points [N,3], point_idxs_all [T,5], shot_feat [N,352], normal [N,3])."""
    branch = "shot"

    def forward(self, points, point_idxs_all, shot_feat, normal):
        return self._run(points, point_idxs_all, shot_feat, normal)

    def forward_sampled(self, points, point_idxs_all, shot_feat, normal, u01=None, seed: int = 0):
        """forward + softmax + one multinomial draw per (tuple, coordinate) (eval.py:221-229) in one kernel chain."""
        return self._run(points, point_idxs_all, shot_feat, normal, sample=dict(u01=u01, seed=seed))


class BeyondCPPFDINO(BeyondCPPF):
    """train_dino.py:58-133.  forward(points [N,3], point_descs [N,1024], point_idxs_all [T,5])."""
    branch = "dino"

    def forward(self, points, point_descs, point_idxs_all):
        return self._run(points, point_idxs_all, point_descs, None)

    def forward_sampled(self, points, point_descs, point_idxs_all, u01=None, seed: int = 0):
        return self._run(points, point_idxs_all, point_descs, None, sample=dict(u01=u01, seed=seed))
