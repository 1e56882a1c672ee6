"""The on-disk formats either side of the hot path in the reference's evaluation script.

Upstream of the path, eval.py reads one detection-result pickle per frame -- `results_*.pkl` written by the mask
network (SAR-Net's Mask-RCNN results, eval.py:74-77, 103-127): a dict (or a list of dicts) with `image_path`,
`pred_bboxes [n,4]`, `pred_masks [H,W,n]`, `pred_class_ids [n]`, `gt_RTs`, `gt_scales`, `gt_class_ids` and, in newer
files, `gt_handle_visibility`.  Downstream it writes the same dict back with `pred_RTs [n,4,4]` (R * ||scale|| and t) and
`pred_scales [n,3]` (scale / ||scale||) filled in (eval.py:143-144, 369-371, 399), which is what `compute_degree_cm_mAP`
(utils/util.py:2736-2955; here `cppf2_b200.scoring.compute_degree_cm_mAP`, offline host code) consumes.

This module is host code only (pickle, numpy, paths); the hot path itself runs through `PoseEstimator.estimate_frame`.
"""
from __future__ import annotations

import os
import pickle
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np

# dataset.py:28-37 (category2id / id2category); 0 is the background class of the detector
ID2CATEGORY = {1: "bottle", 2: "bowl", 3: "camera", 4: "can", 5: "laptop", 6: "mug"}
WHITELIST = ("can", "bowl", "laptop", "bottle", "camera", "mug")          # eval.py:78
REAL275_INTRINSICS = np.array([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]])   # eval.py:82


def load_results(log_dir) -> List[dict]:
    """eval.py:103-127: every `results_*.pkl` under `log_dir`, in sorted order, flattened to one list of per-frame dicts;
    `gt_handle_visibility` defaults to ones when the file predates it."""
    paths = sorted(Path(log_dir).glob("results_*.pkl"))
    if not paths:
        raise FileNotFoundError(f"no results_*.pkl under {log_dir}")          # the reference asserts here (eval.py:108)
    final: List[dict] = []
    for p in paths:
        with open(p, "rb") as f:
            result = pickle.load(f)
        items = result if isinstance(result, list) else [result]
        if not isinstance(result, (list, dict)):
            raise TypeError(f"{p}: expected a dict or a list of dicts, got {type(result).__name__}")
        for r in items:
            if "gt_handle_visibility" not in r:
                r["gt_handle_visibility"] = np.ones_like(r["gt_class_ids"])
            elif len(r["gt_handle_visibility"]) != len(r["gt_class_ids"]):
                raise ValueError(f"{p}: gt_handle_visibility and gt_class_ids differ in length")
            final.append(r)
    return final


def image_stem(res: dict, src: str = "data/real/test", dst: str = "NOCS/real_test") -> str:
    """eval.py:133: the frame's path stem; `<stem>_color.png` / `<stem>_depth.png` are the images."""
    return res["image_path"].replace(src, dst)


def output_path(out_dir, image_path: str) -> str:
    """eval.py:134: one pickle per frame, named after the path components below the first."""
    return os.path.join(str(out_dir), "_".join(image_path.split("/")[1:]) + ".pkl")


def _read_depth(path: str) -> np.ndarray:
    """uint16 millimetres, as cv2.imread(path, -1) returns it (eval.py:139)."""
    try:
        import cv2
        depth = cv2.imread(path, -1)
    except ImportError:
        from PIL import Image
        depth = np.asarray(Image.open(path))
    if depth is None:
        raise FileNotFoundError(path)
    return depth


def fill_frame(res: dict, poses: Sequence, instance_ids: Sequence[int]) -> dict:
    """eval.py:143-144 and 369-371: identity / ones for every detection, then the estimated pose of the instances that
    ran (`poses[k]` belongs to detection `instance_ids[k]`; None = skipped by the reference's guards, left at identity)."""
    n = len(res["pred_bboxes"])
    res["pred_RTs"] = np.stack([np.eye(4) for _ in range(n)]) if n else np.zeros((0, 4, 4))
    res["pred_scales"] = np.stack([np.ones((3,)) for _ in range(n)]) if n else np.zeros((0, 3))
    for pose, i in zip(poses, instance_ids):
        if pose is None:
            continue
        res["pred_RTs"][i] = pose.RT
        res["pred_scales"][i] = pose.scale
    return res


def run_results(estimator, results: Iterable[dict], out_dir=None, intrinsics=REAL275_INTRINSICS,
                desc_fn: Optional[Callable] = None, read_depth: Callable[[str], np.ndarray] = _read_depth,
                stem_of: Callable[[dict], str] = image_stem, id2category: Dict[int, str] = ID2CATEGORY,
                whitelist: Sequence[str] = WHITELIST) -> List[dict]:
    """The frame loop of eval.py:132-399 around the device path: per result dict read the depth image, run
    `estimator.estimate_frame` on the detections whose category is whitelisted and has heads (eval.py:163-166), write
    `pred_RTs` / `pred_scales` and, when `out_dir` is given, dump the dict like eval.py:399.  `desc_fn(res, i, pix)` supplies
    the DINOv2 key-point descriptors of detection i (the backbone is not part of this path); None runs the SHOT branch only."""
    done = []
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
    for k, res in enumerate(results):
        stem = stem_of(res)
        depth = read_depth(stem + "_depth.png")
        masks = np.asarray(res["pred_masks"])
        ids, cats, frame_masks = [], [], []
        for i in range(len(res["pred_bboxes"])):
            cat = id2category.get(int(res["pred_class_ids"][i]))
            if cat is None or cat not in whitelist or cat not in estimator.models:
                continue
            ids.append(i)
            cats.append(cat)
            frame_masks.append(np.ascontiguousarray(masks[:, :, i]).astype(bool))
        poses = []
        if ids:
            fn = None if desc_fn is None else (lambda j, pix, _ids=ids, _res=res: desc_fn(_res, _ids[j], pix))
            poses = estimator.estimate_frame(depth, frame_masks, cats, intrinsics, desc_fn=fn, depth_div=1000.0, frame_seed=k)
        fill_frame(res, poses, ids)
        if out_dir is not None:
            with open(output_path(out_dir, stem), "wb") as f:
                pickle.dump(res, f)
        done.append(res)
    return done


def degree_cm_error(RT_pred: np.ndarray, RT_gt: np.ndarray, symmetric_y: bool) -> tuple:
    """The yardstick of the parity tolerance (utils/util.py:588-663 `compute_RT_degree_cm_symmetry`, the no-handle-flip
    core): rotation error in degrees (about y only for the symmetric categories) and translation error in centimetres of two
    4x4 [sR | t] matrices, scale divided out like the reference does with cbrt(det)."""
    R1 = RT_pred[:3, :3] / np.cbrt(np.linalg.det(RT_pred[:3, :3]))
    R2 = RT_gt[:3, :3] / np.cbrt(np.linalg.det(RT_gt[:3, :3]))
    if symmetric_y:
        y = np.array([0.0, 1.0, 0.0])
        c = (R1 @ y) @ (R2 @ y) / (np.linalg.norm(R1 @ y) * np.linalg.norm(R2 @ y))
    else:
        c = (np.trace(R1 @ R2.T) - 1.0) / 2.0
    theta = float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))
    shift = float(np.linalg.norm(RT_pred[:3, 3] - RT_gt[:3, 3]) * 100.0)
    return theta, shift
