"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/map_eval.pkl from the UNMODIFIED reference scoring code
(/root/reference/utils/util.py: compute_degree_cm_mAP, compute_3d_iou_new, compute_RT_degree_cm_symmetry, imported under
the stubs of oracle/ref_import.py; the figures go to the matplotlib stand-in) on a seeded synthetic result set:

    python oracle/make_map_golden.py

The golden holds the inputs (per-frame result dicts as eval.py writes them), the reference's AP tables for the thresholds
eval.py:400-412 passes, and sample box pairs with the reference's IoU and (degree, cm) errors.
"""
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

SYNSETS = ["BG", "bottle", "bowl", "camera", "can", "laptop", "mug"]


def random_rotation(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def small_rotation(rng, max_deg):
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = np.radians(rng.uniform(0, max_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def make_results(seed=2024, frames=14):
    rng = np.random.default_rng(seed)
    out = []
    for f in range(frames):
        n_gt = int(rng.integers(2, 7))
        cls = rng.integers(1, 7, n_gt)
        gt_RTs, gt_scales = np.zeros((n_gt, 4, 4)), np.zeros((n_gt, 3))
        for i in range(n_gt):
            size = rng.uniform(0.08, 0.35)                       # metric size = ||scale|| factor inside RT
            shape = rng.uniform(0.3, 1.0, 3)
            shape /= np.linalg.norm(shape)
            gt_RTs[i] = np.eye(4)
            gt_RTs[i, :3, :3] = random_rotation(rng) * size
            gt_RTs[i, :3, 3] = rng.uniform([-0.3, -0.3, 0.6], [0.3, 0.3, 1.8])
            gt_scales[i] = shape
        pred_cls, pred_RTs, pred_scales, pred_scores = [], [], [], []
        for i in range(n_gt):
            if rng.random() < 0.12:                              # missed detection
                continue
            rt = gt_RTs[i].copy()
            size = np.cbrt(np.linalg.det(rt[:3, :3]))
            R = rt[:3, :3] / size
            R = small_rotation(rng, 25.0) @ R
            rt[:3, :3] = R * size * rng.uniform(0.85, 1.15)
            rt[:3, 3] += rng.normal(0, rng.choice([0.01, 0.03, 0.08]), 3)
            sc = gt_scales[i] * rng.uniform(0.85, 1.15, 3)
            pred_cls.append(int(cls[i]) if rng.random() > 0.05 else int(rng.integers(1, 7)))
            pred_RTs.append(rt)
            pred_scales.append(sc / np.linalg.norm(sc))
            pred_scores.append(rng.uniform(0.3, 1.0))
        for _ in range(int(rng.integers(0, 2))):                 # false positive
            rt = np.eye(4)
            rt[:3, :3] = random_rotation(rng) * rng.uniform(0.08, 0.35)
            rt[:3, 3] = rng.uniform([-0.3, -0.3, 0.6], [0.3, 0.3, 1.8])
            sc = rng.uniform(0.3, 1.0, 3)
            pred_cls.append(int(rng.integers(1, 7)))
            pred_RTs.append(rt)
            pred_scales.append(sc / np.linalg.norm(sc))
            pred_scores.append(rng.uniform(0.05, 0.6))
        n_pred = len(pred_cls)
        out.append({
            "image_path": f"data/real/test/scene_{f // 7 + 1}/{f % 7:04d}",
            "gt_class_ids": cls.astype(np.int32),
            "gt_RTs": gt_RTs,
            "gt_scales": gt_scales,
            "gt_handle_visibility": (rng.random(n_gt) > 0.4).astype(np.int32),
            "pred_class_ids": np.array(pred_cls, dtype=np.int32),
            "pred_RTs": np.array(pred_RTs).reshape(n_pred, 4, 4),
            "pred_scales": np.array(pred_scales).reshape(n_pred, 3),
            "pred_scores": np.array(pred_scores),
            "pred_bboxes": np.ones((n_pred, 4)),
        })
    return out


def main():
    ref_import.load_reference()
    util = sys.modules["utils.util"]
    util.plt = ref_import._Anything()          # the figures of compute_degree_cm_mAP go nowhere (matplotlib is absent)
    results = make_results()
    import copy
    with tempfile.TemporaryDirectory() as tmp:
        iou_aps, pose_aps = util.compute_degree_cm_mAP(copy.deepcopy(results), SYNSETS, tmp, degree_thresholds=[5, 10, 15],
                                                       shift_thresholds=[5, 10, 15], iou_3d_thresholds=np.linspace(0, 1, 101),
                                                       iou_pose_thres=0.1, use_matches_for_pose=True, num_proc=1)
        iou_aps_d, pose_aps_d = util.compute_degree_cm_mAP(copy.deepcopy(results), SYNSETS, tmp, degree_thresholds=[5, 10],
                                                           shift_thresholds=[2, 5], iou_3d_thresholds=[0.25, 0.5, 0.75],
                                                           iou_pose_thres=0.1, use_matches_for_pose=False, num_proc=1)
    # box pairs: reference IoU and pose error per pair
    rng = np.random.default_rng(7)
    pairs = []
    for k in range(60):
        c = int(rng.integers(1, 7))
        rt1, rt2 = np.eye(4), np.eye(4)
        rt1[:3, :3] = random_rotation(rng) * rng.uniform(0.1, 0.3)
        rt1[:3, 3] = rng.uniform(-0.1, 0.1, 3)
        rt2[:3, :3] = small_rotation(rng, 40.0) @ rt1[:3, :3] * rng.uniform(0.8, 1.2)
        rt2[:3, 3] = rt1[:3, 3] + rng.normal(0, 0.03, 3)
        s1, s2 = rng.uniform(0.3, 1.0, 3), rng.uniform(0.3, 1.0, 3)
        hv = int(rng.integers(0, 2))
        if k < 4:                                                  # far apart: no overlap
            rt2[:3, 3] += 2.0
        if k == 4:                                                 # identical boxes
            rt2, s2 = rt1.copy(), s1.copy()
        iou = util.compute_3d_iou_new(rt1.copy(), rt2.copy(), s1.copy(), s2.copy(), hv, SYNSETS[c], SYNSETS[c])
        err = util.compute_RT_degree_cm_symmetry(rt1.copy(), rt2.copy(), c, hv, SYNSETS)
        pairs.append({"class_id": c, "RT_1": rt1, "RT_2": rt2, "scales_1": s1, "scales_2": s2, "handle_visibility": hv,
                      "iou": float(iou), "err": np.asarray(err, dtype=np.float64)})
    out = {"results": results, "synsets": SYNSETS, "iou_aps": iou_aps, "pose_aps": pose_aps,
           "iou_aps_detection": iou_aps_d, "pose_aps_detection": pose_aps_d, "pairs": pairs}
    path = os.path.join(ROOT, "tests", "golden", "map_eval.pkl")
    with open(path, "wb") as f:
        pickle.dump(out, f, protocol=4)
    print("wrote", path, os.path.getsize(path), "bytes; mean IoU25 / IoU50 =", iou_aps[-1, 25], iou_aps[-1, 50], "; 10 deg 5 cm =", pose_aps[-1, 1, 0])


if __name__ == "__main__":
    main()
