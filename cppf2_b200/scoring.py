"""NOCS-style scoring of the pose records the hot path writes: 3D-IoU AP and
rotation / translation AP over a list of per-frame result dicts (`results.py`), as eval.py:400-412 calls
`compute_degree_cm_mAP` (reference utils/util.py:2610-2955) after the instance loop.

Offline host code (numpy + scipy's Qhull binding), no device work, no plotting: the reference's figures are left out, its
two pickles (`IoU_3D_AP_*.pkl`, `Pose_*AP_*.pkl`) are written when `log_dir` is given.  Conventions kept from the
reference, each cited where it is implemented:

* poses are `RT = [s R | t]`; rotation and scale are separated by `cbrt(det)` (util.py:2619-2621, 2632-2635);
* the box IoU is the exact IoU of two ORIENTED boxes (util.py:505-514 -> utils/iou.py), maximised over 36 rotations about
  y for the symmetric categories (util.py:519-543);
* the rotation error of symmetric categories is the angle between the y axes (util.py:640-649);
* matching is greedy in score order, one ground truth per prediction (util.py:1727-1752, 1897-1926);
* AP is the VOC area under the monotone precision envelope (util.py:1757-1782).
"""
from __future__ import annotations

import math
import os
import pickle
from typing import List, Optional, Sequence, Tuple

import numpy as np

SYMMETRIC_ALWAYS = ("bottle", "bowl", "can")          # util.py:519, 640
SYMMETRIC_WITHOUT_HANDLE = ("mug",)                   # util.py:519, 646 (the reference's longer list has no other NOCS class)
NOCS_SYNSETS = ["BG", "bottle", "bowl", "camera", "can", "laptop", "mug"]      # eval.py:400-406
_PLANE_EPS = 1e-6                                     # thickness of a clipping plane, metres (utils/iou.py:9)


# ---------------------------------------------------------------------------------------------------------------------
# oriented boxes
# ---------------------------------------------------------------------------------------------------------------------
def _unit_rotation(m: np.ndarray) -> np.ndarray:
    """util.py:507-508: the 3x3 block divided by the cube root of its determinant."""
    return m / np.cbrt(np.linalg.det(m))


def _corners(R: np.ndarray, t: np.ndarray, s: np.ndarray) -> np.ndarray:
    """[8,3]: corner (i,j,k) = t + R (+-s/2); index = 4 i + 2 j + k with 0 -> -, 1 -> +."""
    signs = np.array([[i, j, k] for i in (-0.5, 0.5) for j in (-0.5, 0.5) for k in (-0.5, 0.5)])
    return (signs * s) @ R.T + t


# the six faces as corner indices (any cyclic order around the face)
_FACES = np.array([[0, 1, 3, 2], [4, 5, 7, 6], [0, 1, 5, 4], [2, 3, 7, 6], [0, 2, 6, 4], [1, 3, 7, 5]])


def _half_spaces(R: np.ndarray, t: np.ndarray, s: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """The box as {x : n_i . x <= d_i}, i < 6: n = +-column k of R, d = +-(column . t) + s_k / 2."""
    n = np.concatenate([R.T, -R.T], 0)
    c = R.T @ t
    d = np.concatenate([c + 0.5 * s, -c + 0.5 * s])
    return n, d


def _clip(poly: List[np.ndarray], n: np.ndarray, d: float) -> List[np.ndarray]:
    """Sutherland-Hodgman: the part of a convex polygon inside n . x <= d.  Points within the plane's thickness count as on
    the plane and are kept once."""
    if len(poly) <= 1:
        return []
    dist = [float(n @ p) - d for p in poly]
    side = [0 if abs(v) <= _PLANE_EPS else (1 if v > 0 else -1) for v in dist]       # +1 outside, -1 inside
    if all(v == 0 for v in side):
        return poly
    out: List[np.ndarray] = []
    m = len(poly)
    for i in range(m):
        p, q = poly[i - 1], poly[i]
        sp, sq = side[i - 1], side[i]
        if sq < 0:                                   # current vertex inside
            if sp > 0:                               # entering: the crossing point first
                a = dist[i - 1] / (dist[i - 1] - dist[i])
                out.append(p + a * (q - p))
            elif sp == 0 and not (out and np.array_equal(out[-1], p)):
                out.append(p)
            out.append(q)
        elif sq > 0:                                 # current vertex outside
            if sp < 0:                               # leaving
                a = dist[i - 1] / (dist[i - 1] - dist[i])
                out.append(p + a * (q - p))
            elif sp == 0 and not (out and np.array_equal(out[-1], p)):
                out.append(p)
        else:                                        # current vertex on the plane
            if sp != 0:
                out.append(q)
    return out


def oriented_box_iou(R1, t1, s1, R2, t2, s2) -> float:
    """Exact IoU of two oriented boxes {t + R u : |u_k| <= s_k / 2} (utils/iou.py:25-40): the vertices of the intersection
    polytope are collected by clipping every face of one box against the half-spaces of the other, both ways, and its
    volume is the volume of their convex hull.  Returns 0 when the boxes do not overlap in a solid."""
    from scipy.spatial import ConvexHull, QhullError

    R1, R2 = np.asarray(R1, np.float64), np.asarray(R2, np.float64)
    t1, t2 = np.asarray(t1, np.float64).reshape(3), np.asarray(t2, np.float64).reshape(3)
    s1, s2 = np.asarray(s1, np.float64).reshape(3), np.asarray(s2, np.float64).reshape(3)
    # bounding spheres apart: no solid overlap (most prediction / ground-truth pairs of a frame end here)
    r1 = 0.5 * math.sqrt(float(((np.linalg.norm(R1, axis=0) * s1) ** 2).sum()))
    r2 = 0.5 * math.sqrt(float(((np.linalg.norm(R2, axis=0) * s2) ** 2).sum()))
    if float(np.linalg.norm(t1 - t2)) > r1 + r2:
        return 0.0
    pts: List[np.ndarray] = []
    for (Ra, ta, sa), (Rb, tb, sb) in (((R1, t1, s1), (R2, t2, s2)), ((R2, t2, s2), (R1, t1, s1))):
        n, d = _half_spaces(Ra, ta, sa)
        corners = _corners(Rb, tb, sb)
        for face in _FACES:
            poly = [corners[i] for i in face]
            for k in range(6):
                poly = _clip(poly, n[k], float(d[k]))
                if not poly:
                    break
            pts.extend(poly)
    if len(pts) < 4:
        return 0.0
    try:
        inter = float(ConvexHull(np.asarray(pts)).volume)
    except (QhullError, ValueError):
        return 0.0                                   # flat or degenerate contact (the reference returns 0 on any exception, util.py:513)
    v1 = abs(float(np.linalg.det(R1))) * float(np.prod(s1))
    v2 = abs(float(np.linalg.det(R2))) * float(np.prod(s2))
    return inter / (v1 + v2 - inter)


def _y_rotation(theta: float) -> np.ndarray:
    c, s = math.cos(theta), math.sin(theta)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def box_iou_3d(RT_1, RT_2, scales_1, scales_2, handle_visibility, class_name_1: str, class_name_2: str) -> float:
    """compute_3d_iou_new (util.py:475-548).  Symmetric categories (and a mug whose handle is hidden) take the maximum over
    36 rotations of the first box about its y axis."""
    if RT_1 is None or RT_2 is None:
        return -1
    RT_1, RT_2 = np.asarray(RT_1, np.float64), np.asarray(RT_2, np.float64)

    def one(rot1: np.ndarray) -> float:
        try:
            return oriented_box_iou(_unit_rotation(rot1), RT_1[:3, 3], scales_1, _unit_rotation(RT_2[:3, :3]), RT_2[:3, 3], scales_2)
        except Exception:                            # util.py:513: any failure counts as no overlap
            return 0.0

    same = class_name_1 == class_name_2
    if same and (class_name_1 in SYMMETRIC_ALWAYS or (class_name_1 in SYMMETRIC_WITHOUT_HANDLE and handle_visibility == 0)):
        n = 36
        return max([0.0] + [one(RT_1[:3, :3] @ _y_rotation(2.0 * math.pi * i / float(n))) for i in range(n)])
    return one(RT_1[:3, :3])


def rotation_translation_error(RT_1, RT_2, class_id: int, handle_visibility, synset_names: Sequence[str]):
    """compute_RT_degree_cm_symmetry (util.py:588-663): [rotation error in degrees, translation error in centimetres]."""
    if RT_1 is None or RT_2 is None:
        return -1
    RT_1, RT_2 = np.asarray(RT_1, np.float64), np.asarray(RT_2, np.float64)
    if not (np.array_equal(RT_1[3], RT_2[3]) and np.array_equal(RT_1[3], np.array([0, 0, 0, 1]))):
        raise ValueError(f"last rows must be [0, 0, 0, 1]: {RT_1[3]} {RT_2[3]}")       # the reference prints and exits
    R1, R2 = _unit_rotation(RT_1[:3, :3]), _unit_rotation(RT_2[:3, :3])
    name = synset_names[class_id]
    if name in SYMMETRIC_ALWAYS or (name in SYMMETRIC_WITHOUT_HANDLE and handle_visibility == 0):
        y = np.array([0, 1, 0])
        y1, y2 = R1 @ y, R2 @ y
        theta = np.arccos(y1.dot(y2) / (np.linalg.norm(y1) * np.linalg.norm(y2)))
    else:
        R = R1 @ R2.transpose()
        theta = np.arccos((np.trace(R) - 1) / 2)
    theta *= 180 / np.pi
    shift = np.linalg.norm(RT_1[:3, 3] - RT_2[:3, 3]) * 100
    return np.array([theta, shift])


# ---------------------------------------------------------------------------------------------------------------------
# matching and AP
# ---------------------------------------------------------------------------------------------------------------------
def match_by_iou(gt_class_ids, gt_RTs, gt_scales, gt_handle_visibility, synset_names, pred_class_ids, pred_scores, pred_RTs,
                 pred_scales, iou_thresholds, score_threshold: float = 0):
    """compute_3d_matches (util.py:1665-1754).  Predictions are visited in descending score order (the returned `order`);
    each takes the unmatched ground truth of its class with the largest IoU above the threshold.  Returns
    (gt_matches [n_thr, n_gt], pred_matches [n_thr, n_pred] in score order, overlaps [n_pred, n_gt] float32, order)."""
    num_pred, num_gt = len(pred_class_ids), len(gt_class_ids)
    order = np.zeros(0)
    if num_pred:
        order = np.argsort(pred_scores)[::-1]
        pred_class_ids, pred_scores = pred_class_ids[order], pred_scores[order]
        pred_scales, pred_RTs = pred_scales[order], pred_RTs[order]
    overlaps = np.zeros((num_pred, num_gt), dtype=np.float32)
    for i in range(num_pred):
        for j in range(num_gt):
            overlaps[i, j] = box_iou_3d(pred_RTs[i], gt_RTs[j], pred_scales[i], gt_scales[j], gt_handle_visibility[j],
                                        synset_names[pred_class_ids[i]], synset_names[gt_class_ids[j]])
    pred_matches = -1 * np.ones([len(iou_thresholds), num_pred])
    gt_matches = -1 * np.ones([len(iou_thresholds), num_gt])
    for s, thr in enumerate(iou_thresholds):
        for i in range(num_pred):
            by_iou = np.argsort(overlaps[i])[::-1]
            low = np.where(overlaps[i, by_iou] < score_threshold)[0]
            if low.size > 0:
                by_iou = by_iou[:low[0]]
            for j in by_iou:
                if gt_matches[s, j] > -1:
                    continue
                iou = overlaps[i, j]
                if iou < thr:
                    break
                if not pred_class_ids[i] == gt_class_ids[j]:
                    continue
                if iou > thr:
                    gt_matches[s, j] = i
                    pred_matches[s, i] = j
                    break
    return gt_matches, pred_matches, overlaps, order


def pose_errors(gt_class_ids, gt_RTs, gt_handle_visibility, pred_class_ids, pred_RTs, synset_names) -> np.ndarray:
    """compute_RT_overlaps (util.py:1785-1808): [n_pred, n_gt, 2] = (degrees, centimetres)."""
    out = np.zeros((len(pred_class_ids), len(gt_class_ids), 2))
    for i in range(out.shape[0]):
        for j in range(out.shape[1]):
            out[i, j, :] = rotation_translation_error(pred_RTs[i], gt_RTs[j], gt_class_ids[j], gt_handle_visibility[j], synset_names)
    return out


def match_by_pose(errors: np.ndarray, pred_class_ids, gt_class_ids, degree_thresholds, shift_thresholds):
    """compute_match_from_degree_cm (util.py:1883-1928): per (degree, shift) threshold pair, every prediction takes the
    unmatched ground truth of its class with the smallest degree + centimetre sum that is within both thresholds."""
    num_pred, num_gt = len(pred_class_ids), len(gt_class_ids)
    pred_matches = -1 * np.ones((len(degree_thresholds), len(shift_thresholds), num_pred))
    gt_matches = -1 * np.ones((len(degree_thresholds), len(shift_thresholds), num_gt))
    if num_pred == 0 or num_gt == 0:
        return gt_matches, pred_matches
    assert errors.shape == (num_pred, num_gt, 2)
    nearest = [np.argsort(np.sum(errors[i], axis=-1)) for i in range(num_pred)]
    for d, deg in enumerate(degree_thresholds):
        for s, cm in enumerate(shift_thresholds):
            for i in range(num_pred):
                for j in nearest[i]:
                    if gt_matches[d, s, j] > -1 or pred_class_ids[i] != gt_class_ids[j]:
                        continue
                    if errors[i, j, 0] > deg or errors[i, j, 1] > cm:
                        continue
                    gt_matches[d, s, j] = i
                    pred_matches[d, s, i] = j
                    break
    return gt_matches, pred_matches


def average_precision(pred_match: np.ndarray, pred_scores: np.ndarray, gt_match: np.ndarray) -> float:
    """compute_ap_from_matches_scores (util.py:1757-1782)."""
    assert pred_match.shape[0] == pred_scores.shape[0]
    order = np.argsort(pred_scores)[::-1]
    hit = pred_match[order] > -1
    precisions = np.cumsum(hit) / (np.arange(len(hit)) + 1)
    recalls = np.cumsum(hit).astype(np.float32) / len(gt_match)
    precisions = np.concatenate([[0], precisions, [0]])
    recalls = np.concatenate([[0], recalls, [1]])
    for i in range(len(precisions) - 2, -1, -1):
        precisions[i] = np.maximum(precisions[i], precisions[i + 1])
    steps = np.where(recalls[:-1] != recalls[1:])[0] + 1
    return float(np.sum((recalls[steps] - recalls[steps - 1]) * precisions[steps]))


# ---------------------------------------------------------------------------------------------------------------------
# one frame, then the whole set
# ---------------------------------------------------------------------------------------------------------------------
def _split_scale(RTs: np.ndarray, scales: np.ndarray):
    """util.py:2619-2621: rotation blocks divided by cbrt(det) + 1e-7, the factor moved into the scales."""
    norm = np.stack([np.cbrt(np.linalg.det(rt[:3, :3])) for rt in RTs])
    RTs[:, :3, :3] = RTs[:, :3, :3] / (norm[:, None, None] + 1e-7)
    return RTs, scales * norm[:, None]


def score_frame(res: dict, synset_names: Sequence[str], iou_thresholds, degree_thresholds, shift_thresholds,
                use_matches_for_pose: bool, iou_pose_thres: float):
    """`work` (util.py:2610-2733) for one per-frame result dict: per class the match / score arrays of the frame."""
    num_classes = len(synset_names)
    n_iou, n_deg, n_cm = len(iou_thresholds), len(degree_thresholds), len(shift_thresholds)
    gt_class_ids = np.array(res["gt_class_ids"]).astype(np.int32)
    gt_RTs = np.array(res["gt_RTs"], dtype=np.float64)
    gt_scales = np.array(res["gt_scales"], dtype=np.float64)
    gt_handle = np.array(res["gt_handle_visibility"])
    if len(gt_RTs):
        gt_RTs, gt_scales = _split_scale(gt_RTs, gt_scales)
    pred_class_ids = np.asarray(res["pred_class_ids"])
    pred_scales = np.asarray(res["pred_scales"], dtype=np.float64)
    pred_scores = np.asarray(res["pred_scores"])
    pred_RTs = np.array(res["pred_RTs"], dtype=np.float64)
    if len(pred_RTs) > 0:
        pred_RTs, pred_scales = _split_scale(pred_RTs, pred_scales)

    iou_pred = [np.zeros((n_iou, 0)) for _ in range(num_classes)]
    iou_score = [np.zeros((n_iou, 0)) for _ in range(num_classes)]
    iou_gt = [np.zeros((n_iou, 0)) for _ in range(num_classes)]
    pose_pred = [np.zeros((n_deg, n_cm, 0)) for _ in range(num_classes)]
    pose_score = [np.zeros((n_deg, n_cm, 0)) for _ in range(num_classes)]
    pose_gt = [np.zeros((n_deg, n_cm, 0)) for _ in range(num_classes)]
    if len(gt_class_ids) == 0 and len(pred_class_ids) == 0:
        return iou_pred, iou_score, iou_gt, pose_pred, pose_score, pose_gt

    for cls_id in range(1, num_classes):
        g = gt_class_ids == cls_id
        c_gt_ids = gt_class_ids[g] if len(gt_class_ids) else np.zeros(0)
        c_gt_scales = gt_scales[g] if len(gt_class_ids) else np.zeros((0, 3))
        c_gt_RTs = gt_RTs[g] if len(gt_class_ids) else np.zeros((0, 4, 4))
        p = pred_class_ids == cls_id
        c_pred_ids = pred_class_ids[p] if len(pred_class_ids) else np.zeros(0)
        c_pred_scores = pred_scores[p] if len(pred_class_ids) else np.zeros(0)
        c_pred_RTs = pred_RTs[p] if len(pred_class_ids) else np.zeros((0, 4, 4))
        c_pred_scales = pred_scales[p] if len(pred_class_ids) else np.zeros((0, 3))
        if synset_names[cls_id] != "mug":              # util.py:2669-2672: only the mug's handle can hide
            c_gt_handle = np.ones_like(c_gt_ids)
        else:
            c_gt_handle = gt_handle[g] if len(gt_class_ids) else np.ones(0)

        gt_m, pred_m, _, order = match_by_iou(c_gt_ids, c_gt_RTs, c_gt_scales, c_gt_handle, synset_names, c_pred_ids, c_pred_scores,
                                              c_pred_RTs, c_pred_scales, iou_thresholds)
        if len(order):
            c_pred_ids, c_pred_RTs, c_pred_scores = c_pred_ids[order], c_pred_RTs[order], c_pred_scores[order]
        iou_pred[cls_id] = np.concatenate((iou_pred[cls_id], pred_m), axis=-1)
        iou_score[cls_id] = np.concatenate((iou_score[cls_id], np.tile(c_pred_scores, (n_iou, 1))), axis=-1)
        iou_gt[cls_id] = np.concatenate((iou_gt[cls_id], gt_m), axis=-1)

        if use_matches_for_pose:                       # util.py:2692-2711: only the detections matched at iou_pose_thres are posed
            k = list(iou_thresholds).index(iou_pose_thres)
            keep_p, keep_g = pred_m[k, :] > -1, gt_m[k, :] > -1
            c_pred_ids = c_pred_ids[keep_p] if len(keep_p) > 0 else np.zeros(0)
            c_pred_RTs = c_pred_RTs[keep_p] if len(keep_p) > 0 else np.zeros((0, 4, 4))
            c_pred_scores = c_pred_scores[keep_p] if len(keep_p) > 0 else np.zeros(0)
            c_gt_ids = c_gt_ids[keep_g] if len(keep_g) > 0 else np.zeros(0)
            c_gt_RTs = c_gt_RTs[keep_g] if len(keep_g) > 0 else np.zeros((0, 4, 4))
            c_gt_handle = c_gt_handle[keep_g] if len(keep_g) > 0 else np.zeros(0)

        errs = pose_errors(c_gt_ids, c_gt_RTs, c_gt_handle, c_pred_ids, c_pred_RTs, synset_names)
        gt_pm, pred_pm = match_by_pose(errs, c_pred_ids, c_gt_ids, degree_thresholds, shift_thresholds)
        pose_pred[cls_id] = np.concatenate((pose_pred[cls_id], pred_pm), axis=-1)
        pose_score[cls_id] = np.concatenate((pose_score[cls_id], np.tile(c_pred_scores, (n_deg, n_cm, 1))), axis=-1)
        pose_gt[cls_id] = np.concatenate((pose_gt[cls_id], gt_pm), axis=-1)
    return iou_pred, iou_score, iou_gt, pose_pred, pose_score, pose_gt


def compute_degree_cm_mAP(final_results: List[dict], synset_names: Sequence[str] = NOCS_SYNSETS, log_dir: Optional[str] = None,
                          degree_thresholds=(360,), shift_thresholds=(100,), iou_3d_thresholds=(0.1,), iou_pose_thres: float = 0.1,
                          use_matches_for_pose: bool = False, num_proc: int = 1):
    """The reference's entry point (util.py:2736-2955) without the figures: returns (iou_3d_aps [classes + 1, n_iou],
    pose_aps [classes + 1, n_deg + 1, n_cm + 1]); row -1 is the mean over the object classes, the last degree / shift
    threshold is the catch-all 360 degrees / 100 cm the reference appends.  `num_proc` > 1 scores the frames in a process
    pool (frame order is kept, so ties in the scores resolve the same way on every run)."""
    num_classes = len(synset_names)
    deg_list = list(degree_thresholds) + [360]
    cm_list = list(shift_thresholds) + [100]
    iou_list = list(iou_3d_thresholds)
    if use_matches_for_pose and iou_pose_thres not in iou_list:
        raise ValueError("iou_pose_thres must be one of iou_3d_thresholds")

    args = [(r, synset_names, iou_list, deg_list, cm_list, use_matches_for_pose, iou_pose_thres) for r in final_results]
    if num_proc > 1 and len(args) > 1:
        from multiprocessing import get_context
        with get_context("fork").Pool(num_proc) as pool:
            per_frame = pool.starmap(score_frame, args, chunksize=max(1, len(args) // (4 * num_proc)))
    else:
        per_frame = [score_frame(*a) for a in args]

    def gather(slot: int, cls_id: int, empty_shape):
        parts = [f[slot][cls_id] for f in per_frame]
        return np.concatenate(parts, -1) if parts else np.zeros(empty_shape)

    iou_aps = np.zeros((num_classes + 1, len(iou_list)))
    pose_aps = np.zeros((num_classes + 1, len(deg_list), len(cm_list)))
    for cls_id in range(1, num_classes):
        pm, sc, gm = (gather(k, cls_id, (len(iou_list), 0)) for k in range(3))
        for s in range(len(iou_list)):
            iou_aps[cls_id, s] = average_precision(pm[s, :], sc[s, :], gm[s, :])
        pm, sc, gm = (gather(k, cls_id, (len(deg_list), len(cm_list), 0)) for k in range(3, 6))
        for i in range(len(deg_list)):
            for j in range(len(cm_list)):
                pose_aps[cls_id, i, j] = average_precision(pm[i, j, :], sc[i, j, :], gm[i, j, :])
    iou_aps[-1, :] = np.mean(iou_aps[1:-1, :], axis=0)
    pose_aps[-1] = np.mean(pose_aps[1:-1], axis=0)

    if log_dir is not None:                            # util.py:2802-2838: the two pickles next to the (omitted) figures
        os.makedirs(log_dir, exist_ok=True)
        with open(os.path.join(log_dir, "IoU_3D_AP_{}-{}.pkl".format(iou_list[0], iou_list[-1])), "wb") as f:
            pickle.dump({"thres_list": iou_list, "aps": iou_aps}, f)
        prefix = "Pose_Only_" if use_matches_for_pose else "Pose_Detection_"
        name = prefix + "AP_{}-{}degree_{}-{}cm.pkl".format(deg_list[0], deg_list[-2], cm_list[0], cm_list[-2])
        with open(os.path.join(log_dir, name), "wb") as f:
            pickle.dump({"degree_thres": deg_list, "shift_thres_list": cm_list, "aps": pose_aps}, f)
    return iou_aps, pose_aps


def summary_lines(iou_aps: np.ndarray, pose_aps: np.ndarray, synset_names: Sequence[str], iou_thresholds, degree_thresholds,
                  shift_thresholds) -> List[str]:
    """The mean rows the reference prints at the end (util.py:2929-2948)."""
    iou_list = list(iou_thresholds)
    deg_list, cm_list = list(degree_thresholds) + [360], list(shift_thresholds) + [100]
    lines = []
    for want in (0.25, 0.5):
        near = [k for k, v in enumerate(iou_list) if abs(v - want) < 1e-9]
        if near:
            lines.append("3D IoU at {:d}: {:.1f}".format(int(want * 100), iou_aps[-1, near[0]] * 100))
    for i, d in enumerate(deg_list):
        for j, c in enumerate(cm_list):
            lines.append("{} degree, {}cm: {:.1f}".format(d, c, pose_aps[-1, i, j] * 100))
    return lines
