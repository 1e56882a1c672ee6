"""mAP scoring of the per-frame pose records (cppf2_b200/scoring.py; eval.py:400-412 -> utils/util.py:2610-2955) against
tests/golden/map_eval.pkl, minted by oracle/make_map_golden.py from the UNMODIFIED reference scoring code on a seeded
synthetic result set (14 frames, misses, false positives, wrong classes, hidden mug handles).  Host code only."""
import copy
import os
import pickle

import numpy as np
import pytest

from cppf2_b200 import scoring as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "map_eval.pkl")

# The reference's box IoU goes through a plane-thickness clipper (utils/iou.py); ours clips with the same thickness but
# collects the polytope's vertices in another order, so the volumes agree to rounding, not bit for bit.
IOU_TOL = 1e-6


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN, "rb") as f:
        return pickle.load(f)


def test_box_pairs_match_reference(gold):
    worst = 0.0
    for p in gold["pairs"]:
        name = gold["synsets"][p["class_id"]]
        iou = S.box_iou_3d(p["RT_1"].copy(), p["RT_2"].copy(), p["scales_1"], p["scales_2"], p["handle_visibility"], name, name)
        worst = max(worst, abs(iou - p["iou"]))
        err = S.rotation_translation_error(p["RT_1"], p["RT_2"], p["class_id"], p["handle_visibility"], gold["synsets"])
        np.testing.assert_allclose(err, p["err"], rtol=0, atol=1e-9)
    assert worst <= IOU_TOL, worst
    assert gold["pairs"][0]["iou"] == 0.0 and gold["pairs"][4]["iou"] > 0.999        # the fixture's far-apart and identical boxes


def test_none_and_bad_last_row():
    assert S.box_iou_3d(None, np.eye(4), np.ones(3), np.ones(3), 1, "camera", "camera") == -1
    assert S.rotation_translation_error(None, np.eye(4), 3, 1, S.NOCS_SYNSETS) == -1
    bad = np.eye(4)
    bad[3, 3] = 2.0
    with pytest.raises(ValueError):
        S.rotation_translation_error(bad, np.eye(4), 3, 1, S.NOCS_SYNSETS)


def test_ap_tables_match_reference_pose_only(gold, tmp_path):
    """The call of eval.py:407-412: 101 IoU thresholds, 5/10/15 degrees x 5/10/15 cm, poses scored on the IoU > 0.1 matches."""
    res = copy.deepcopy(gold["results"])
    iou_aps, pose_aps = S.compute_degree_cm_mAP(res, gold["synsets"], str(tmp_path), degree_thresholds=[5, 10, 15],
                                                shift_thresholds=[5, 10, 15], iou_3d_thresholds=np.linspace(0, 1, 101),
                                                iou_pose_thres=0.1, use_matches_for_pose=True)
    assert iou_aps.shape == gold["iou_aps"].shape and pose_aps.shape == gold["pose_aps"].shape
    np.testing.assert_allclose(pose_aps, gold["pose_aps"], rtol=0, atol=1e-12)
    # an IoU within IOU_TOL of one of the 101 thresholds may flip one match; none does on this fixture
    np.testing.assert_allclose(iou_aps, gold["iou_aps"], rtol=0, atol=1e-12)
    names = sorted(os.listdir(tmp_path))
    assert names == ["IoU_3D_AP_0.0-1.0.pkl", "Pose_Only_AP_5-15degree_5-15cm.pkl"]
    with open(tmp_path / names[1], "rb") as f:
        d = pickle.load(f)
    assert d["degree_thres"] == [5, 10, 15, 360] and d["shift_thres_list"] == [5, 10, 15, 100]
    np.testing.assert_array_equal(d["aps"], pose_aps)
    for r, g in zip(res, gold["results"]):                 # the caller's dicts are not modified
        np.testing.assert_array_equal(r["pred_RTs"], g["pred_RTs"])
        np.testing.assert_array_equal(r["gt_RTs"], g["gt_RTs"])


def test_ap_tables_match_reference_detection(gold):
    iou_aps, pose_aps = S.compute_degree_cm_mAP(copy.deepcopy(gold["results"]), gold["synsets"], None, degree_thresholds=[5, 10],
                                                shift_thresholds=[2, 5], iou_3d_thresholds=[0.25, 0.5, 0.75],
                                                iou_pose_thres=0.1, use_matches_for_pose=False)
    np.testing.assert_allclose(iou_aps, gold["iou_aps_detection"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(pose_aps, gold["pose_aps_detection"], rtol=0, atol=1e-12)


def test_process_pool_gives_the_same_tables(gold):
    kw = dict(degree_thresholds=[5, 10], shift_thresholds=[5], iou_3d_thresholds=[0.25, 0.5], iou_pose_thres=0.25,
              use_matches_for_pose=True)
    a = S.compute_degree_cm_mAP(copy.deepcopy(gold["results"]), gold["synsets"], None, num_proc=1, **kw)
    b = S.compute_degree_cm_mAP(copy.deepcopy(gold["results"]), gold["synsets"], None, num_proc=3, **kw)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_empty_frames_and_missing_classes():
    empty = dict(gt_class_ids=np.zeros(0, np.int32), gt_RTs=np.zeros((0, 4, 4)), gt_scales=np.zeros((0, 3)),
                 gt_handle_visibility=np.zeros(0, np.int32), pred_class_ids=np.zeros(0, np.int32), pred_RTs=np.zeros((0, 4, 4)),
                 pred_scales=np.zeros((0, 3)), pred_scores=np.zeros(0))
    rt = np.eye(4)
    rt[:3, :3] *= 0.2
    rt[:3, 3] = [0, 0, 1]
    one = dict(gt_class_ids=np.array([3], np.int32), gt_RTs=rt[None].copy(), gt_scales=np.array([[0.5, 0.6, 0.62]]),
               gt_handle_visibility=np.array([1], np.int32), pred_class_ids=np.array([3], np.int32), pred_RTs=rt[None].copy(),
               pred_scales=np.array([[0.5, 0.6, 0.62]]), pred_scores=np.array([0.9]))
    missed = dict(one, pred_class_ids=np.zeros(0, np.int32), pred_RTs=np.zeros((0, 4, 4)), pred_scales=np.zeros((0, 3)),
                  pred_scores=np.zeros(0))
    iou_aps, pose_aps = S.compute_degree_cm_mAP([empty, one, missed], S.NOCS_SYNSETS, None, degree_thresholds=[5],
                                                shift_thresholds=[5], iou_3d_thresholds=[0.5], iou_pose_thres=0.5)
    cam = S.NOCS_SYNSETS.index("camera")
    assert iou_aps[cam, 0] == pytest.approx(0.5)           # one of the two ground truths found, at precision 1
    assert pose_aps[cam, 0, 0] == pytest.approx(0.5)
    assert iou_aps[1, 0] == 0.0 and pose_aps[1].sum() == 0.0


def test_summary_lines(gold):
    lines = S.summary_lines(gold["iou_aps"], gold["pose_aps"], gold["synsets"], np.linspace(0, 1, 101), [5, 10, 15], [5, 10, 15])
    assert lines[0].startswith("3D IoU at 25: ") and lines[1].startswith("3D IoU at 50: ")
    assert "10 degree, 5cm: {:.1f}".format(gold["pose_aps"][-1, 1, 0] * 100) in lines
