"""Host-side result formats around the hot path (cppf2_b200/results.py; eval.py:103-151, 399): the detection pickles the frame
loop reads and the per-frame pickles it writes.  No GPU: the estimator is a stand-in that records what it was asked."""
import pickle
from types import SimpleNamespace

import numpy as np
import pytest

from cppf2_b200 import results as R


def _frame(path, cls_ids, h=48, w=64):
    n = len(cls_ids)
    masks = np.zeros((h, w, n), dtype=bool)
    for i in range(n):
        masks[4 + 6 * i: 10 + 6 * i, 8:40, i] = True
    return dict(image_path=path, pred_bboxes=np.tile(np.array([4, 8, 30, 40]), (n, 1)), pred_masks=masks,
                pred_class_ids=np.array(cls_ids), pred_scores=np.ones(n), gt_RTs=np.stack([np.eye(4)] * 2),
                gt_scales=np.ones((2, 3)), gt_class_ids=np.array([1, 5]))


class _FakeEstimator:
    def __init__(self, cats):
        self.models = {c: {} for c in cats}
        self.calls = []

    def estimate_frame(self, depth, masks, categories, intrinsics, desc_fn=None, depth_div=1000.0, frame_seed=0):
        self.calls.append(dict(depth=depth, masks=masks, cats=list(categories), K=np.asarray(intrinsics), div=depth_div,
                               seed=frame_seed, descs=None if desc_fn is None else [desc_fn(j, np.arange(3)) for j in range(len(masks))]))
        out = []
        for j, c in enumerate(categories):
            if c == "bowl":
                out.append(None)                       # an instance the reference's guards skip
                continue
            RT = np.eye(4)
            RT[:3, :3] *= 0.25 + j
            RT[:3, 3] = [0.1 * j, 0.2, 0.9]
            out.append(SimpleNamespace(RT=RT, scale=np.array([0.5, 0.6, 0.62]) / 1.0))
        return out


def test_load_results_flattens_and_defaults_handle_visibility(tmp_path):
    a = _frame("data/real/test/scene_1/0000", [1, 5])
    b = [_frame("data/real/test/scene_1/0001", [6]), dict(_frame("data/real/test/scene_2/0000", [2]), gt_handle_visibility=np.array([1, 0]))]
    pickle.dump(b, open(tmp_path / "results_b.pkl", "wb"))
    pickle.dump(a, open(tmp_path / "results_a.pkl", "wb"))
    pickle.dump({"x": 1}, open(tmp_path / "other.pkl", "wb"))
    res = R.load_results(tmp_path)
    assert [r["image_path"] for r in res] == ["data/real/test/scene_1/0000", "data/real/test/scene_1/0001", "data/real/test/scene_2/0000"]
    assert np.array_equal(res[0]["gt_handle_visibility"], np.ones(2, dtype=res[0]["gt_class_ids"].dtype))
    assert np.array_equal(res[2]["gt_handle_visibility"], [1, 0])
    with pytest.raises(FileNotFoundError):
        R.load_results(tmp_path / "empty")
    bad = dict(_frame("data/real/test/s/1", [1]), gt_handle_visibility=np.array([1]))
    (tmp_path / "bad").mkdir()
    pickle.dump(bad, open(tmp_path / "bad" / "results_x.pkl", "wb"))
    with pytest.raises(ValueError):
        R.load_results(tmp_path / "bad")


def test_paths_follow_the_reference():
    res = dict(image_path="data/real/test/scene_3/0042")
    stem = R.image_stem(res)
    assert stem == "NOCS/real_test/scene_3/0042"
    assert R.output_path("out", stem) == "out/real_test_scene_3_0042.pkl"          # '_'.join(path.split('/')[1:]) + '.pkl'


def test_run_results_fills_and_dumps(tmp_path):
    frames = [_frame("data/real/test/scene_1/0000", [1, 0, 2, 5]), _frame("data/real/test/scene_1/0001", [3])]
    est = _FakeEstimator(["bottle", "bowl", "laptop"])         # no camera heads: detection of class 3 is left untouched
    depths = {}

    def read_depth(path):
        depths[path] = np.full((48, 64), 900, np.uint16)
        return depths[path]

    seen = []
    out = R.run_results(est, frames, out_dir=tmp_path / "o", read_depth=read_depth,
                        desc_fn=lambda res, i, pix: seen.append((res["image_path"], i, len(pix))) or np.zeros((len(pix), 1024), np.float32))
    assert list(depths) == ["NOCS/real_test/scene_1/0000_depth.png", "NOCS/real_test/scene_1/0001_depth.png"]
    assert len(est.calls) == 1                                                    # frame 2 has no runnable detection
    call = est.calls[0]
    assert call["cats"] == ["bottle", "bowl", "laptop"] and call["div"] == 1000.0 and call["seed"] == 0
    assert np.array_equal(call["K"], R.REAL275_INTRINSICS)
    assert all(m.dtype == bool and m.shape == (48, 64) and m.flags["C_CONTIGUOUS"] for m in call["masks"])
    assert np.array_equal(call["masks"][2], frames[0]["pred_masks"][:, :, 3])
    assert seen == [("data/real/test/scene_1/0000", 0, 3), ("data/real/test/scene_1/0000", 2, 3), ("data/real/test/scene_1/0000", 3, 3)]
    r0 = out[0]
    assert r0["pred_RTs"].shape == (4, 4, 4) and r0["pred_scales"].shape == (4, 3)
    assert np.array_equal(r0["pred_RTs"][1], np.eye(4)) and np.array_equal(r0["pred_scales"][1], np.ones(3))   # background class
    assert np.array_equal(r0["pred_RTs"][2], np.eye(4))                                                        # skipped (None)
    assert np.allclose(r0["pred_RTs"][0][:3, 3], [0.0, 0.2, 0.9]) and np.allclose(r0["pred_RTs"][3][:3, 3], [0.2, 0.2, 0.9])
    assert np.allclose(r0["pred_scales"][3], [0.5, 0.6, 0.62])
    assert np.array_equal(out[1]["pred_RTs"], np.eye(4)[None])
    dumped = pickle.load(open(tmp_path / "o" / "real_test_scene_1_0000.pkl", "rb"))
    assert np.array_equal(dumped["pred_RTs"], r0["pred_RTs"]) and "gt_RTs" in dumped
    assert (tmp_path / "o" / "real_test_scene_1_0001.pkl").exists()


def test_degree_cm_error_matches_the_reference_formula():
    ang = np.radians(10.0)
    Ry = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(ang), -np.sin(ang)], [0, np.sin(ang), np.cos(ang)]])
    A, B = np.eye(4), np.eye(4)
    A[:3, :3] = 0.3 * Ry
    B[:3, :3] = 0.7 * np.eye(3)
    B[:3, 3] = [0.03, 0.0, 0.04]
    th, sh = R.degree_cm_error(A, B, symmetric_y=False)
    assert abs(th - 10.0) < 1e-6 and abs(sh - 5.0) < 1e-9
    th, _ = R.degree_cm_error(A, B, symmetric_y=True)          # a rotation about y is free for bottle / bowl / can
    assert th < 1e-5
    A[:3, :3] = 0.3 * Rx
    th, _ = R.degree_cm_error(A, B, symmetric_y=True)
    assert abs(th - 10.0) < 1e-6
