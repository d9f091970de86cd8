"""CPU: dataset-level golden.  tests/golden/dataset_items.npz holds what the reference's OWN ``ArgoverseIntentNetDataset``
(unmodified, run by oracle/make_golden_dataset.py) returns for every sample of the synthetic AV2-format log under
tests/golden/mini_av2/.  Here the host side of this repo reads the same directory (scenario scan, sliding windows, feather
decode with the reference's skip rules, pose prep) and the ORACLE reproduces the reference's tensors from it."""
import random

import numpy as np

from golden_util import GOLD, dense, load
from oracle import c_oracle as co

ROOT = GOLD / "mini_av2"


def _frames(bg):
    import pyarrow.feather as feather
    seqs = bg.av2_dataset.make_sequences(bg.av2_dataset.find_scenarios(ROOT))
    out = []
    for s in seqs:
        poses = feather.read_feather(f"{s['log_dir']}/city_SE3_egovehicle.feather")
        out.append((s, bg.sweep_io.load_frame(s["log_dir"], s["sweep_ts_list"], s["current_ts_ns"], poses, layout="columns", use_cache=False)))
    return out


def test_scan_and_windows_match_the_reference_dataset():
    import bevgen_b200 as bg
    z = load("dataset_items.npz")
    sc = bg.av2_dataset.find_scenarios(ROOT)
    assert [s["log_id"] for s in sc] == ["minilog0000"] and sc[0]["map_json_path"].endswith("PIT_city_00000.json")
    seqs = bg.av2_dataset.make_sequences(sc)
    assert [s["current_ts_ns"] for s in seqs] == z["eval_current_ts"].tolist()
    assert [s["sweep_ts_list"] for s in seqs] == z["eval_sweep_ts"].tolist()
    assert bg.av2_dataset.find_scenarios(ROOT / "nope") == []
    assert bg.av2_dataset.make_sequences(sc, num_sweeps=13) == []          # fewer sweeps than the window: no samples (dataset.py:189)


def test_oracle_reproduces_the_reference_datasets_items():
    import json
    import bevgen_b200 as bg
    z = load("dataset_items.npz")
    frames = _frames(bg)
    assert len(frames) == 3
    md = json.loads(open(frames[0][0]["map_json_path"]).read())
    flat = co.flatten_map(md)
    random.seed(int(z["seed"]))
    params = [bg.augment.draw_aug_params(random) for _ in frames]           # the reference's draws, sample by sample
    assert any(p.needs_gather for p in params)
    for i, (seq, fr) in enumerate(frames):
        assert np.diff(fr.sweep_offsets)[[4 - i, 6 - i]].tolist() == [0, 0]    # log sweeps 4 (no pose row) and 6 (empty table) are None
        pts4 = np.stack([c.astype(np.float32) for c in fr.columns], 1)
        rel = bg.pose.sweep_rel_transforms(fr.current_pose, fr.sweep_poses)
        lid = co.lidar_bev_fused(pts4, fr.sweep_offsets, rel, 10)
        assert np.array_equal(lid, dense(z, f"eval{i}_lidar")), i
        T3 = bg.pose.map_matrix_from_pose(fr.current_pose[0], fr.current_pose[1], fr.current_pose[3:7])
        mp = co.map_raster_flat(flat, T3)
        want_m = np.unpackbits(z[f"eval{i}_map"])[: 9 * 400 * 720].reshape(9, 400, 720).astype(np.float32)
        assert np.array_equal(mp, want_m), i
        p = params[i]
        for g, name in ((lid, "lidar"), (mp, "mapf")):
            o = g
            if p.flip: o = co.flip_w(o)
            if p.angle_deg is not None: o = co.rotate(o, p.angle_deg)
            if p.scale is not None: o = co.scale(o, p.scale)
            o = np.ascontiguousarray(o).copy()
            if p.rects: co.dropout(o, p.rects)
            assert np.array_equal(o, dense(z, f"train{i}_{name}")), (i, name)
        boxes, intents = z[f"eval{i}_boxes"].copy(), z[f"eval{i}_intents"].copy()
        bg.augment.apply_gt(p, boxes, intents)
        assert np.array_equal(boxes.astype(np.float32), z[f"train{i}_boxes"]) and np.array_equal(intents, z[f"train{i}_intents"])


def test_annotation_gt_mirrors_prepare_gt_for_frame():
    """AnnotationGT = utils.py:184-225 on annotations_with_intent.feather.  strict_reference reproduces what the reference
    returned in this container (the golden); the default converts the quaternion to float64 first, as the reference's pinned
    scipy would, and yields the filtered rows' boxes"""
    import pandas as pd
    from scipy.spatial.transform import Rotation as R
    import bevgen_b200 as bg
    z = load("dataset_items.npz")
    seqs = bg.av2_dataset.make_sequences(bg.av2_dataset.find_scenarios(ROOT))
    strict, dflt = bg.av2_dataset.AnnotationGT(strict_reference=True), bg.av2_dataset.AnnotationGT()
    ann = pd.read_feather(ROOT / "minilog0000" / "annotations_with_intent.feather")
    for i, s in enumerate(seqs):
        g = strict(s)
        assert np.array_equal(g["boxes_xywha"], z[f"eval{i}_boxes"]) and np.array_equal(g["intentions"], z[f"eval{i}_intents"])
        d = dflt(s)
        rows = ann[(ann["timestamp_ns"] == s["current_ts_ns"]) & ann["category"].isin(bg.constants.VEHICLE_CATEGORIES) & (ann["heuristic_intent"] != -1)]
        assert len(rows) > 0 and d["boxes_xywha"].shape == (len(rows), 5) and d["boxes_xywha"].dtype == np.float32
        assert d["intentions"].tolist() == rows["heuristic_intent"].astype(int).tolist() and d["intentions"].dtype == np.int64
        q = rows[["qx", "qy", "qz", "qw"]].to_numpy(np.float64)
        want = np.stack([rows["tx_m"], rows["ty_m"], rows["width_m"].abs(), rows["length_m"].abs(), R.from_quat(q).as_euler("xyz")[:, 2]], 1)
        assert np.array_equal(d["boxes_xywha"], want.astype(np.float32))
    assert dflt({"log_id": "none", "log_dir": str(ROOT / "nope"), "current_ts_ns": 0})["boxes_xywha"].shape == (0, 5)
