"""Generates tests/golden/*.npz by running the UNMODIFIED reference functions (loaded read-only from
/root/reference, see oracle/ref_loader.py) on small seeded inputs.  Build container only:

    python -m oracle.make_golden

The reference ships no golden vectors of its own, so these fixtures -- inputs AND the
reference's outputs -- are the committed pin that travels to the GPU box.  cv2.resize is run with IPP
disabled (the IPP-dispatched path is CPU dependent); everything else uses library defaults.
Versions used are recorded in tests/golden/META.json.
"""
from __future__ import annotations

import hashlib
import json
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"


def sparse(a: np.ndarray) -> dict:
    flat = np.ascontiguousarray(a).reshape(-1)
    idx = np.flatnonzero(flat).astype(np.int32)
    return {"idx": idx, "val": flat[idx].copy(), "shape": np.asarray(a.shape, np.int64),
            "sha256": np.frombuffer(hashlib.sha256(flat.tobytes()).digest(), np.uint8)}


def pref(d: dict, p: str) -> dict:
    return {f"{p}_{k}": v for k, v in d.items()}


def main() -> None:
    import cv2
    import pandas as pd
    import scipy
    import torch
    from oracle import ref_loader
    import bevgen_b200 as bg                      # only the pure-host synthetic generators are used here

    u = ref_loader.utils()
    GOLD.mkdir(parents=True, exist_ok=True)
    cv2.ipp.setUseIPP(False)

    # ---- transform_points + create_intentnet_lidar_bev (utils.py:27-33, :62-106) -------------------------
    f = bg.synthetic.frame_c2(frame_idx=7, n=1500, S=10)
    pl, il = [], []
    for s in range(10):
        sw = f.sweep(s)
        pl.append(u.transform_points(sw[:, :3], f.rel_tf[s]))
        il.append(sw[:, 3].astype(np.float32))
    pl[3] = None; il[3] = None                    # missing sweep
    bev = u.create_intentnet_lidar_bev(pl, il)
    fused_pts = f.points4.copy()
    off = f.sweep_offsets.copy()
    # the fused API expresses "sweep 3 missing" as an empty range
    keep = np.ones(fused_pts.shape[0], bool); keep[off[3]:off[4]] = False
    off2 = off.copy(); off2[4:] -= (off[4] - off[3])
    np.savez_compressed(GOLD / "lidar_c2_small.npz", points4=fused_pts[keep], sweep_offsets=off2, rel_tf=f.rel_tf,
                        sweep_poses=f.sweep_poses, current_pose=f.current_pose,
                        transformed_sweep0=pl[0], **pref(sparse(bev), "bev"))

    # bin-edge / non-finite points, one sweep, identity pose
    rng = np.random.default_rng(7)
    n = 6000
    xs = (60.0 - 0.2 * np.arange(-5, 406)).astype(np.float32)
    ys = (-72.0 + 0.2 * np.arange(-5, 726)).astype(np.float32)
    zs = (-2.0 + 0.2 * np.arange(-2, 32)).astype(np.float32)
    p = np.empty((n, 4), np.float32)
    p[:, 0] = rng.choice(xs, n); p[:, 1] = rng.choice(ys, n); p[:, 2] = rng.choice(zs, n); p[:, 3] = rng.integers(1, 256, n)
    p[:20, 0] = np.nan; p[20:40, 1] = np.inf; p[40:60, 2] = -np.inf; p[60:80, 0] = 1e30; p[80:100, 3] = 0.0
    ego = u.transform_points(p[:, :3], np.eye(4))
    bev_e = u.create_intentnet_lidar_bev([ego], [p[:, 3].copy()], 1)
    np.savez_compressed(GOLD / "lidar_edges.npz", points4=p, **pref(sparse(bev_e), "bev"))

    # ---- rasterize_map_ego_centric (utils.py:108-182) ---------------------------------------------------------
    md = bg.synthetic.synthetic_map(seed=11, rows=9, cols=8, n_cross=5)
    mp_path = GOLD / "map_small.json"
    mp_path.write_text(json.dumps(md))
    from scipy.spatial.transform import Rotation as R
    poses, outs = [], []
    prng = np.random.default_rng(3)
    for k in range(3):
        tx, ty = 1500 + prng.uniform(-40, 40), -700 + prng.uniform(-40, 40)
        q = R.from_euler("xyz", [0.01, -0.02, prng.uniform(-3.1, 3.1)]).as_quat()
        ser = pd.Series(dict(tx_m=tx, ty_m=ty, tz_m=0.0, qx=q[0], qy=q[1], qz=q[2], qw=q[3]))
        poses.append([tx, ty, 0.0, *q])
        outs.append(np.packbits(u.rasterize_map_ego_centric(str(mp_path), ser).astype(np.uint8), axis=None))
    np.savez_compressed(GOLD / "map_small.npz", poses=np.asarray(poses), packed=np.stack(outs))

    # ---- augmentation (utils.py:394-517), seeded through the reference's own augment_bev --------------------
    grng = np.random.default_rng(21)
    lid = np.zeros((2, 400, 720), np.float32)
    ii = grng.integers(0, lid.size, 5000); lid.reshape(-1)[ii] = grng.integers(1, 256, ii.size)
    mapb = (grng.random((1, 400, 720)) < 0.02).astype(np.float32)
    boxes = np.array([[10.0, -3.0, 2.0, 4.5, 0.3], [30.0, 8.0, 2.1, 5.0, -2.9], [-5.0, 20.0, 1.9, 4.2, 1.4]], np.float32)
    intents = np.array([1, 4, 0], np.int64)
    aug = {"lidar_in_idx": sparse(lid)["idx"], "lidar_in_val": sparse(lid)["val"], "map_in_idx": sparse(mapb)["idx"],
           "boxes": boxes, "intents": intents}
    seeds = [0, 1, 2, 3, 5, 8, 13, 21, 34, 58, 144]        # cover flip / rotate / scale / dropout combinations
    for sd in seeds:
        random.seed(sd)
        lo, mo, gt = u.augment_bev(lid, mapb, {"boxes_xywha": torch.from_numpy(boxes.copy()), "intentions": torch.from_numpy(intents.copy())})
        aug.update(pref(sparse(lo), f"s{sd}_lidar")); aug.update(pref(sparse(mo), f"s{sd}_map"))
        aug[f"s{sd}_boxes"] = gt["boxes_xywha"].numpy(); aug[f"s{sd}_intents"] = gt["intentions"].numpy()
    aug["seeds"] = np.asarray(seeds)
    np.savez_compressed(GOLD / "augment_seeded.npz", **aug)

    meta = {"numpy": np.__version__, "cv2": cv2.__version__, "scipy": scipy.__version__, "pandas": pd.__version__,
            "torch": torch.__version__, "cv2_ipp": False,
            "reference": "Nadeem202020/VisionTransformer-Intention-Prediction (utils.py, constants.py; unmodified, loaded read-only)",
            "generator": "python -m oracle.make_golden"}
    (GOLD / "META.json").write_text(json.dumps(meta, indent=1))
    for fpath in sorted(GOLD.iterdir()):
        print(f"{fpath.name:28s} {fpath.stat().st_size/1024:8.1f} KiB")


if __name__ == "__main__":
    main()
