"""Generates tests/golden/lidar_av2_columns.npz: the reference's OWN result for sweeps in AV2's on-disk dtypes, read the way
its Dataset reads them -- the literal statements of dataset.py:311-340, executed against the UNMODIFIED reference functions
(loaded read-only from /root/reference, see oracle/ref_loader.py):

    sweep_df = pd.read_feather(path, columns=['x','y','z','intensity'])     dataset.py:311   (float16 x 3, uint8)
    pts      = sweep_df[['x','y','z']].values                               dataset.py:320   (float16 (N,3))
    inten    = sweep_df['intensity'].values.astype(np.float32)              dataset.py:321
    rel_tf   = ego_SE3_world @ sw_tf_world_ego                              dataset.py:339
    transform_points(pts, rel_tf)                                           dataset.py:340 -> utils.py:27-33
    create_intentnet_lidar_bev(points_list, intensity_list)                 dataset.py:347 -> utils.py:62-106

The fixture stores the column arrays exactly as they sit in the feather files plus the reference's grid, and pins the
column-format CUDA entry points (bevgen_lidar_bev_fused_cols, 7 B/point) and the oracle against the reference itself.
Build container only:   python -m oracle.make_golden_av2
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.make_golden import GOLD, pref, sparse  # noqa: E402


def main() -> None:
    import pandas as pd
    from scipy.spatial.transform import Rotation as R
    from oracle import ref_loader
    import bevgen_b200 as bg                      # only the pure-host synthetic generators are used here

    u = ref_loader.utils()
    S = 10
    f = bg.synthetic.frame_c2(frame_idx=91, n=1801, S=S)          # odd sweep length: unaligned column starts
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        cols = [[], [], [], []]
        counts = []
        for s in range(S):
            sw = f.sweep(s)
            n = sw.shape[0] if s != 4 else 0                        # sweep 4: empty table -> None sweep (dataset.py:312-314)
            df = pd.DataFrame({"x": sw[:n, 0].astype(np.float16), "y": sw[:n, 1].astype(np.float16), "z": sw[:n, 2].astype(np.float16),
                               "intensity": sw[:n, 3].astype(np.uint8), "laser_number": np.zeros(n, np.uint8),
                               "offset_ns": np.zeros(n, np.int32)})
            df.to_feather(tmp / f"{s}.feather")
        # ---- the reference's loop, dataset.py:293-347 ----
        cp = f.current_pose
        world_SE3_ego = np.eye(4)
        world_SE3_ego[:3, :3] = R.from_quat(cp[3:7]).as_matrix()
        world_SE3_ego[:3, 3] = cp[:3]
        ego_SE3_world = np.linalg.inv(world_SE3_ego)
        points_list, intensity_list, rels = [], [], []
        for s in range(S):
            sweep_df = pd.read_feather(tmp / f"{s}.feather", columns=["x", "y", "z", "intensity"])
            sp = f.sweep_poses[s]
            sw_tf = np.eye(4)
            sw_tf[:3, :3] = R.from_quat(sp[3:7]).as_matrix()
            sw_tf[:3, 3] = sp[:3]
            rel_tf = ego_SE3_world @ sw_tf
            rels.append(rel_tf)
            if sweep_df.empty:
                points_list.append(None); intensity_list.append(None); counts.append(0)
                continue
            pts_world = sweep_df[["x", "y", "z"]].values
            intensity = sweep_df["intensity"].values.astype(np.float32)
            assert pts_world.dtype == np.float16 and sweep_df["intensity"].dtype == np.uint8
            points_list.append(u.transform_points(pts_world, rel_tf))
            intensity_list.append(intensity)
            for k, c in enumerate(("x", "y", "z", "intensity")):
                cols[k].append(sweep_df[c].values)
            counts.append(len(sweep_df))
        bev = u.create_intentnet_lidar_bev(points_list, intensity_list)
    off = np.zeros(S + 1, np.int32)
    off[1:] = np.cumsum(counts)
    np.savez_compressed(GOLD / "lidar_av2_columns.npz", x=np.concatenate(cols[0]), y=np.concatenate(cols[1]), z=np.concatenate(cols[2]),
                        intensity=np.concatenate(cols[3]), sweep_offsets=off, rel_tf=np.stack(rels), sweep_poses=f.sweep_poses,
                        current_pose=f.current_pose, **pref(sparse(bev), "bev"))
    print(f"lidar_av2_columns.npz {(GOLD / 'lidar_av2_columns.npz').stat().st_size / 1024:.1f} KiB, {int(off[-1])} points, "
          f"{np.count_nonzero(bev)} non-zero voxels")


if __name__ == "__main__":
    main()
