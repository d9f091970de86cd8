"""Synthetic frames / maps for the five BASELINE.json configs.

Everything is drawn from ``np.random.default_rng(1000*config + frame_idx)``.  Points are
float32 rows (x, y, z, intensity) in the *sweep* frame, concatenated over sweeps with a
CSR offset array, which is the layout the fused device API consumes.
"""
from __future__ import annotations

import dataclasses
import json
from typing import Optional

import numpy as np
from scipy.spatial.transform import Rotation as R

from . import pose as _pose
from .constants import LIDAR_SWEEPS


@dataclasses.dataclass
class Frame:
    points4: np.ndarray          # float32 [sum N, 4]  x,y,z,intensity (sweep frame)
    sweep_offsets: np.ndarray    # int32   [S+1]
    sweep_poses: np.ndarray      # float64 [S, 7]  tx,ty,tz,qx,qy,qz,qw  (oldest -> newest)
    current_pose: np.ndarray     # float64 [7]
    rel_tf: np.ndarray           # float64 [S,4,4]  (dataset.py:339)
    map_data: Optional[dict] = None

    @property
    def num_sweeps(self) -> int:
        return len(self.sweep_offsets) - 1

    @property
    def num_points(self) -> int:
        return int(self.sweep_offsets[-1])

    def sweep(self, s: int) -> np.ndarray:
        return self.points4[self.sweep_offsets[s]:self.sweep_offsets[s + 1]]

    def pose_series(self):
        """pandas Series shaped like the reference's ego-pose row (utils.py:121-122)."""
        import pandas as pd
        keys = ["tx_m", "ty_m", "tz_m", "qx", "qy", "qz", "qw"]
        return pd.Series(dict(zip(keys, self.current_pose.tolist())))


def _poses(S: int, origin=(0.0, 0.0, 0.0), yaw0: float = 0.0) -> np.ndarray:
    out = np.zeros((S, 7), np.float64)
    for k in range(S):
        out[k, :3] = [origin[0] + 0.5 * k, origin[1] + 0.01 * k, origin[2]]
        out[k, 3:] = R.from_euler("xyz", [1e-3 * k, -1e-3 * k, yaw0 + 0.002 * k]).as_quat()
    return out


def _finish(pts_per_sweep, poses, map_data=None) -> Frame:
    offs = np.zeros(len(pts_per_sweep) + 1, np.int32)
    offs[1:] = np.cumsum([p.shape[0] for p in pts_per_sweep])
    pts = (np.concatenate(pts_per_sweep, 0) if pts_per_sweep else np.zeros((0, 4), np.float32)).astype(np.float32)
    cur = poses[-1].copy()
    return Frame(pts, offs, poses, cur, _pose.sweep_rel_transforms(cur, poses), map_data)


def frame_c1(frame_idx: int = 0, n: int = 100_000) -> Frame:
    """Config 1: one loaded sweep, uniform box, identity pose."""
    rng = np.random.default_rng(1000 * 1 + frame_idx)
    p = np.empty((n, 4), np.float32)
    p[:, 0] = rng.uniform(-30, 70, n)
    p[:, 1] = rng.uniform(-80, 80, n)
    p[:, 2] = rng.uniform(-3, 5, n)
    p[:, 3] = rng.integers(0, 256, n)
    poses = np.zeros((1, 7)); poses[0, 6] = 1.0
    return _finish([p], poses)


def frame_c2(frame_idx: int = 0, n: int = 100_000, S: int = LIDAR_SWEEPS, config: int = 2,
             origin=(0.0, 0.0, 0.0), yaw0: float = 0.0) -> Frame:
    """Config 2: S sweeps, N(0, 35 m) xy, synthetic moving ego."""
    rng = np.random.default_rng(1000 * config + frame_idx)
    sweeps = []
    for _ in range(S):
        p = np.empty((n, 4), np.float32)
        p[:, 0] = rng.normal(0, 35, n)
        p[:, 1] = rng.normal(0, 35, n)
        p[:, 2] = rng.uniform(-3, 5, n)
        p[:, 3] = rng.integers(0, 256, n)
        sweeps.append(p)
    return _finish(sweeps, _poses(S, origin, yaw0))


def frame_c4(frame_idx: int = 0, n: int = 100_000, S: int = LIDAR_SWEEPS) -> Frame:
    """Config 4: dense / skewed -- S*n points in 4 clusters, sigma 0.3 m (atomic contention)."""
    rng = np.random.default_rng(1000 * 4 + frame_idx)
    centres = np.array([[10.0, -5.0], [25.0, 12.0], [-4.0, 30.0], [40.0, -40.0]])
    sweeps = []
    for _ in range(S):
        c = centres[rng.integers(0, 4, n)]
        p = np.empty((n, 4), np.float32)
        p[:, 0] = c[:, 0] + rng.normal(0, 0.3, n)
        p[:, 1] = c[:, 1] + rng.normal(0, 0.3, n)
        p[:, 2] = rng.uniform(-1, 1, n)
        p[:, 3] = rng.integers(0, 256, n)
        sweeps.append(p)
    return _finish(sweeps, _poses(S))


def sliding_log(num_sweeps: int, n: int = 100_000, seed: int = 0, dense: bool = False):
    """A synthetic LOG: ``num_sweeps`` consecutive sweeps (config-2 point distribution, or config-4 clusters with
    ``dense``; coordinates already rounded to float16 values, as AV2 stores them) with a moving ego, from which the
    reference's Dataset would cut 10-sweep sliding windows (dataset.py:174-211).  Returns (sweeps, poses): a list of
    float32 ``[n,4]`` arrays and float64 ``[num_sweeps,7]``."""
    rng = np.random.default_rng(1000 * 6 + seed)
    centres = np.array([[10.0, -5.0], [25.0, 12.0], [-4.0, 30.0], [40.0, -40.0]])
    sweeps = []
    for _ in range(num_sweeps):
        p = np.empty((n, 4), np.float32)
        if dense:
            c = centres[rng.integers(0, 4, n)]
            p[:, 0] = c[:, 0] + rng.normal(0, 0.3, n); p[:, 1] = c[:, 1] + rng.normal(0, 0.3, n); p[:, 2] = rng.uniform(-1, 1, n)
        else:
            p[:, 0] = rng.normal(0, 35, n); p[:, 1] = rng.normal(0, 35, n); p[:, 2] = rng.uniform(-3, 5, n)
        p[:, 3] = rng.integers(0, 256, n)
        p[:, :3] = p[:, :3].astype(np.float16).astype(np.float32)
        sweeps.append(p)
    return sweeps, _poses(num_sweeps)


def window_frame(sweeps, poses: np.ndarray, first: int, S: int = LIDAR_SWEEPS, map_data=None) -> Frame:
    """The sample whose sweeps are ``first .. first + S - 1`` of a log (current pose = the newest sweep's, as in
    dataset.py:285-301 where the sample's timestamp is its last sweep's)."""
    return _finish([sweeps[first + k] for k in range(S)], poses[first:first + S].copy(), map_data)


def quantize_f16(frame: Frame) -> Frame:
    """Round the point coordinates to float16-representable values (what AV2 sweeps hold on disk: float16 x,y,z and
    uint8 intensity), in place.  Such frames can be packed at 8 B/point (FrameBatch.pack(points_dtype=np.float16))."""
    frame.points4 = frame.points4.astype(np.float16).astype(np.float32)
    return frame


# ----------------------------------------------------------------------------- maps
EGO_XY = (1500.0, -700.0)
EGO_YAW = 0.6
_LANE_TYPES = ["VEHICLE", "BUS", "BIKE"]
_MARKS = ["DASHED_WHITE", "SOLID_WHITE", "SOLID_YELLOW", "NONE"]


def synthetic_map(seed: int = 3000, rows: int = 25, cols: int = 17, n_cross: int = 20) -> dict:
    """Map dict in the reference's JSON schema (utils.py:118-180): rows x cols lane segments on a
    15 x 12 m lattice around the ego, random headings, 10 vertices per boundary, half-width
    1.8 m; crosswalk quads under the key 'polygon' (the key utils.py:176 reads)."""
    rng = np.random.default_rng(seed)
    cy, sy = np.cos(EGO_YAW), np.sin(EGO_YAW)
    lanes = {}
    for i in range(rows):
        for j in range(cols):
            ex, ey = (i - rows // 3) * 15.0 + rng.uniform(-3, 3), (j - cols // 2) * 12.0 + rng.uniform(-3, 3)
            wx, wy = EGO_XY[0] + cy * ex - sy * ey, EGO_XY[1] + sy * ex + cy * ey
            h = rng.uniform(0, 2 * np.pi)
            curv = rng.uniform(-0.02, 0.02)
            s = np.linspace(0, 14.0, 10)
            heading = h + curv * s
            cxs, cys = wx + np.cumsum(np.cos(heading)) * (s[1] - s[0]), wy + np.cumsum(np.sin(heading)) * (s[1] - s[0])
            nx, ny = -np.sin(heading), np.cos(heading)
            left = [{"x": float(a), "y": float(b), "z": 0.0} for a, b in zip(cxs + 1.8 * nx, cys + 1.8 * ny)]
            right = [{"x": float(a), "y": float(b), "z": 0.0} for a, b in zip(cxs - 1.8 * nx, cys - 1.8 * ny)]
            lanes[str(len(lanes))] = {
                "id": len(lanes),
                "is_intersection": bool(rng.random() < 0.2),
                "lane_type": _LANE_TYPES[int(rng.integers(0, 3))],
                "left_lane_boundary": left,
                "right_lane_boundary": right,
                "left_lane_mark_type": _MARKS[int(rng.integers(0, 4))],
                "right_lane_mark_type": _MARKS[int(rng.integers(0, 4))],
            }
    cross = {}
    for k in range(n_cross):
        ex, ey = rng.uniform(-15, 55), rng.uniform(-60, 60)
        wx, wy = EGO_XY[0] + cy * ex - sy * ey, EGO_XY[1] + sy * ex + cy * ey
        h = rng.uniform(0, 2 * np.pi)
        ux, uy, vx, vy = np.cos(h) * 3.0, np.sin(h) * 3.0, -np.sin(h) * 1.5, np.cos(h) * 1.5
        quad = [(wx - ux - vx, wy - uy - vy), (wx + ux - vx, wy + uy - vy),
                (wx + ux + vx, wy + uy + vy), (wx - ux + vx, wy - uy + vy)]
        cross[str(k)] = {"id": k, "polygon": [{"x": float(a), "y": float(b), "z": 0.0} for a, b in quad]}
    return {"lane_segments": lanes, "pedestrian_crossings": cross}


def frame_c3(frame_idx: int = 0, n: int = 100_000, S: int = LIDAR_SWEEPS, map_data: Optional[dict] = None) -> Frame:
    """Config 3: config-2 lidar (own seed stream) + the synthetic HD map; ego near (1500,-700), yaw 0.6."""
    f = frame_c2(frame_idx, n, S, config=3, origin=(EGO_XY[0], EGO_XY[1], 0.0), yaw0=EGO_YAW)
    f.map_data = map_data if map_data is not None else synthetic_map()
    return f


def write_map_json(map_data: dict, path) -> str:
    with open(path, "w") as fh:
        json.dump(map_data, fh)
    return str(path)
