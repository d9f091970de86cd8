"""Sweep decode: AV2 LiDAR feather files -> packed point rows for the device path.

Mirrors the host loop of the reference's ``Dataset.__getitem__`` (dataset.py:303-342) up to -- but not including -- the
hot path: which sweeps count as missing (file absent, empty table, corrupt file, no pose row, invalid quaternion ->
``None`` sweep, dataset.py:307-334), which columns are read (x, y, z, intensity; dataset.py:311), and the pose rows
looked up per sweep timestamp (dataset.py:323-331).  Instead of transforming on the CPU it returns the raw sweep-frame
points packed as ``[N,4]`` rows (x, y, z, intensity) in the narrowest exact dtype:

* float16 when every column is exactly representable (AV2 stores x,y,z as float16 and intensity as uint8) -> 8 B/point;
* float32 otherwise.

Decoded sweeps are cached per path (one parse per file instead of one per sample that uses the sweep: with 10-sweep
sliding windows every file is otherwise read ten times per epoch).
"""
from __future__ import annotations

import dataclasses
from pathlib import Path
from typing import Dict, Optional, Sequence

import numpy as np

from . import pose as _pose

_COLUMNS = ["x", "y", "z", "intensity"]
_CACHE: Dict[str, Optional[np.ndarray]] = {}


def decode_sweep(path, use_cache: bool = True) -> Optional[np.ndarray]:
    """One sweep file -> ``[N,4]`` float16/float32 rows, or ``None`` for the reference's "skip this sweep" cases
    (missing file dataset.py:307-309, empty table :312-314, unreadable file :315-318)."""
    import pyarrow as pa
    import pyarrow.feather as feather
    key = str(path)
    if use_cache and key in _CACHE:
        return _CACHE[key]
    out: Optional[np.ndarray] = None
    p = Path(key)
    if p.is_file():
        try:
            table = feather.read_table(key, columns=_COLUMNS)
            if table.num_rows > 0:
                cols = [table.column(c).to_numpy(zero_copy_only=False) for c in _COLUMNS]
                xyz = np.stack(cols[:3], axis=1)
                inten = cols[3].astype(np.float32)               # dataset.py:321
                exact16 = xyz.dtype == np.float16 and np.array_equal(inten.astype(np.float16).astype(np.float32), inten)
                dt = np.float16 if exact16 else np.float32
                out = np.empty((table.num_rows, 4), dt)
                out[:, :3] = xyz
                out[:, 3] = inten
        except (pa.ArrowInvalid, OSError):
            print(f"Warning: Corrupt LiDAR sweep file: {key}. Skipping sweep.")
            out = None
    if use_cache:
        _CACHE[key] = out
    return out


def clear_cache() -> None:
    _CACHE.clear()


@dataclasses.dataclass
class SweepFrame:
    """What pipeline.FrameBatch.pack consumes (same fields as synthetic.Frame)."""
    points4: np.ndarray
    sweep_offsets: np.ndarray
    sweep_poses: np.ndarray
    current_pose: np.ndarray
    gt: Optional[dict] = None


def _pose_row(ego_poses_df, ts) -> Optional[np.ndarray]:
    rows = ego_poses_df[ego_poses_df["timestamp_ns"] == ts]
    if rows.empty:
        return None
    r = rows.iloc[0]
    return np.array([r["tx_m"], r["ty_m"], r["tz_m"], r["qx"], r["qy"], r["qz"], r["qw"]], dtype=np.float64)


def load_frame(log_dir, sweep_ts_list: Sequence[int], current_ts_ns: int, ego_poses_df) -> Optional[SweepFrame]:
    """dataset.py:285-342 without the CPU transform: returns ``None`` where the reference returns ``None``
    (no current pose row :286-287, invalid current quaternion :294-296, every sweep missing :344-345)."""
    cur = _pose_row(ego_poses_df, current_ts_ns)
    if cur is None:
        return None
    try:
        _pose.se3_from_pose(cur[:3], cur[3:7])
    except ValueError:
        print(f"Warning: Invalid quaternion for ego pose at ts {current_ts_ns}. Skipping item.")
        return None
    base = Path(log_dir) / "sensors" / "lidar"
    S = len(sweep_ts_list)
    pts, counts = [], np.zeros(S, np.int64)
    poses = np.zeros((S, 7), np.float64)
    poses[:, 6] = 1.0                                            # identity for missing sweeps (never used: empty range)
    any_sweep = False
    for s, ts in enumerate(sweep_ts_list):
        sweep = decode_sweep(base / f"{ts}.feather")
        if sweep is None:
            continue
        prow = _pose_row(ego_poses_df, ts)                       # dataset.py:323-326
        if prow is None:
            continue
        try:
            _pose.se3_from_pose(prow[:3], prow[3:7])             # dataset.py:330-334: invalid quaternion -> None sweep
        except ValueError:
            continue
        poses[s] = prow
        pts.append(sweep)
        counts[s] = sweep.shape[0]
        any_sweep = True
    if not any_sweep:
        return None
    dt = np.float16 if all(p.dtype == np.float16 for p in pts) else np.float32
    off = np.zeros(S + 1, np.int32)
    off[1:] = np.cumsum(counts)
    return SweepFrame(np.concatenate([p.astype(dt, copy=False) for p in pts], 0), off, poses, cur)
