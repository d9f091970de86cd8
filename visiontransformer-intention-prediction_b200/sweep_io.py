"""Sweep decode: AV2 LiDAR feather files -> packed point rows for the device path.

Mirrors the host loop of the reference's ``Dataset.__getitem__`` (dataset.py:303-342) up to -- but not including -- the
hot path: which sweeps count as missing (file absent, empty table, corrupt file, no pose row, invalid quaternion ->
``None`` sweep, dataset.py:307-334), which columns are read (x, y, z, intensity; dataset.py:311), and the pose rows
looked up per sweep timestamp (dataset.py:323-331).  Instead of transforming on the CPU it returns the raw sweep-frame
points packed as ``[N,4]`` rows (x, y, z, intensity) in the narrowest exact dtype:

* float16 when every column is exactly representable (AV2 stores x,y,z as float16 and intensity as uint8) -> 8 B/point;
* float32 otherwise.

Decoded sweeps go through a byte-bounded LRU (default 256 MiB, ``set_cache_bytes``): with 10-sweep sliding windows
every file is otherwise parsed ten times per epoch, and a small window of recent sweeps captures that reuse.
"""
from __future__ import annotations

import collections
import dataclasses
import threading
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import pose as _pose

_COLUMNS = ["x", "y", "z", "intensity"]


class _SweepCache:
    """Byte-bounded LRU of decoded sweeps.  With 10-sweep sliding windows consecutive samples of a log share 9 of their
    10 files, so a small window of recent sweeps captures nearly all the reuse; an unbounded cache would grow to the whole
    split (~100 k files x ~0.8 MB) over an epoch.  The reference keeps no per-sweep state at all (dataset.py:305-321)."""

    DEFAULT_BYTES = 256 << 20          # ~300 sweeps of 100 k points at 8 B/point

    def __init__(self, max_bytes: int = DEFAULT_BYTES) -> None:
        self.max_bytes = int(max_bytes)
        self._d: "collections.OrderedDict[str, Optional[np.ndarray]]" = collections.OrderedDict()
        self._bytes = 0
        self._lock = threading.Lock()
        self.hits = self.misses = 0

    @staticmethod
    def _size(v: Optional[np.ndarray]) -> int:
        return int(getattr(v, "nbytes", 0)) + 64

    def get(self, key: str):
        with self._lock:
            if key in self._d:
                self._d.move_to_end(key)
                self.hits += 1
                return True, self._d[key]
            self.misses += 1
            return False, None

    def put(self, key: str, value: Optional[np.ndarray]) -> None:
        sz = self._size(value)
        with self._lock:
            if key in self._d:
                self._bytes -= self._size(self._d.pop(key))
            if sz > self.max_bytes:
                return                  # larger than the whole budget: not cached
            self._d[key] = value
            self._bytes += sz
            while self._bytes > self.max_bytes and self._d:
                _, old = self._d.popitem(last=False)
                self._bytes -= self._size(old)

    def resize(self, max_bytes: int) -> None:
        with self._lock:
            self.max_bytes = int(max_bytes)
            while self._bytes > self.max_bytes and self._d:
                _, old = self._d.popitem(last=False)
                self._bytes -= self._size(old)

    def clear(self) -> None:
        with self._lock:
            self._d.clear()
            self._bytes = 0

    @property
    def nbytes(self) -> int:
        return self._bytes

    def __len__(self) -> int:
        return len(self._d)


_CACHE = _SweepCache()


def set_cache_bytes(max_bytes: int) -> None:
    """Resize the decoded-sweep LRU (0 disables caching)."""
    _CACHE.resize(max_bytes)


def decode_sweep(path, use_cache: bool = True) -> Optional[np.ndarray]:
    """One sweep file -> ``[N,4]`` float16/float32 rows, or ``None`` for the reference's "skip this sweep" cases
    (missing file dataset.py:307-309, empty table :312-314, unreadable file :315-318)."""
    import pyarrow as pa
    import pyarrow.feather as feather
    key = str(path)
    if use_cache:
        hit, val = _CACHE.get(key)
        if hit:
            return val
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
        _CACHE.put(key, out)
    return out


def decode_sweep_columns(path, use_cache: bool = True):
    """One sweep file -> its (x, y, z, intensity) COLUMNS exactly as stored (AV2: float16, float16, float16, uint8), or
    ``None`` for the reference's "skip this sweep" cases -- no row interleaving, no dtype conversion: the arrays are the
    Arrow column buffers.  A readable file whose columns are not (float16 x 3, uint8) yields ``NOT_COLUMNAR``: use
    ``decode_sweep`` for it (``load_frame`` falls back by itself)."""
    import pyarrow as pa
    import pyarrow.feather as feather
    key = "cols:" + str(path)
    if use_cache:
        hit, val = _CACHE.get(key)
        if hit:
            return val
    out = None
    p = Path(str(path))
    if p.is_file():
        try:
            table = feather.read_table(str(path), columns=_COLUMNS)
            if table.num_rows > 0:
                cols = tuple(table.column(c).to_numpy(zero_copy_only=False) for c in _COLUMNS)
                columnar = all(c.dtype == np.float16 for c in cols[:3]) and cols[3].dtype == np.uint8
                out = _Columns(cols) if columnar else NOT_COLUMNAR
        except (pa.ArrowInvalid, OSError):
            print(f"Warning: Corrupt LiDAR sweep file: {path}. Skipping sweep.")
            out = None
    if use_cache:
        _CACHE.put(key, out)
    return out


NOT_COLUMNAR = "not-columnar"       # decode_sweep_columns: readable sweep, but not AV2's (float16 x 3, uint8) dtypes


class _Columns(tuple):
    """(x, y, z, intensity) arrays of one sweep; ``nbytes`` / ``shape`` so that the LRU can size it like an array."""

    @property
    def nbytes(self) -> int:
        return int(sum(c.nbytes for c in self))

    @property
    def shape(self):
        return (self[0].shape[0], 4)


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
    columns: Optional[tuple] = None      # layout="columns": (x f16, y f16, z f16, intensity u8) of all sweeps; points4 is None


def _pose_row(ego_poses_df, ts) -> Optional[np.ndarray]:
    rows = ego_poses_df[ego_poses_df["timestamp_ns"] == ts]
    if rows.empty:
        return None
    r = rows.iloc[0]
    return np.array([r["tx_m"], r["ty_m"], r["tz_m"], r["qx"], r["qy"], r["qz"], r["qw"]], dtype=np.float64)


def load_frame(log_dir, sweep_ts_list: Sequence[int], current_ts_ns: int, ego_poses_df, *,
               use_cache: bool = True, layout: str = "rows") -> Optional[SweepFrame]:
    """dataset.py:285-342 without the CPU transform: returns ``None`` where the reference returns ``None``
    (no current pose row :286-287, invalid current quaternion :294-296, every sweep missing :344-345).
    ``use_cache=False`` decodes every file afresh (the reference's behaviour); the default goes through the
    byte-bounded LRU (``set_cache_bytes``).  ``layout="columns"`` keeps AV2's stored columns (float16 x,y,z + uint8
    intensity, 7 B/point; ``SweepFrame.columns``) instead of interleaving them into ``[N,4]`` rows -- the frame falls back
    to rows when a file holds other dtypes."""
    if layout not in ("rows", "columns"):
        raise ValueError("layout must be 'rows' or 'columns'")
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
        sweep = None
        if layout == "columns":
            sweep = decode_sweep_columns(base / f"{ts}.feather", use_cache=use_cache)
            if sweep is NOT_COLUMNAR:                                  # other dtypes on disk: this frame goes by rows
                layout, sweep = "rows", decode_sweep(base / f"{ts}.feather", use_cache=use_cache)
        else:
            sweep = decode_sweep(base / f"{ts}.feather", use_cache=use_cache)
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
    off = np.zeros(S + 1, np.int32)
    off[1:] = np.cumsum(counts)
    if layout == "columns":
        cols = tuple(np.concatenate([p[k] for p in pts]) for k in range(4))
        return SweepFrame(None, off, poses, cur, columns=cols)
    pts = [np.stack([c.astype(np.float32) for c in p], 1) if isinstance(p, _Columns) else p for p in pts]
    dt = np.float16 if all(p.dtype == np.float16 for p in pts) else np.float32
    return SweepFrame(np.concatenate([p.astype(dt, copy=False) for p in pts], 0), off, poses, cur)
