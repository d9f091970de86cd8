"""HD-map pre-parse: the reference's map JSON -> flat CSR arrays the rasterizer consumes.

The reference re-reads and walks the JSON for every sample (utils.py:111-119, :148-180).
Here a map is parsed ONCE into

* ``verts``   float64 [V,2]   world x,y of every usable vertex (entries that are dicts with
                              'x' and 'y'; anything else is skipped exactly as utils.py:134 does)
* ``pl_off``  int32 [P+1]     polyline CSR offsets into ``verts``
* ``lanes``   int32 [L,4]     (left polyline id | -1, right polyline id | -1, flags, 0)
                              flags: bit0 is_intersection (utils.py:156), bit1 lane_type == "BUS"
                              (utils.py:158), bits 4-7 / 8-11 = channel of the left / right lane-mark
                              type (6 DASHED_WHITE, 7 SOLID_WHITE, 8 SOLID_YELLOW, 0 none; utils.py:166-173)
* ``cws``     int32 [CW]      polyline ids of crosswalk 'polygon' lists (utils.py:175-180)

and cached per path, so per frame only the 3x3 ego matrix changes.
"""
from __future__ import annotations

import collections
import dataclasses
import json
from typing import List

import numpy as np

_MARK_TO_CHANNEL = {"DASHED_WHITE": 6, "SOLID_WHITE": 7, "SOLID_YELLOW": 8}
FLAG_INTERSECTION, FLAG_BUS = 1, 2
MAX_POLYLINE_VERTS = 1024     # device scratch bound per boundary (csrc/map_raster.cu)


@dataclasses.dataclass
class FlatMap:
    verts: np.ndarray
    pl_off: np.ndarray
    lanes: np.ndarray
    cws: np.ndarray

    @property
    def num_polylines(self) -> int:
        return len(self.pl_off) - 1

    @property
    def num_prims(self) -> int:
        """lanes + crosswalks = work items of the raster kernel (one warp each)."""
        return int(self.lanes.shape[0] + self.cws.shape[0])

    def max_polyline_len(self) -> int:
        return int(np.diff(self.pl_off).max()) if self.num_polylines else 0


class _Builder:
    def __init__(self) -> None:
        self.xy: List[tuple] = []
        self.off: List[int] = [0]

    def polyline(self, entries) -> int:
        if not entries or not isinstance(entries, (list, tuple)):
            return -1
        n0 = len(self.xy)
        for e in entries:
            if isinstance(e, dict) and "x" in e and "y" in e:
                self.xy.append((e["x"], e["y"]))
        if len(self.xy) == n0:
            return -1
        self.off.append(len(self.xy))
        return len(self.off) - 2


def parse_map_dict(map_data: dict) -> FlatMap:
    b = _Builder()
    lanes = []
    for lane in map_data.get("lane_segments", {}).values():
        left = b.polyline(lane.get("left_lane_boundary", []))
        right = b.polyline(lane.get("right_lane_boundary", []))
        flags = (FLAG_INTERSECTION if lane.get("is_intersection", False) else 0)
        flags |= (FLAG_BUS if lane.get("lane_type") == "BUS" else 0)
        flags |= _MARK_TO_CHANNEL.get(lane.get("left_lane_mark_type", ""), 0) << 4
        flags |= _MARK_TO_CHANNEL.get(lane.get("right_lane_mark_type", ""), 0) << 8
        lanes.append((left, right, flags, 0))
    cws = []
    for cw in map_data.get("pedestrian_crossings", {}).values():
        pid = b.polyline(cw.get("polygon", []))
        if pid >= 0:
            cws.append(pid)
    fm = FlatMap(
        verts=np.asarray(b.xy, dtype=np.float64).reshape(-1, 2),
        pl_off=np.asarray(b.off, dtype=np.int32),
        lanes=np.asarray(lanes, dtype=np.int32).reshape(-1, 4),
        cws=np.asarray(cws, dtype=np.int32),
    )
    if fm.max_polyline_len() > MAX_POLYLINE_VERTS:
        raise ValueError(f"map boundary with {fm.max_polyline_len()} vertices exceeds the "
                         f"rasterizer's bound of {MAX_POLYLINE_VERTS}")
    return fm


class LruDict(collections.OrderedDict):
    """Small entry-bounded LRU (one map per AV2 log: a few hundred KB each; a training run touches ~850 logs)."""

    def __init__(self, max_entries: int) -> None:
        super().__init__()
        self.max_entries = int(max_entries)

    def lookup(self, key):
        if key in self:
            self.move_to_end(key)
            return self[key]
        return None

    def store(self, key, value) -> None:
        self[key] = value
        self.move_to_end(key)
        while len(self) > self.max_entries:
            self.popitem(last=False)


MAX_CACHED_MAPS = 256
_CACHE = LruDict(MAX_CACHED_MAPS)


def load_map_json(path: str, use_cache: bool = True) -> FlatMap:
    """Parse (once) the map JSON at ``path``.  I/O and JSON errors propagate to the caller,
    which maps them to the reference's 'all-zero map + warning' convention (utils.py:111-116)."""
    key = str(path)
    if use_cache:
        hit = _CACHE.lookup(key)
        if hit is not None:
            return hit
    with open(key, "r") as fh:
        data = json.load(fh)
    fm = parse_map_dict(data)
    if use_cache:
        _CACHE.store(key, fm)
    return fm


def clear_cache() -> None:
    _CACHE.clear()
