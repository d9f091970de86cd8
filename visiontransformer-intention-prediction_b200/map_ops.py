"""Host side of the HD-map rasterizer: device-resident map tables + the reference's signatures.

* ``MapTable``                  several pre-parsed maps (maps.FlatMap) concatenated on the device
* ``rasterize_maps``            batched: one 3x3 ego matrix per frame -> ``[B,9,400,720]`` CUDA tensor
* ``rasterize_map_ego_centric`` drop-in for utils.py:108-182 (same arguments, numpy float32 out,
                                same "print a warning and return zeros" error conventions)
* ``world_to_bev_pixel``        drop-in for utils.py:47-60
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native, maps, pose
from .bev_ops import _aug_fusion_args, _dev, _ptr, _require_cuda, _stream_ptr, to_numpy_fresh, workspace
from .constants import GRID_HEIGHT_PX, GRID_WIDTH_PX, MAP_CHANNELS


class MapTable:
    """Concatenation of FlatMaps on one device; ``frame_rows[m]`` = (lane_begin, lane_count, cw_begin, cw_count)."""

    def __init__(self, flat_maps: Sequence[maps.FlatMap], device: Optional[torch.device] = None) -> None:
        device = torch.device(device) if device is not None else _require_cuda()
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        verts, pl_off, lanes, cws, rows = [], [np.zeros(1, np.int64)], [], [], []
        v_base = p_base = l_base = c_base = 0
        max_len = 0
        for fm in flat_maps:
            verts.append(fm.verts)
            pl_off.append(fm.pl_off[1:].astype(np.int64) + v_base)
            ln = fm.lanes.copy()
            for col in (0, 1):
                ln[:, col] = np.where(ln[:, col] >= 0, ln[:, col] + p_base, -1)
            lanes.append(ln)
            cws.append(fm.cws + p_base)
            rows.append((l_base, ln.shape[0], c_base, fm.cws.shape[0]))
            v_base += fm.verts.shape[0]; p_base += fm.num_polylines
            l_base += ln.shape[0]; c_base += fm.cws.shape[0]
            max_len = max(max_len, fm.max_polyline_len())
        if max_len > maps.MAX_POLYLINE_VERTS:
            raise ValueError("polyline longer than the rasterizer bound")
        cat = lambda xs, dt, shape: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)).reshape(shape)
        self.device = device
        self.frame_rows = np.asarray(rows, np.int32).reshape(-1, 4)
        self.max_polyline_len = max(max_len, 1)
        self.verts = _dev(cat(verts, np.float64, (-1, 2)) if v_base else np.zeros((1, 2)), torch.float64, device)
        self.pl_off = _dev(np.concatenate(pl_off).astype(np.int32), torch.int32, device)
        self.lanes = _dev(cat(lanes, np.int32, (-1, 4)) if l_base else np.zeros((1, 4), np.int32), torch.int32, device)
        self.cws = _dev(cat(cws, np.int32, (-1,)) if c_base else np.zeros(1, np.int32), torch.int32, device)

    @property
    def num_maps(self) -> int:
        return self.frame_rows.shape[0]


def _frame_rows(table: MapTable, map_ids) -> np.ndarray:
    """(lane_begin, lane_count, cw_begin, cw_count) per frame.  -1 is FrameBatch's "no map" marker and must not reach
    numpy indexing (it would wrap to the last map): ids outside [0, num_maps) are an error."""
    ids = np.asarray(map_ids, np.int64).reshape(-1)
    if ids.size and (ids.min() < 0 or ids.max() >= table.num_maps):
        raise ValueError(f"map ids must be in [0, {table.num_maps}); got min {int(ids.min())}, max {int(ids.max())} "
                         "(-1 means 'no map': such frames cannot be rasterised)")
    return table.frame_rows[ids]


def rasterize_maps(table: MapTable, map_ids, ego_T, *, flip=None, drop_rects=None, drop_counts=None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """utils.py:108-182 for a batch: frame b uses map ``map_ids[b]`` of ``table`` and the 3x3 matrix
    ``ego_T[b]`` (pose.map_matrix_from_pose).  Returns float32 CUDA ``[B,9,400,720]``."""
    device = table.device
    rows = _frame_rows(table, map_ids)
    B = rows.shape[0]
    shape = (B, MAP_CHANNELS, GRID_HEIGHT_PX, GRID_WIDTH_PX)
    with torch.cuda.device(device):
        T = _dev(np.asarray(ego_T, np.float64).reshape(B, 9) if not isinstance(ego_T, torch.Tensor) else ego_T,
                 torch.float64, device).reshape(B, 9)
        fmap = _dev(rows, torch.int32, device)
        max_prims = int((rows[:, 1] + rows[:, 3]).max()) if B else 0
        f, r, c = _aug_fusion_args(B, flip, drop_rects, drop_counts, device)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != device:
            raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape} on {device}")
        L = _native.lib()
        wbytes = int(L.bevgen_map_workspace_bytes(B))
        ws = workspace.get("map", wbytes, device)
        rc = L.bevgen_map_raster(_ptr(table.verts), _ptr(table.pl_off), _ptr(table.lanes), _ptr(table.cws), _ptr(fmap),
                                 _ptr(T), B, max_prims, int(table.max_polyline_len), _ptr(f), _ptr(r), _ptr(c), _ptr(out),
                                 _ptr(ws), wbytes, _stream_ptr())
    _native.check(rc, "bevgen_map_raster")
    return out


def rasterize_maps_augmented(table: MapTable, map_ids, ego_T, params, *, out: Optional[torch.Tensor] = None,
                             packed: Optional[np.ndarray] = None) -> torch.Tensor:
    """utils.py:108-182 followed by the image part of augment_bev (utils.py:508-511) on the 9 map channels, for a batch:
    ``params`` = one ``augment.AugParams`` per frame.  Frames that rotate / scale are resampled straight from the
    rasterizer's bit planes (no intermediate float grid); the others get flip / dropout fused into the expand."""
    import ctypes as C
    from . import augment as _aug
    device = table.device
    rows = _frame_rows(table, map_ids)
    B = rows.shape[0]
    if len(params) != B:
        raise ValueError("one AugParams per frame")
    shape = (B, MAP_CHANNELS, GRID_HEIGHT_PX, GRID_WIDTH_PX)
    with torch.cuda.device(device):
        T = _dev(np.asarray(ego_T, np.float64).reshape(B, 9) if not isinstance(ego_T, torch.Tensor) else ego_T,
                 torch.float64, device).reshape(B, 9)
        fmap = _dev(rows, torch.int32, device)
        max_prims = int((rows[:, 1] + rows[:, 3]).max()) if B else 0
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != device:
            raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape} on {device}")
        hparams = _aug.pack_params(params) if packed is None else packed
        dparams = _dev(hparams, torch.uint8, device)
        L = _native.lib()
        wbytes = int(L.bevgen_map_workspace_bytes(B))
        ws = workspace.get("map", wbytes, device)
        rc = L.bevgen_map_raster_augmented(_ptr(table.verts), _ptr(table.pl_off), _ptr(table.lanes), _ptr(table.cws),
                                           _ptr(fmap), _ptr(T), B, max_prims, int(table.max_polyline_len), _ptr(dparams),
                                           C.c_void_p(hparams.ctypes.data), _ptr(out), _ptr(ws), wbytes, _stream_ptr())
    _native.check(rc, "bevgen_map_raster_augmented")
    return out


_TABLES = maps.LruDict(maps.MAX_CACHED_MAPS)     # (path, device) -> one-map device table


def table_for_path(map_json_path: str) -> MapTable:
    """One-map device table per (JSON path, current device), parsed and uploaded once (the reference re-parses per
    sample); least-recently-used tables beyond ``maps.MAX_CACHED_MAPS`` are dropped."""
    key = (str(map_json_path), torch.cuda.current_device() if torch.cuda.is_available() else -1)
    t = _TABLES.lookup(key)
    if t is None:
        t = MapTable([maps.load_map_json(key[0])])
        _TABLES.store(key, t)
    return t


def clear_cache() -> None:
    _TABLES.clear()
    maps.clear_cache()


def rasterize_map_ego_centric(map_json_path: str, current_ego_pose, *, as_tensor: bool = False):
    """Drop-in for utils.py:108-182.  ``current_ego_pose`` is the reference's pandas Series (or any mapping)
    with tx_m, ty_m, qx, qy, qz, qw.  Unreadable JSON or an invalid quaternion -> warning + all-zero map
    (utils.py:111-116, :123-127)."""
    shape = (MAP_CHANNELS, GRID_HEIGHT_PX, GRID_WIDTH_PX)

    def empty():
        if as_tensor:
            return torch.zeros(shape, dtype=torch.float32, device=_require_cuda())
        return np.zeros(shape, dtype=np.float32)

    try:
        table = table_for_path(map_json_path)
    except Exception as e:  # noqa: BLE001 -- the reference catches everything here (utils.py:114)
        print(f"Error loading map JSON {map_json_path}: {e}. Returning empty map.")
        return empty()
    tx, ty = current_ego_pose["tx_m"], current_ego_pose["ty_m"]
    q = np.array([current_ego_pose[k] for k in ("qx", "qy", "qz", "qw")])
    try:
        T3 = pose.map_matrix_from_pose(tx, ty, q)
    except ValueError:
        print(f"Warning: Invalid ego quaternion. Returning empty map for {map_json_path}")
        return empty()
    out = rasterize_maps(table, [0], T3[None])[0]
    return out if as_tensor else to_numpy_fresh(out)


def world_to_bev_pixel(points_world_xy: np.ndarray, ego_tf_matrix: np.ndarray) -> np.ndarray:
    """Drop-in for utils.py:47-60: (M,2) world xy -> (M,2) int pixel [px, py] (half-to-even rounding)."""
    pts = np.asarray(points_world_xy)
    if pts.shape[0] == 0:
        return np.empty((0, 2), dtype=int)
    device = _require_cuda()
    xy = _dev(np.ascontiguousarray(pts[:, :2], dtype=np.float64), torch.float64, device)
    T = np.ascontiguousarray(ego_tf_matrix, dtype=np.float64)
    out = torch.empty((pts.shape[0], 2), dtype=torch.int64, device=device)
    rc = _native.lib().bevgen_world_to_bev_pixel(_ptr(xy), pts.shape[0], C.c_void_p(T.ctypes.data), _ptr(out), _stream_ptr())
    _native.check(rc, "bevgen_world_to_bev_pixel")
    return out.cpu().numpy().astype(int)
