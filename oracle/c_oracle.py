"""ctypes binding of oracle/bevgen_oracle.c (test infrastructure; see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libbevgen_oracle.so"

GRID_H, GRID_W, ZBINS, MAP_CH = 400, 720, 29, 9


def build(force: bool = False) -> Path:
    """Compile the C oracle with the recipe in oracle/Makefile (gcc only)."""
    src = _HERE / "bevgen_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        p = C.c_void_p
        i64, i32, f64 = C.c_int64, C.c_int, C.c_double
        L.orc_abi_version.restype = C.c_int
        L.orc_constants.argtypes = [p]
        L.orc_transform_points_f64.argtypes = [p, i64, i64, p, p]
        L.orc_transform_points_f32.argtypes = [p, i64, i64, p, p]
        L.orc_lidar_sweep_f64.argtypes = [p, p, i64, i32, p]
        L.orc_lidar_frame_fused.argtypes = [p, p, p, i32, p]
        L.orc_world_to_bev_pixel.argtypes = [p, i64, p, p]
        L.orc_polyline.argtypes = [p, p, i32]
        L.orc_fill_poly.argtypes = [p, p, i32]
        L.orc_map_raster.argtypes = [p, p, i32, p, i32, p, i32, p, p]
        L.orc_u8_to_f32.argtypes = [p, p, i64]
        L.orc_flip_w.argtypes = [p, p, i32]
        L.orc_rotation_matrix.argtypes = [f64, p]
        L.orc_invert_affine.argtypes = [p, p]
        L.orc_warp_affine.argtypes = [p, p, i32, p]
        L.orc_scale_dims.argtypes = [f64, p, p]
        L.orc_scale.argtypes = [p, p, i32, f64]
        L.orc_dropout.argtypes = [p, i32, p, i32]
        _lib = L
    return _lib


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def constants() -> dict:
    out = np.zeros(10, np.float64)
    lib().orc_constants(_ptr(out))
    keys = ["GRID_H", "GRID_W", "VOXEL", "OFF_X", "OFF_Y", "Z_MIN", "Z_MAX", "ZBINS", "MAP_CH", "Z_SPAN"]
    return dict(zip(keys, out.tolist()))


# --------------------------------------------------------------------------- transform
def transform_points(points: np.ndarray, T: np.ndarray) -> np.ndarray:
    """utils.py:27-33 -- returns float64 (N,3)."""
    n = points.shape[0]
    if n == 0:
        return np.empty((0, 3), dtype=points.dtype)
    T = _c(T, np.float64)
    out = np.empty((n, 3), np.float64)
    if points.dtype == np.float64:
        p = np.ascontiguousarray(points)
        lib().orc_transform_points_f64(_ptr(p), n, p.shape[1], _ptr(T), _ptr(out))
    else:
        p = _c(points, np.float32) if points.dtype != np.float32 else np.ascontiguousarray(points)
        lib().orc_transform_points_f32(_ptr(p), n, p.shape[1], _ptr(T), _ptr(out))
    return out


# --------------------------------------------------------------------------- lidar
def lidar_bev(points_list, intensity_list, num_expected_sweeps: int = 10) -> np.ndarray:
    """utils.py:62-106 on float64 ego-frame points (the Dataset path, dataset.py:340-347)."""
    bev = np.zeros((ZBINS * num_expected_sweeps, GRID_H, GRID_W), np.float32)
    loaded = min(len(points_list), len(intensity_list))
    for s in range(loaded):
        pts, inten = points_list[s], intensity_list[s]
        if pts is None or inten is None or pts.shape[0] == 0:
            continue
        if s >= num_expected_sweeps:
            raise IndexError("sweep index beyond num_expected_sweeps (reference: utils.py:102)")
        p = _c(pts[:, :3], np.float64)
        it = _c(inten, np.float32)
        lib().orc_lidar_sweep_f64(_ptr(p), _ptr(it), p.shape[0], s, _ptr(bev))
    return bev


def lidar_bev_fused(points4: np.ndarray, sweep_offsets: np.ndarray, rel_tf: np.ndarray,
                    num_expected_sweeps: int = 10) -> np.ndarray:
    """dataset.py:339-340 + utils.py:62-106 for one frame from raw float32 (x,y,z,intensity)."""
    S = len(sweep_offsets) - 1
    assert S <= num_expected_sweeps
    p = _c(points4, np.float32)
    off = _c(sweep_offsets, np.int32)
    tf = _c(rel_tf, np.float64).reshape(S, 16)
    bev = np.zeros((ZBINS * num_expected_sweeps, GRID_H, GRID_W), np.float32)
    lib().orc_lidar_frame_fused(_ptr(p), _ptr(off), _ptr(tf), S, _ptr(bev))
    return bev


# --------------------------------------------------------------------------- map
def world_to_bev_pixel(xy: np.ndarray, T3: np.ndarray) -> np.ndarray:
    """utils.py:47-60 -- returns int64 (M,2) [px,py]."""
    m = xy.shape[0]
    if m == 0:
        return np.empty((0, 2), dtype=int)
    a = _c(xy, np.float64)
    T = _c(T3, np.float64)
    out = np.empty((m, 2), np.int64)
    lib().orc_world_to_bev_pixel(_ptr(a), m, _ptr(T), _ptr(out))
    return out


def polyline(img: np.ndarray, pts: np.ndarray) -> None:
    v = _c(pts, np.int32).reshape(-1, 2)
    assert img.dtype == np.uint8 and img.shape == (GRID_H, GRID_W) and img.flags.c_contiguous
    lib().orc_polyline(_ptr(img), _ptr(v), v.shape[0])


def fill_poly(img: np.ndarray, pts: np.ndarray) -> None:
    v = _c(pts, np.int32).reshape(-1, 2)
    assert img.dtype == np.uint8 and img.shape == (GRID_H, GRID_W) and img.flags.c_contiguous
    lib().orc_fill_poly(_ptr(img), _ptr(v), v.shape[0])


MARK_CHANNEL = {"DASHED_WHITE": 6, "SOLID_WHITE": 7, "SOLID_YELLOW": 8}   # utils.py:166


def flatten_map(map_data: dict) -> dict:
    """Oracle-side flattening of the reference's map JSON dict (utils.py:118-119, :131-134,
    :148-180) into the arrays orc_map_raster takes.  Independent of the product's parser."""
    verts: list = []
    pl_off = [0]

    def add(lst) -> int:
        if not lst:                                   # utils.py:133
            return -1
        good = [(p["x"], p["y"]) for p in lst if isinstance(p, dict) and "x" in p and "y" in p]
        if not good:                                  # utils.py:135
            return -1
        verts.extend(good)
        pl_off.append(len(verts))
        return len(pl_off) - 2

    lanes = []
    for _, lane in map_data.get("lane_segments", {}).items():
        lp = add(lane.get("left_lane_boundary", []))
        rp = add(lane.get("right_lane_boundary", []))
        flags = 0
        if lane.get("is_intersection", False):
            flags |= 1
        if lane.get("lane_type") == "BUS":
            flags |= 2
        flags |= MARK_CHANNEL.get(lane.get("left_lane_mark_type", ""), 0) << 4
        flags |= MARK_CHANNEL.get(lane.get("right_lane_mark_type", ""), 0) << 8
        lanes.append((lp, rp, flags, 0))
    cws = []
    for _, cw in map_data.get("pedestrian_crossings", {}).items():
        poly = cw.get("polygon", [])
        if poly:
            pid = add(poly)
            if pid >= 0:
                cws.append(pid)
    return {
        "verts": np.asarray(verts, np.float64).reshape(-1, 2),
        "pl_off": np.asarray(pl_off, np.int32),
        "lanes": np.asarray(lanes, np.int32).reshape(-1, 4),
        "cws": np.asarray(cws, np.int32),
    }


def map_raster_flat(flat: dict, T3: np.ndarray) -> np.ndarray:
    """utils.py:146-182 given the flat map and the 3x3 world->ego matrix; float32 [9,400,720]."""
    out = np.zeros((MAP_CH, GRID_H, GRID_W), np.uint8)
    v, po, la, cw = (_c(flat["verts"], np.float64), _c(flat["pl_off"], np.int32),
                     _c(flat["lanes"], np.int32), _c(flat["cws"], np.int32))
    T = _c(T3, np.float64)
    lib().orc_map_raster(_ptr(v), _ptr(po), len(po) - 1, _ptr(la), la.shape[0], _ptr(cw), cw.shape[0],
                         _ptr(T), _ptr(out))
    return out.astype(np.float32)


# --------------------------------------------------------------------------- augmentation
def flip_w(grid: np.ndarray) -> np.ndarray:
    g = _c(grid, np.float32)
    out = np.empty_like(g)
    lib().orc_flip_w(_ptr(g), _ptr(out), g.shape[0])
    return out


def rotation_matrix(angle_deg: float) -> np.ndarray:
    M = np.zeros(6, np.float64)
    lib().orc_rotation_matrix(float(angle_deg), _ptr(M))
    return M.reshape(2, 3)


def invert_affine(M: np.ndarray) -> np.ndarray:
    Mi = np.zeros(6, np.float64)
    Mc = _c(M, np.float64).reshape(6)
    lib().orc_invert_affine(_ptr(Mc), _ptr(Mi))
    return Mi.reshape(2, 3)


def rotate(grid: np.ndarray, angle_deg: float) -> np.ndarray:
    """utils.py:425-439 for every channel of grid [C,400,720]."""
    g = _c(grid, np.float32)
    out = np.empty_like(g)
    Mi = invert_affine(rotation_matrix(angle_deg)).reshape(6)
    lib().orc_warp_affine(_ptr(g), _ptr(out), g.shape[0], _ptr(Mi))
    return out


def scale(grid: np.ndarray, s: float) -> np.ndarray:
    """utils.py:457-476 for every channel of grid [C,400,720]."""
    g = _c(grid, np.float32)
    out = np.empty_like(g)
    lib().orc_scale(_ptr(g), _ptr(out), g.shape[0], float(s))
    return out


def dropout(grid: np.ndarray, rects) -> np.ndarray:
    """utils.py:496-497 in place; rects = [(y0, x0, ph, pw), ...]."""
    assert grid.dtype == np.float32 and grid.flags.c_contiguous
    r = _c(np.asarray(rects, np.int32).reshape(-1, 4), np.int32)
    lib().orc_dropout(_ptr(grid), grid.shape[0], _ptr(r), r.shape[0])
    return grid
