"""Loader for the sm_100a shared library behind the C ABI of include/bevgen.h.

There is NO fallback: if ``csrc/libbevgen_b200.so`` is missing or fails to load, every op
raises.  ``build()`` is the in-tree nvcc recipe (also run by ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

from . import constants as K

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libbevgen_b200.so"
SOURCES = ["cabi.cu", "lidar.cu", "map_raster.cu", "augment.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                                   # no implicit FMA contraction: bit-exact binning / bilinear
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libbevgen_b200.so")
    return exe


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    deps = srcs + list(CSRC.glob("*.cuh")) + [CSRC.parent.parent / "include" / "bevgen.h"]
    if not force and LIB_PATH.exists() and all(LIB_PATH.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(LIB_PATH), *map(str, srcs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


class AugParams(C.Structure):
    """mirror of struct bevgen_aug_params (include/bevgen.h)"""
    _fields_ = [("minv", C.c_double * 6), ("scale", C.c_double),
                ("do_flip", C.c_int32), ("do_rotate", C.c_int32), ("do_scale", C.c_int32), ("n_rects", C.c_int32),
                ("rects", C.c_int32 * 20)]


_lib = None

# name -> (restype, argtypes); every symbol include/bevgen.h declares
_p, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
PROTOTYPES = {
    "bevgen_abi_version": (_i, []),
    "bevgen_constants": (None, [_p]),
    "bevgen_error_string": (C.c_char_p, [_i]),
    "bevgen_launch_count": (_i64, []),
    "bevgen_timing_enable": (None, [_i]),
    "bevgen_timing_collect": (_i, [C.c_char_p, _sz]),
    "bevgen_transform_points_f32": (_i, [_p, _i64, _i64, _p, _p, _p]),
    "bevgen_transform_points_f64": (_i, [_p, _i64, _i64, _p, _p, _p]),
    "bevgen_lidar_workspace_bytes": (_sz, [_i64, _i, _i]),
    "bevgen_lidar_overflow_total": (_i, [_p, _sz, _i64, _i, _i, _p, _p]),
    "bevgen_lidar_bev_fused": (_i, [_p, _p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _p, _sz, _i, _p]),
    "bevgen_lidar_bev_fused_f16": (_i, [_p, _p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _p, _sz, _i, _p]),
    "bevgen_lidar_bev_fused_cols": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _p, _sz, _i, _p]),
    "bevgen_lidar_bev_augmented_cols": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _sz, _p]),
    "bevgen_store_put_cols": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "bevgen_lidar_bev_store_cols": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _p, _sz, _i, _p]),
    "bevgen_lidar_bev_store_augmented_cols": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _sz, _p]),
    "bevgen_lidar_bev_ego_f64": (_i, [_p, _p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _p, _sz, _i, _p]),
    "bevgen_map_workspace_bytes": (_sz, [_i]),
    "bevgen_map_raster": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "bevgen_world_to_bev_pixel": (_i, [_p, _i64, _p, _p, _p]),
    "bevgen_augment": (_i, [_p, _p, _i, _i, _p, _p, _p, _sz, _p]),
    "bevgen_map_raster_augmented": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "bevgen_aug_fused_supported": (_i, [_p, _i]),
    "bevgen_lidar_aug_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "bevgen_lidar_bev_augmented": (_i, [_p, _i, _p, _p, _i, _i, _i64, _i, _p, _p, _p, _p, _sz, _p]),
}


def lib() -> C.CDLL:
    """The loaded library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(L, name)          # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    if L.bevgen_abi_version() != 2:
        raise RuntimeError("libbevgen_b200.so ABI version mismatch")
    got = (C.c_double * 10)()
    L.bevgen_constants(got)
    want = [K.GRID_HEIGHT_PX, K.GRID_WIDTH_PX, K.VOXEL_SIZE_M, K.BEV_PIXEL_OFFSET_X, K.BEV_PIXEL_OFFSET_Y,
            K.Z_MIN, K.Z_MAX, K.LIDAR_HEIGHT_CHANNELS, K.MAP_CHANNELS, K.Z_MAX - K.Z_MIN]
    if list(got) != [float(w) for w in want]:
        raise RuntimeError(f"grid constants differ between constants.py and the CUDA library: {list(got)} vs {want}")
    _lib = L
    return L


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().bevgen_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def launch_count() -> int:
    return int(lib().bevgen_launch_count())


def timing_enable(on: bool) -> None:
    lib().bevgen_timing_enable(1 if on else 0)


def timing_collect() -> dict:
    """{kernel name: (launches, total_ms)} since the last collect (synchronises the recorded events)."""
    buf = C.create_string_buffer(1 << 14)
    lib().bevgen_timing_collect(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out
