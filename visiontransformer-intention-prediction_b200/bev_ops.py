"""Host-side mirror of the reference's BEV-generation functions on top of the C ABI.

Two layers, both thin (PyTorch is used for device memory and streams only):

* the reference's Dataset-facing names with identical signatures and return types
  (numpy in -> numpy out), so ``dataset.py`` can rebind them unchanged:
    ``transform_points``              utils.py:27-33
    ``create_intentnet_lidar_bev``    utils.py:62-106
    ``rasterize_map_ego_centric``     utils.py:108-182      (map_ops.py)
    ``augment_bev`` / ``random_*``    utils.py:394-517      (augment.py)
* the batched, device-resident fast path (``lidar_bev_from_sweeps`` ...) that replaces
  transform + scatter for a whole batch of frames in three kernel launches and writes the
  ``[B, 290, 400, 720]`` tensor the models consume (dataset.py:146-147, train_vit.py:154-159).

No CPU fallback: every function here ends in a call into ``libbevgen_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native
from .constants import (GRID_HEIGHT_PX, GRID_WIDTH_PX, LIDAR_HEIGHT_CHANNELS, LIDAR_SWEEPS)

# LiDAR algorithms (bit-identical; bench.py / DESIGN.md section 3.1 record the measured ranking):
#   TILED (default)   one binning pass into per-tile record bins -> per 28.8 KB tile: shared-memory tile,
#                     __match_any_sync-aggregated atomicMax, ONE TMA bulk store (each output byte written once)
#   ATOMIC            TMA zero-fill kernel + one global RED.MAX scatter kernel (faster on dense / skewed scenes)
#   MEMSET_ATOMIC     cudaMemsetAsync + the same scatter kernel (library fill, comparison point only)
#   AUTO              (pipeline.BevGenerator only) TILED, switching to ATOMIC while the scenes keep overflowing the tile bins
ALGO_TILED, ALGO_ATOMIC, ALGO_MEMSET_ATOMIC, ALGO_AUTO = 0, 1, 2, 3
MAX_DROP_RECTS = 5


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("bevgen_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _device_of(*objs) -> torch.device:
    """Device the call runs on: that of the first CUDA tensor among ``objs`` (a BevGenerator on cuda:1 keeps working
    when cuda:0 is the current device), else the current device.  Callers launch under ``torch.cuda.device(dev)``."""
    for o in objs:
        if isinstance(o, torch.Tensor) and o.is_cuda:
            return o.device
    return _require_cuda()


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_offsets(host_off: np.ndarray, n_points: int, max_sweep_points: Optional[int]) -> int:
    """CSR offsets that are available on the host are validated here (device-resident ones are clamped inside the
    kernels instead): monotone, inside the point array; returns the largest sweep.  ``max_sweep_points`` only sizes the
    launch grid -- a value below the largest sweep is corrected rather than silently dropping points."""
    off = np.asarray(host_off, np.int64)
    if off.size == 0:
        return 0
    d = np.diff(off)
    if off[0] < 0 or (d < 0).any() or off[-1] > n_points:
        raise ValueError(f"sweep_offsets must be non-decreasing and within [0, {n_points}]")
    biggest = int(d.max()) if d.size else 0
    return biggest if max_sweep_points is None else max(int(max_sweep_points), biggest)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _dev(a, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """numpy / tensor -> contiguous device tensor of `dtype` (no copy when already right)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device, non_blocking=True)


class PointColumns:
    """A batch of sweeps in AV2's on-disk column form (dataset.py:311 reads exactly these columns): ``x``, ``y``, ``z``
    float16 and ``intensity`` uint8, sweeps concatenated inside each column -- 7 bytes per point instead of the 16 of a
    float32 ``[N,4]`` row array, and no interleaving work on the host (Arrow column buffers are copied as they are).

    Held as ONE uint8 buffer ``[7 * stride]`` = x | y | z | intensity (``stride`` = rows per column, a multiple of 8 so
    every column starts 16-byte aligned and whole 8-row vectors are readable), so a batch is one host->device copy."""

    def __init__(self, buf: torch.Tensor, n: int, stride: int) -> None:
        if buf.dtype != torch.uint8 or buf.dim() != 1 or stride % 8 or buf.numel() != 7 * stride or not 0 <= n <= stride:
            raise ValueError("PointColumns: uint8 buffer of 7 * stride bytes, stride a multiple of 8, n <= stride")
        self.buf, self.n, self.stride = buf, int(n), int(stride)

    @staticmethod
    def stride_for(n: int) -> int:
        return max(8, (int(n) + 7) // 8 * 8)

    @classmethod
    def from_arrays(cls, x, y, z, intensity, pinned: bool = True) -> "PointColumns":
        """Pack host columns (lossless: raises when a value is not representable in float16 / uint8)."""
        cols = [np.asarray(c).reshape(-1) for c in (x, y, z, intensity)]
        n = cols[0].shape[0]
        if any(c.shape[0] != n for c in cols):
            raise ValueError("x, y, z, intensity must have the same length")
        stride = cls.stride_for(n)
        # written in place: one pass over the data (pinned blocks come from torch's caching host allocator, so repeated
        # packs of similar size reuse them); only the padding rows need clearing
        buf = torch.empty(7 * stride, dtype=torch.uint8, pin_memory=bool(pinned and torch.cuda.is_available()))
        host = buf.numpy()
        if stride > n:
            for k in range(3):
                host[2 * (k * stride + n):2 * (k + 1) * stride] = 0
            host[6 * stride + n:7 * stride] = 0
        for k, c in enumerate(cols[:3]):
            h = c if c.dtype == np.float16 else c.astype(np.float16)
            if h is not c and not np.array_equal(h.astype(np.float64), c.astype(np.float64), equal_nan=True):
                raise ValueError("coordinates are not exactly representable in float16; use the float32 row format")
            host[2 * k * stride:2 * k * stride + 2 * n].view(np.float16)[:] = h
        i8 = cols[3] if cols[3].dtype == np.uint8 else cols[3].astype(np.uint8)
        if i8 is not cols[3] and not np.array_equal(i8.astype(np.float64), cols[3].astype(np.float64)):
            raise ValueError("intensities are not integers in [0, 255]; use the float32 row format")
        host[6 * stride:6 * stride + n] = i8
        return cls(buf, n, stride)

    @classmethod
    def from_points4(cls, points4: np.ndarray, pinned: bool = True) -> "PointColumns":
        p = np.asarray(points4).reshape(-1, 4)
        return cls.from_arrays(p[:, 0], p[:, 1], p[:, 2], p[:, 3], pinned=pinned)

    # -- tensor-like surface used by the pipeline
    @property
    def device(self) -> torch.device:
        return self.buf.device

    @property
    def is_cuda(self) -> bool:
        return self.buf.is_cuda

    @property
    def nbytes(self) -> int:
        return int(self.buf.numel())

    def to(self, device, non_blocking: bool = False) -> "PointColumns":
        return PointColumns(self.buf.to(device, non_blocking=non_blocking), self.n, self.stride)

    def record_stream(self, stream) -> None:
        self.buf.record_stream(stream)

    def ptrs(self):
        b = self.buf.data_ptr()
        return tuple(C.c_void_p(b + k * self.stride) for k in (0, 2, 4, 6))

    def to_points4(self) -> np.ndarray:
        """float32 [n,4] rows (host copy; for tests and the CPU oracle)."""
        host = self.buf.cpu().numpy()
        s, n = self.stride, self.n
        xyz = [host[2 * k * s:2 * k * s + 2 * n].view(np.float16).astype(np.float32) for k in range(3)]
        return np.stack(xyz + [host[6 * s:6 * s + n].astype(np.float32)], axis=1)


class _Workspace:
    """Grow-only scratch buffer per (purpose, device, CUDA stream); kernels never allocate.

    The stream is part of the key: launches on different streams (two generators, a generator's LiDAR and map streams,
    host threads with their own current streams) may overlap on the GPU and must not share counters / bins / bit planes.
    A buffer is allocated under, and only ever used on, the stream it is keyed by, so torch's caching allocator needs no
    ``record_stream``.  Launches on ONE stream are ordered and can share."""

    def __init__(self) -> None:
        self._bufs = {}
        self._lock = threading.Lock()

    def get(self, key, nbytes: int, device: torch.device) -> torch.Tensor:
        k = (key, device, torch.cuda.current_stream(device).cuda_stream)
        with self._lock:
            buf = self._bufs.get(k)
            if buf is None or buf.numel() < nbytes:
                with torch.cuda.device(device):
                    buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
                self._bufs[k] = buf
            return buf

    def clear(self) -> None:
        with self._lock:
            self._bufs.clear()


workspace = _Workspace()

class _PinnedPool:
    """Pinned host buffers handed out as the backing store of the numpy arrays the reference-signature functions return.

    The reference returns a FRESH array per call (utils.py:70, :146, :182).  Allocating 334 MB of pageable memory per call
    costs more than the whole GPU path (page faults + a pageable device->host copy), and so does a pinned staging
    buffer plus a host memcpy.  Instead the array returned IS a pinned buffer: the device->host copy runs at full PCIe rate
    straight into it, and the buffer goes back to the pool only when the array (and every view / ``torch.from_numpy`` alias
    of it) has been garbage-collected -- so no two live results ever share memory, which is all "fresh" means to a caller.
    Small results (< 8 MiB, e.g. transform_points' output) take the plain path."""

    MIN_BYTES = 8 << 20
    MAX_IDLE_BYTES = 8 << 30          # idle buffers kept for reuse; anything above is released to the OS

    def __init__(self) -> None:
        self._free = {}               # nbytes -> [uint8 pinned tensors]
        self._idle_bytes = 0

    def _release(self, buf: torch.Tensor) -> None:
        n = buf.numel()
        if self._idle_bytes + n > self.MAX_IDLE_BYTES:
            return                    # dropped: freed by torch
        self._free.setdefault(n, []).append(buf)
        self._idle_bytes += n

    def take(self, nbytes: int) -> torch.Tensor:
        lst = self._free.get(nbytes)
        if lst:
            self._idle_bytes -= nbytes
            return lst.pop()
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory()

    def to_numpy(self, t: torch.Tensor) -> np.ndarray:
        import weakref
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        buf = self.take(nbytes)
        host = buf.view(t.dtype).reshape(t.shape)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()        # the tensor's device, which need not be the current one
        arr = host.numpy()
        weakref.finalize(arr, self._release, buf)     # numpy views keep `arr` alive through .base, so this fires last
        return arr

    def clear(self) -> None:
        self._free.clear()
        self._idle_bytes = 0


pinned_pool = _PinnedPool()


def to_numpy_fresh(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> NEW numpy array (what the reference's functions return); large results live in pooled pinned
    memory (see _PinnedPool), small ones are plain copies."""
    if t.numel() * t.element_size() < _PinnedPool.MIN_BYTES:
        return t.cpu().numpy()
    return pinned_pool.to_numpy(t)


def _aug_fusion_args(B, flip, drop_rects, drop_counts, device):
    f = None if flip is None else _dev(flip, torch.uint8, device).reshape(B)
    if (drop_rects is None) != (drop_counts is None):
        raise ValueError("drop_rects and drop_counts go together")
    r = c = None
    if drop_rects is not None:
        r = _dev(drop_rects, torch.int32, device).reshape(B, MAX_DROP_RECTS, 4)
        c = _dev(drop_counts, torch.int32, device).reshape(B)
    return f, r, c


# =========================================================================== batched fast path
def lidar_bev_from_sweeps(points4, sweep_offsets, rel_tf, num_frames: int, num_sweeps: int = LIDAR_SWEEPS, *,
                          max_sweep_points: Optional[int] = None, flip=None, drop_rects=None, drop_counts=None,
                          out: Optional[torch.Tensor] = None, algo: int = ALGO_TILED, sweep_start=None,
                          total_points: Optional[int] = None) -> torch.Tensor:
    """Fused dataset.py:339-340 + utils.py:62-106 for ``num_frames`` frames.

    points4        float32 (16 B/pt) or float16 (8 B/pt) [N,4]  (x, y, z, intensity) in each sweep's own frame, sweeps
                   concatenated -- or a ``PointColumns`` (AV2's on-disk columns: float16 x,y,z + uint8 intensity, 7 B/pt)
    sweep_offsets  int32 [B*S+1]  CSR offsets; an empty range = missing sweep (its 29 channels stay 0)
    rel_tf         float64 [B*S,4,4]  ego_SE3_world @ sweep_SE3 (pose.sweep_rel_transforms)
    sweep_start    sweep STORE form (PointColumns only): int32 [B*S] first row of every sweep in a resident store; the sweeps'
                   lengths are then ``diff(sweep_offsets)`` and ``total_points`` their sum (sweep_store.SweepStore)
    returns        float32 CUDA tensor [B, S*29, 400, 720]
    """
    cols = points4 if isinstance(points4, PointColumns) else None
    if sweep_start is not None and cols is None:
        raise ValueError("sweep_start (sweep store form) needs the points as PointColumns")
    device = _device_of(cols.buf if cols is not None else points4, sweep_offsets, rel_tf, out)
    B, S = int(num_frames), int(num_sweeps)
    host_off = None
    if not isinstance(sweep_offsets, torch.Tensor) or not sweep_offsets.is_cuda:
        host_off = np.ascontiguousarray(sweep_offsets.numpy() if isinstance(sweep_offsets, torch.Tensor) else sweep_offsets,
                                        dtype=np.int32)
    half = (isinstance(points4, torch.Tensor) and points4.dtype == torch.float16) or \
           (isinstance(points4, np.ndarray) and points4.dtype == np.float16)
    with torch.cuda.device(device):
        off = _dev(sweep_offsets, torch.int32, device)
        tf = _dev(rel_tf, torch.float64, device).reshape(-1, 16)
        if cols is not None:
            pts = cols if cols.is_cuda else cols.to(device, non_blocking=True)
            n = pts.n
        else:
            pts = _dev(points4, torch.float16 if half else torch.float32, device)
            if pts.dim() != 2 or pts.shape[1] != 4:
                raise ValueError("points4 must be [N,4] float32 or float16 (x,y,z,intensity), or a PointColumns")
            n = int(pts.shape[0])
        if off.numel() != B * S + 1 or tf.shape[0] != B * S:
            raise ValueError("sweep_offsets must have B*S+1 entries and rel_tf B*S matrices")
        store_rows = n
        start = None
        if sweep_start is not None:
            start = _dev(sweep_start, torch.int32, device)
            if start.numel() != B * S or total_points is None:
                raise ValueError("sweep_start must have B*S entries and total_points (sum of the sweep lengths) must be given")
            n = int(total_points)              # the offsets are a prefix sum of lengths here: they add up to n, not to the store
        if host_off is not None:
            max_sweep_points = _check_offsets(host_off, n, max_sweep_points)
        elif max_sweep_points is None:
            max_sweep_points = _check_offsets(off.cpu().numpy(), n, None)
        f, r, c = _aug_fusion_args(B, flip, drop_rects, drop_counts, device)
        shape = (B, S * LIDAR_HEIGHT_CHANNELS, GRID_HEIGHT_PX, GRID_WIDTH_PX)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != device:
            raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape} on {device}")
        L = _native.lib()
        ws = None
        wbytes = 0
        if algo == ALGO_TILED:
            wbytes = int(L.bevgen_lidar_workspace_bytes(n, B, S))
            ws = workspace.get("lidar", wbytes, device)
        if start is not None:
            rc = L.bevgen_lidar_bev_store_cols(*pts.ptrs(), store_rows, _ptr(start), _ptr(off), _ptr(tf), B, S, n,
                                               int(max_sweep_points), _ptr(f), _ptr(r), _ptr(c), _ptr(out), _ptr(ws), wbytes,
                                               int(algo), _stream_ptr())
        elif cols is not None:
            rc = L.bevgen_lidar_bev_fused_cols(*pts.ptrs(), _ptr(off), _ptr(tf), B, S, n, int(max_sweep_points),
                                               _ptr(f), _ptr(r), _ptr(c), _ptr(out), _ptr(ws), wbytes, int(algo), _stream_ptr())
        else:
            fn = L.bevgen_lidar_bev_fused_f16 if half else L.bevgen_lidar_bev_fused
            rc = fn(_ptr(pts), _ptr(off), _ptr(tf), B, S, n, int(max_sweep_points),
                    _ptr(f), _ptr(r), _ptr(c), _ptr(out), _ptr(ws), wbytes, int(algo), _stream_ptr())
    _native.check(rc, "bevgen_lidar_bev_fused")
    return out


def overflow_total_async(total_points: int, num_frames: int, num_sweeps: int, device: torch.device):
    """After a TILED ``lidar_bev_from_sweeps`` call on the current stream: (pinned int32 tensor, event) -- once the event has
    completed the tensor holds the number of records that call sent to its overflow lists."""
    L = _native.lib()
    with torch.cuda.device(device):
        wbytes = int(L.bevgen_lidar_workspace_bytes(int(total_points), int(num_frames), int(num_sweeps)))
        ws = workspace.get("lidar", wbytes, device)
        word = torch.empty(1, dtype=torch.int32, device=device)
        rc = L.bevgen_lidar_overflow_total(_ptr(ws), wbytes, int(total_points), int(num_frames), int(num_sweeps), _ptr(word), _stream_ptr())
        _native.check(rc, "bevgen_lidar_overflow_total")
        host = torch.empty(1, dtype=torch.int32, pin_memory=True)
        host.copy_(word, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
    return host, ev


def lidar_bev_from_ego_points(points3, intensity, sweep_offsets, num_frames: int, num_sweeps: int = LIDAR_SWEEPS, *,
                              max_sweep_points: Optional[int] = None, flip=None, drop_rects=None, drop_counts=None,
                              out: Optional[torch.Tensor] = None, algo: int = ALGO_TILED) -> torch.Tensor:
    """utils.py:62-106 on float64 points that are already in the ego frame (batched)."""
    device = _device_of(points3, intensity, sweep_offsets, out)
    B, S = int(num_frames), int(num_sweeps)
    host_off = None
    if not isinstance(sweep_offsets, torch.Tensor) or not sweep_offsets.is_cuda:
        host_off = np.ascontiguousarray(sweep_offsets.numpy() if isinstance(sweep_offsets, torch.Tensor) else sweep_offsets,
                                        dtype=np.int32)
    with torch.cuda.device(device):
        pts = _dev(points3, torch.float64, device)
        inten = _dev(intensity, torch.float32, device)
        off = _dev(sweep_offsets, torch.int32, device)
        if pts.dim() != 2 or pts.shape[1] != 3 or inten.numel() != pts.shape[0] or off.numel() != B * S + 1:
            raise ValueError("points3 [N,3] float64, intensity [N] float32, sweep_offsets [B*S+1] expected")
        n = int(pts.shape[0])
        if host_off is not None:
            max_sweep_points = _check_offsets(host_off, n, max_sweep_points)
        elif max_sweep_points is None:
            max_sweep_points = _check_offsets(off.cpu().numpy(), n, None)
        f, r, c = _aug_fusion_args(B, flip, drop_rects, drop_counts, device)
        shape = (B, S * LIDAR_HEIGHT_CHANNELS, GRID_HEIGHT_PX, GRID_WIDTH_PX)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != device:
            raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape} on {device}")
        L = _native.lib()
        ws, wbytes = None, 0
        if algo == ALGO_TILED:
            wbytes = int(L.bevgen_lidar_workspace_bytes(n, B, S))
            ws = workspace.get("lidar", wbytes, device)
        rc = L.bevgen_lidar_bev_ego_f64(_ptr(pts), _ptr(inten), _ptr(off), B, S, n, int(max_sweep_points),
                                        _ptr(f), _ptr(r), _ptr(c), _ptr(out), _ptr(ws), wbytes, int(algo), _stream_ptr())
    _native.check(rc, "bevgen_lidar_bev_ego_f64")
    return out


# =========================================================================== reference signatures
def transform_points(points: np.ndarray, transform_matrix: np.ndarray) -> np.ndarray:
    """Drop-in for utils.py:27-33: float64 (N,3) = (T @ [x y z 1]^T)^T[:, :3] computed on the GPU.

    Empty input returns ``np.empty((0, 3), dtype=points.dtype)`` exactly as utils.py:29-30."""
    points = np.asarray(points)
    if points.shape[0] == 0:
        return np.empty((0, 3), dtype=points.dtype)
    if points.ndim != 2 or points.shape[1] < 3:
        raise ValueError("points must be (N, >=3)")
    device = _require_cuda()
    T = np.ascontiguousarray(transform_matrix, dtype=np.float64)
    if T.shape != (4, 4):
        raise ValueError("transform_matrix must be 4x4")
    n = points.shape[0]
    out = torch.empty((n, 3), dtype=torch.float64, device=device)
    L = _native.lib()
    if points.dtype == np.float64:
        p = _dev(points[:, :3], torch.float64, device)
        rc = L.bevgen_transform_points_f64(_ptr(p), n, 3, C.c_void_p(T.ctypes.data), _ptr(out), _stream_ptr())
    else:   # float16 / float32 / ints widen exactly to float64 through float32 only if they fit; be exact:
        if points.dtype in (np.float16, np.float32):
            p = _dev(points[:, :3].astype(np.float32, copy=False), torch.float32, device)
            rc = L.bevgen_transform_points_f32(_ptr(p), n, 3, C.c_void_p(T.ctypes.data), _ptr(out), _stream_ptr())
        else:
            p = _dev(points[:, :3].astype(np.float64), torch.float64, device)
            rc = L.bevgen_transform_points_f64(_ptr(p), n, 3, C.c_void_p(T.ctypes.data), _ptr(out), _stream_ptr())
    _native.check(rc, "bevgen_transform_points")
    return to_numpy_fresh(out)


class DeferredPoints:
    """What ``transform_points_deferred`` returns: a sweep's raw points and its matrix, standing in for the float64 ``(N,3)``
    array ``transform_points`` would have returned.  ``create_intentnet_lidar_bev`` recognises it and runs transform +
    binning as ONE fused pass over the raw points (what ``bevgen_lidar_bev_fused`` is for) instead of ten device round
    trips of float64 arrays; anything else that touches it (``np.asarray``, indexing, arithmetic through ``__array__``)
    gets the transformed array, computed on first use -- same bits either way."""

    ndim = 2
    dtype = np.dtype(np.float64)

    def __init__(self, raw: np.ndarray, T: np.ndarray) -> None:
        self.raw, self.matrix, self._value = raw, T, None        # raw: float16 / float32 [N,3], own copy; matrix: float64 [4,4], own copy

    @property
    def shape(self):
        return (self.raw.shape[0], 3)

    def __len__(self) -> int:
        return self.raw.shape[0]

    def materialize(self) -> np.ndarray:
        if self._value is None:
            self._value = transform_points(self.raw, self.matrix)
        return self._value

    def __array__(self, dtype=None, copy=None):
        v = self.materialize()
        return v if dtype is None else v.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __iter__(self):
        return iter(self.materialize())

    def __getattr__(self, name):                 # anything else an ndarray has (.T, .mean(), .astype, ...): the real array's
        if name.startswith("__") or name in ("raw", "matrix", "_value"):      # (own fields not set yet: copy / pickle protocols)
            raise AttributeError(name)
        return getattr(self.materialize(), name)


def _forward_operator(name):
    def op(self, *args):
        return getattr(self.materialize(), name)(*args)
    op.__name__ = name
    return op


for _name in ("add", "sub", "mul", "truediv", "floordiv", "mod", "pow", "matmul", "lt", "le", "gt", "ge", "eq", "ne"):
    for _dunder in (f"__{_name}__", f"__r{_name}__"):
        if hasattr(np.ndarray, _dunder):
            setattr(DeferredPoints, _dunder, _forward_operator(_dunder))
for _dunder in ("__neg__", "__abs__"):
    setattr(DeferredPoints, _dunder, _forward_operator(_dunder))
DeferredPoints.__hash__ = None                   # like ndarray: == is element-wise


def transform_points_deferred(points: np.ndarray, transform_matrix: np.ndarray):
    """``transform_points`` for the one place the reference calls it (dataset.py:340, whose results only feed
    ``create_intentnet_lidar_bev`` at :347): float16 / float32 points -- what AV2's feather files hold, dataset.py:320 --
    come back as a ``DeferredPoints``; every other input is transformed right away, exactly as ``transform_points``."""
    points = np.asarray(points)
    if points.ndim != 2 or points.shape[0] == 0 or points.shape[1] < 3 or points.dtype not in (np.float16, np.float32):
        return transform_points(points, transform_matrix)
    T = np.array(transform_matrix, dtype=np.float64)
    if T.shape != (4, 4):
        raise ValueError("transform_matrix must be 4x4")
    raw = np.array(points[:, :3], order="K")                  # own copy, dtype and memory order kept (pandas' .values is column-major)
    if not (raw.flags.c_contiguous or raw.flags.f_contiguous):
        raw = np.ascontiguousarray(raw)
    return DeferredPoints(raw, T)


def _fused_from_deferred(points_list: Sequence, intensity_list: Sequence, num_expected_sweeps: int, algo: int) -> torch.Tensor:
    """create_intentnet_lidar_bev when every loaded sweep is a DeferredPoints: the sweeps' raw coordinate and intensity
    arrays are copied AS THEY ARE (contiguous memcpys) into one pinned staging buffer, uploaded once, laid out as (x, y, z,
    intensity) float32 rows on the device (float16 -> float32 is exact), and transformed + binned by the fused launch.
    Skip rules and errors as ``_pack_sweeps``."""
    S = num_expected_sweeps
    device = _require_cuda()
    loaded = min(len(points_list), len(intensity_list))
    counts = np.zeros(S, np.int64)
    rel = np.tile(np.eye(4), (S, 1, 1))
    parts, used = [], 0
    for s in range(loaded):
        pts, inten = points_list[s], intensity_list[s]
        if pts is None or inten is None or pts.shape[0] == 0:
            continue
        if s >= S:
            raise IndexError(f"index {s * LIDAR_HEIGHT_CHANNELS} is out of bounds for axis 0 with size {S * LIDAR_HEIGHT_CHANNELS}")
        inten = np.asarray(inten)
        if inten.shape[0] != pts.shape[0]:
            raise IndexError("boolean index did not match: points and intensity lengths differ")
        inten = np.ascontiguousarray(inten, dtype=np.float32).reshape(-1)       # utils.py:101: maximum.at into a float32 grid
        counts[s], rel[s] = pts.shape[0], pts.matrix
        a = used
        b = a + (pts.raw.nbytes + 15) // 16 * 16
        used = b + (inten.nbytes + 15) // 16 * 16
        parts.append((pts.raw, inten, a, b))
    off = np.zeros(S + 1, np.int32)
    off[1:] = np.cumsum(counts)
    total = int(off[-1])
    nbytes = max(1, (used + (4 << 20) - 1) // (4 << 20)) * (4 << 20)            # few distinct sizes: the pool keys on them
    stage = pinned_pool.take(nbytes)
    try:
        h = stage.numpy()
        for raw, inten, a, b in parts:
            flat = raw.reshape(-1, order="F" if raw.flags.f_contiguous and not raw.flags.c_contiguous else "C")
            h[a:a + raw.nbytes] = flat.view(np.uint8)
            h[b:b + inten.nbytes] = inten.view(np.uint8)
        with torch.cuda.device(device):
            d = stage[:used].to(device, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record()
            rows = torch.empty((total, 4), dtype=torch.float32, device=device)
            o = 0
            for raw, inten, a, b in parts:
                n = raw.shape[0]
                src = d[a:a + raw.nbytes].view(torch.float16 if raw.dtype == np.float16 else torch.float32)
                src = src.view(3, n).t() if raw.flags.f_contiguous and not raw.flags.c_contiguous else src.view(n, 3)
                rows[o:o + n, :3].copy_(src)
                rows[o:o + n, 3].copy_(d[b:b + 4 * n].view(torch.float32))
                o += n
            out = lidar_bev_from_sweeps(rows, off, rel.reshape(S, 16), 1, S, max_sweep_points=int(counts.max()) if S else 0, algo=algo)[0]
            copied.synchronize()                                                # the staging buffer is free again (kernels keep running)
    finally:
        pinned_pool._release(stage)
    return out


def _pack_sweeps(points_list: Sequence, intensity_list: Sequence, num_expected_sweeps: int):
    """Concatenate the reference's list-of-arrays form into (points float64 [N,3], intensity float32 [N],
    offsets int32 [S+1]); applies the reference's skip rules (utils.py:71-76)."""
    loaded = min(len(points_list), len(intensity_list))
    P, I = [], []
    counts = np.zeros(num_expected_sweeps, np.int64)
    for s in range(loaded):
        pts, inten = points_list[s], intensity_list[s]
        if pts is None or inten is None or pts.shape[0] == 0:
            continue
        if s >= num_expected_sweeps:
            # the reference indexes bev[channel_offset + h] out of range here (utils.py:102)
            raise IndexError(f"index {s * LIDAR_HEIGHT_CHANNELS} is out of bounds for axis 0 with size "
                             f"{num_expected_sweeps * LIDAR_HEIGHT_CHANNELS}")
        pts = np.asarray(pts)
        inten = np.asarray(inten)
        if inten.shape[0] != pts.shape[0]:
            raise IndexError("boolean index did not match: points and intensity lengths differ")
        P.append(np.ascontiguousarray(pts[:, :3], dtype=np.float64))
        I.append(np.ascontiguousarray(inten, dtype=np.float32))
        counts[s] = pts.shape[0]
    off = np.zeros(num_expected_sweeps + 1, np.int32)
    off[1:] = np.cumsum(counts)
    pts_all = np.concatenate(P, 0) if P else np.zeros((0, 3), np.float64)
    int_all = np.concatenate(I, 0) if I else np.zeros((0,), np.float32)
    return pts_all, int_all, off


def create_intentnet_lidar_bev(points_list, intensity_list, num_expected_sweeps: int = LIDAR_SWEEPS, *,
                               as_tensor: bool = False, algo: int = ALGO_TILED):
    """Drop-in for utils.py:62-106.

    points_list[s]   (N_s, >=3) points already in the current ego frame (float64 on the Dataset
                     path, dataset.py:340) or None;  intensity_list[s]  (N_s,) or None.
    Returns float32 ``[29*num_expected_sweeps, 400, 720]`` -- a fresh numpy array like the
    reference, or the CUDA tensor when ``as_tensor=True`` (skips the 334 MB device->host copy).
    Parity domain: float64 points (other dtypes are widened to float64 first), intensities
    finite or +inf and >= 0.
    """
    live = [p for p, i in zip(points_list, intensity_list) if p is not None and i is not None and p.shape[0] > 0]
    if live and all(isinstance(p, DeferredPoints) for p in live):
        out = _fused_from_deferred(points_list, intensity_list, num_expected_sweeps, algo)    # install() mode, dataset.py:340-347
    else:
        pts, inten, off = _pack_sweeps(points_list, intensity_list, num_expected_sweeps)
        out = lidar_bev_from_ego_points(pts, inten, off, 1, num_expected_sweeps, algo=algo)[0]
    return out if as_tensor else to_numpy_fresh(out)
