"""Device-resident sweep store: upload every sweep ONCE, name it in as many frames as use it.

The reference's samples are 10-sweep sliding windows over a log (dataset.py:174-211), so consecutive samples share 9 of
their 10 sweeps -- and its Dataset re-reads and re-transforms all 10 for every sample (dataset.py:305-342).  The per-sweep
pose transform depends on the sample's own ego pose and cannot be shared, but the raw points can: in AV2's column form a
100 k-point sweep is 0.7 MB, a log of ~150 sweeps ~100 MB, and a whole split fits in a B200's 180 GB of HBM.  With a store a
step moves only the sweeps that are not resident yet over PCIe (one per new sample in a sliding-window stream) instead of
ten per frame; the kernels are the same ones, reading their sweeps through a start row per sweep
(``bevgen_lidar_bev_store_cols``), and the grids are bit-identical.

    store = SweepStore(capacity_points=64_000_000)
    store.put(("log7", ts), x, y, z, intensity)                  # float16 x 3 + uint8, as the feather file holds them
    batch = StoreBatch.pack(store, [StoreFrame(keys, sweep_poses, current_pose), ...])
    out = BevGenerator(...).generate(batch)                       # same call as for a FrameBatch
"""
from __future__ import annotations

import collections
import dataclasses
import weakref
from typing import Hashable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native, bev_ops, pose
from .constants import LIDAR_SWEEPS


def _round8(n: int) -> int:
    return (int(n) + 7) // 8 * 8


class SweepStore:
    """Ring of sweeps in one device ``PointColumns`` buffer; first-in-first-out eviction when it wraps.

    Uploads run on the store's own copy stream and overlap the kernels of earlier steps; ``ready(stream)`` makes a compute
    stream wait for them.  ``StoreBatch.pack`` pins the sweeps it names; ``consumed(stream, batch)`` (made by
    ``BevGenerator.generate``) closes an epoch, unpins them and stamps them with it.  An upload that overwrites a sweep's
    rows first waits for the event of the sweep's last epoch -- not for the newest kernels, so uploads for step i+1 still
    overlap the kernels of step i.  Capacity must cover the sweeps of the batches in flight: a batch whose sweep is gone
    fails in ``StoreBatch.pack`` (KeyError), and an upload that would overwrite a pinned sweep raises -- never silently
    wrong.  Not thread-safe: one store per loader thread (decode in worker threads, ``put`` on the thread that launches)."""

    def __init__(self, capacity_points: int, device: Optional[torch.device] = None) -> None:
        self.device = torch.device(device) if device is not None else bev_ops._require_cuda()
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.capacity = max(8, _round8(capacity_points))
        with torch.cuda.device(self.device):
            buf = torch.zeros(7 * self.capacity, dtype=torch.uint8, device=self.device)
            self.copy_stream = torch.cuda.Stream(device=self.device)
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))     # uploads are ordered after the clear
        self.columns = bev_ops.PointColumns(buf, self.capacity, self.capacity)       # every row readable: n = capacity
        self._entries: "collections.OrderedDict[Hashable, Tuple[int, int]]" = collections.OrderedDict()   # ring order
        self._head = 0
        self._uploaded: Optional[torch.cuda.Event] = None
        self._dirty = False                     # uploads issued since the last event on the copy stream
        self._pending_sources = []              # host arrays of those uploads (kept alive until their event has completed)
        self._inflight = collections.deque()
        self._epoch = 0                         # number of consumed() calls so far
        self._events = {}                       # epoch -> event recorded by consumed()
        self._last_use = {}                     # key -> epoch of the last consumed batch that named it
        self._pins = collections.Counter()      # key -> packed batches that name it and have not been consumed yet
        self._wait_epoch = 0                    # newest epoch a sweep evicted since the last upload was used in
        self.h2d_bytes = 0                      # bytes uploaded since construction (bench accounting)
        self.evictions = 0

    def __contains__(self, key: Hashable) -> bool:
        return key in self._entries

    def __len__(self) -> int:
        return len(self._entries)

    def rows(self, key: Hashable) -> Tuple[int, int]:
        """(first row, number of points) of a resident sweep; KeyError when it is not (never uploaded, or evicted)."""
        return self._entries[key]

    def _pin(self, keys) -> None:
        self._pins.update(keys)

    def _unpin(self, keys) -> None:
        self._pins.subtract(keys)
        for k in keys:
            if self._pins[k] <= 0:
                del self._pins[k]

    def _evict_front(self) -> None:
        key = next(iter(self._entries))
        if self._pins.get(key, 0) > 0:
            raise RuntimeError("SweepStore too small: this upload would overwrite a sweep that a packed batch which has not been "
                               "generated yet still names")
        self._entries.popitem(last=False)
        self._wait_epoch = max(self._wait_epoch, self._last_use.pop(key, 0))
        self.evictions += 1

    def _alloc(self, need: int) -> int:
        if need > self.capacity:
            raise ValueError(f"sweep of {need} rows does not fit a store of {self.capacity}")
        # (_wait_epoch accumulates over the evictions below and is cleared by put() once the wait has been issued: an
        #  allocation that fails half-way -- pinned sweep -- leaves it set for the next one)
        if self._head + need > self.capacity:                       # wrap: the rows up to the end go first ...
            while self._entries and next(iter(self._entries.values()))[0] >= self._head:
                self._evict_front()
            self._head = 0
        lo, hi = self._head, self._head + need                      # ... then whatever lies under the new rows
        while self._entries:
            r, n = next(iter(self._entries.values()))
            if r < hi and lo < r + max(8, _round8(n)):
                self._evict_front()
            else:
                break
        self._head = hi
        return lo

    def put(self, key: Hashable, x, y, z, intensity) -> Tuple[int, int]:
        """Make sweep ``key`` resident (no-op when it is): ``x, y, z`` float16 and ``intensity`` uint8 arrays / tensors of one
        length, as AV2 stores them; pinned sources are copied asynchronously.  Returns (first row, points)."""
        hit = self._entries.get(key)
        if hit is not None:
            return hit
        ptrs, keep, n = [], [], None
        for c, dt, npdt in zip((x, y, z, intensity), (torch.float16,) * 3 + (torch.uint8,), (np.float16,) * 3 + (np.uint8,)):
            if isinstance(c, torch.Tensor):
                if c.dtype != dt or c.is_cuda or not c.is_contiguous():
                    c = None
                else:
                    ptrs.append(c.data_ptr()); m = int(c.numel())
            else:
                c = np.ascontiguousarray(c)
                if c.dtype != npdt:
                    c = None
                else:
                    ptrs.append(c.ctypes.data); m = int(c.size)
            if c is None:
                raise ValueError("SweepStore.put takes contiguous host arrays in AV2's stored dtypes: float16 x, y, z and uint8 "
                                 "intensity (PointColumns.from_arrays converts losslessly or raises)")
            keep.append(c)
            if n is None:
                n = m
            elif m != n:
                raise ValueError("x, y, z, intensity must have the same length")
        row = self._alloc(max(8, _round8(n)))
        cap, base = self.capacity, self.columns.buf.data_ptr()
        ev_used = self._wait_event(self._wait_epoch)
        if ev_used is not None:
            self.copy_stream.wait_event(ev_used)                     # the last kernels that read the rows being overwritten
        self._wait_epoch = 0
        # one call, four async copies on the store's copy stream (pageable sources are staged by the driver before it returns)
        rc = _native.lib().bevgen_store_put_cols(base + 2 * row, base + 2 * (cap + row), base + 2 * (2 * cap + row), base + 6 * cap + row,
                                                 *ptrs, n, self.copy_stream.cuda_stream)
        _native.check(rc, "bevgen_store_put_cols")
        self._pending_sources.append(keep)                           # pinned sources must outlive the copy
        self._dirty = True
        self._entries[key] = (row, n)
        self.h2d_bytes += 7 * n
        return row, n

    def _wait_event(self, epoch: int):
        """event an upload over rows last read in ``epoch`` must wait for: that epoch's own, or -- when it has been pruned --
        the oldest one still kept (recorded later, so waiting for it is merely conservative); ``None``: never read"""
        if epoch <= 0:
            return None
        ev = self._events.get(epoch)
        if ev is None and self._events:
            ev = self._events[min(self._events)]
        return ev

    def check_resident(self, batch: "StoreBatch") -> None:
        """KeyError unless every sweep ``batch`` names still sits where it was when the batch was packed.  A batch is
        protected from eviction until its first ``generate``; generating it AGAIN is fine as long as no upload has taken its
        rows in between -- this check turns that case into an error instead of a grid built from another sweep's points."""
        for k, where in batch.rows.items():
            if self._entries.get(k) != where:
                raise KeyError(f"sweep {k!r} named by this batch is no longer resident in the store (evicted after the batch's "
                               "first generate call); pack the batch again")

    def upload_tables(self, tables: torch.Tensor) -> torch.Tensor:
        """A batch's per-frame tables (one pinned buffer, ``StoreBatch.tables``) -> device, on the store's copy stream like the
        sweeps themselves: the compute stream carries no copies at all, only the wait that ``ready`` inserts.  (Four small
        copies on the compute stream cost a sliding-window step 0.1 ms of idle GPU between its kernels.)"""
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            dt = tables.to(self.device, non_blocking=True)
        dt.record_stream(main)
        self._dirty = True
        self._pending_sources.append(tables)
        return dt

    def ready(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """``stream`` (default: current) waits for every upload issued so far."""
        if self._dirty:
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self._uploaded, self._dirty = ev, False
            self._inflight.append((ev, self._pending_sources))
            self._pending_sources = []
            while self._inflight and self._inflight[0][0].query():   # drop references to sources whose copies have completed
                self._inflight.popleft()
        if self._uploaded is not None:
            (stream or torch.cuda.current_stream(self.device)).wait_event(self._uploaded)

    def consumed(self, stream: Optional[torch.cuda.Stream] = None, batch: Optional["StoreBatch"] = None) -> None:
        """Call after enqueuing the kernels that read ``batch``'s sweeps: uploads that later overwrite those sweeps wait for
        this point of ``stream``."""
        ev = torch.cuda.Event()
        ev.record(stream or torch.cuda.current_stream(self.device))
        self._epoch += 1
        self._events[self._epoch] = ev
        if batch is not None:
            for k in batch.keys:
                if k in self._entries:
                    self._last_use[k] = self._epoch
            batch._release()
        if len(self._events) > 64:              # older epochs: their kernels finished long before any row comes round again
            for k in sorted(self._events)[:-64]:
                del self._events[k]


@dataclasses.dataclass
class StoreFrame:
    """One sample of a sliding-window stream: the keys of its sweeps (oldest first; ``None`` = missing sweep, its 29 channels
    stay zero like the reference's ``None`` entries, utils.py:75-76) and the poses dataset.py:285-339 looks up for them."""
    keys: Sequence[Optional[Hashable]]
    sweep_poses: np.ndarray          # float64 [S,7] tx,ty,tz,qx,qy,qz,qw
    current_pose: np.ndarray         # float64 [7]
    gt: Optional[dict] = None
    rel_tf: Optional[np.ndarray] = None      # [len(keys),4,4] if the caller has done the pose prep already (pose.sweep_rel_transforms)
    ego_T: Optional[np.ndarray] = None       # [3,3]                                                       (pose.map_matrix_from_pose)


@dataclasses.dataclass
class StoreBatch:
    """What ``BevGenerator.generate`` needs for a batch of StoreFrames: a few hundred bytes per frame, pinned."""
    store: SweepStore
    sweep_start: torch.Tensor        # int32 [B*S]   first store row of every sweep (0 for missing ones)
    sweep_offsets: torch.Tensor      # int32 [B*S+1] exclusive prefix sum of the sweep lengths
    rel_tf: torch.Tensor             # float64 [B*S,16]
    ego_T: torch.Tensor              # float64 [B,9]
    map_ids: np.ndarray
    num_frames: int
    num_sweeps: int
    max_sweep_points: int
    total_points: int
    keys: tuple = ()                 # distinct sweeps named (pinned in the store until the batch is generated or dropped)
    tables: Optional[torch.Tensor] = None    # the four arrays above as ONE pinned uint8 buffer (they are views of it): one copy
    layout: tuple = ()               # (name, dtype, byte offset, elements) of each view inside ``tables``
    rows: dict = dataclasses.field(default_factory=dict)     # key -> (first row, points) at pack time (SweepStore.check_resident)
    _finalizer: object = None

    points_format = "av2-store"
    point_bytes = 7

    @property
    def num_points(self) -> int:
        return self.total_points

    @property
    def h2d_bytes(self) -> int:
        """per-call bytes (the sweeps themselves are accounted by ``SweepStore.h2d_bytes`` when they are uploaded)"""
        return int(self.sweep_start.numel() * 4 + self.sweep_offsets.numel() * 4 + self.rel_tf.numel() * 8 + self.ego_T.numel() * 8)

    def _release(self) -> None:
        """unpin the batch's sweeps (idempotent; also runs when the batch is garbage-collected without being generated)"""
        if self._finalizer is not None and self._finalizer.alive:
            self._finalizer()

    @staticmethod
    def pack(store: SweepStore, frames: Sequence[StoreFrame], num_sweeps: int = LIDAR_SWEEPS,
             map_ids: Optional[Sequence[int]] = None) -> "StoreBatch":
        B, S = len(frames), int(num_sweeps)
        start = np.zeros((B, S), np.int32)
        counts = np.zeros((B, S), np.int64)
        rel = np.tile(np.eye(4), (B, S, 1, 1))
        ego = np.tile(np.eye(3), (B, 1, 1))
        for b, f in enumerate(frames):
            ns = len(f.keys)
            if ns > S:
                raise IndexError(f"frame {b}: {ns} sweeps but num_sweeps={S}")      # utils.py:102
            for s, key in enumerate(f.keys):
                if key is None:
                    continue
                start[b, s], counts[b, s] = store.rows(key)                         # KeyError: not resident (any more)
            if f.rel_tf is not None:
                rel[b, :ns] = f.rel_tf
            else:
                rel[b, :ns] = pose.sweep_rel_transforms(f.current_pose, f.sweep_poses)  # dataset.py:285-339
            if f.ego_T is not None:
                ego[b] = f.ego_T
            else:
                cp = np.asarray(f.current_pose, np.float64)
                ego[b] = pose.map_matrix_from_pose(cp[0], cp[1], cp[3:7])           # utils.py:121-129
        off = np.zeros(B * S + 1, np.int64)
        off[1:] = np.cumsum(counts.reshape(-1))
        if off[-1] >= 2 ** 31:
            raise ValueError("more than 2^31 points in one batch")
        # one pinned buffer for the four tables (float64 ones first: every view stays aligned), uploaded with one copy
        parts = [("rel_tf", rel.reshape(-1), torch.float64), ("ego_T", ego.reshape(-1), torch.float64),
                 ("sweep_start", start.reshape(-1), torch.int32), ("sweep_offsets", off.astype(np.int32), torch.int32)]
        nbytes = sum(a.nbytes for _, a, _ in parts)
        tables = torch.empty(max(nbytes, 8), dtype=torch.uint8)
        if torch.cuda.is_available():
            tables = tables.pin_memory()
        h, o, layout, views = tables.numpy(), 0, [], {}
        for name, a, dt in parts:
            h[o:o + a.nbytes] = np.ascontiguousarray(a).view(np.uint8)
            views[name] = tables[o:o + a.nbytes].view(dt)
            layout.append((name, dt, o, a.size))
            o += a.nbytes
        ids = np.full(B, -1, np.int64) if map_ids is None else np.asarray(map_ids, np.int64)
        keys = tuple({k for f in frames for k in f.keys if k is not None})
        batch = StoreBatch(store, views["sweep_start"], views["sweep_offsets"], views["rel_tf"].reshape(B * S, 16), views["ego_T"].reshape(B, 9),
                           ids, B, S, int(counts.max()) if B else 0, int(off[-1]), keys, tables, tuple(layout), {k: store.rows(k) for k in keys})
        store._pin(keys)
        batch._finalizer = weakref.finalize(batch, store._unpin, keys)
        return batch
