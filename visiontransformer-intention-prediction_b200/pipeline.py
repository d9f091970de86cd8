"""Batched frame -> BEV tensors pipeline: what ``Dataset.__getitem__`` x B + ``collate_fn`` do in the
reference (dataset.py:262-361, :137-150), as a handful of kernel launches per batch.

    host frames (raw sweep points, poses, map id)
        -> FrameBatch.pack()      pose prep with the reference's numpy/scipy calls (dataset.py:285-301,
                                  :323-339; utils.py:121-129), everything packed into PINNED host buffers
        -> BevGenerator.generate  H2D copies, LiDAR kernels, map rasterizer, [fused augmentation]
        -> {"lidar_bev": [B,290,400,720], "map_bev": [B,9,400,720]} float32 CUDA tensors, i.e. exactly
           what collate_fn stacks and train_vit.py:154-159 moves to the device and feeds to the model.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence

import numpy as np
import torch

from . import augment as _aug
from . import bev_ops, map_ops, pose
from .constants import (GRID_HEIGHT_PX, GRID_WIDTH_PX, LIDAR_HEIGHT_CHANNELS, LIDAR_SWEEPS, MAP_CHANNELS)


def _pinned(arr: np.ndarray) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(arr))
    return t.pin_memory() if torch.cuda.is_available() else t


@dataclasses.dataclass
class FrameBatch:
    """Host-side (pinned) packed batch."""
    points4: object              # torch float32 / float16 [N,4] rows, or bev_ops.PointColumns (AV2 columns, 7 B / point)
    sweep_offsets: torch.Tensor  # int32 [B*S+1]
    rel_tf: torch.Tensor         # float64 [B*S,16]
    ego_T: torch.Tensor          # float64 [B,9]
    map_ids: np.ndarray          # int64 [B]  (-1 = no map)
    num_frames: int
    num_sweeps: int
    max_sweep_points: int

    @property
    def num_points(self) -> int:
        return self.points4.n if isinstance(self.points4, bev_ops.PointColumns) else int(self.points4.shape[0])

    @property
    def points_format(self) -> str:
        if isinstance(self.points4, bev_ops.PointColumns):
            return "av2"
        return "f16" if self.points4.dtype == torch.float16 else "f32"

    @property
    def point_bytes(self) -> int:
        return {"av2": 7, "f16": 8, "f32": 16}[self.points_format]

    @property
    def h2d_bytes(self) -> int:
        pb = self.points4.nbytes if isinstance(self.points4, bev_ops.PointColumns) else self.points4.numel() * self.points4.element_size()
        return int(pb + self.sweep_offsets.numel() * 4 + self.rel_tf.numel() * 8 + self.ego_T.numel() * 8)

    @staticmethod
    def pack(frames: Sequence, num_sweeps: int = LIDAR_SWEEPS, map_ids: Optional[Sequence[int]] = None,
             points_dtype=None) -> "FrameBatch":
        """frames: objects with points4 [n,4] (or ``columns`` = (x, y, z, intensity) arrays), sweep_offsets [s+1],
        sweep_poses [s,7], current_pose [7] (e.g. synthetic.Frame, sweep_io.SweepFrame).  Sweeps beyond ``num_sweeps`` are
        an error, as in the reference (utils.py:102).  ``points_dtype``: np.float32 (16 B/point rows), np.float16 (8 B/point
        rows) or "av2" (AV2's on-disk columns, 7 B/point); default = the narrowest the frames already come in."""
        B, S = len(frames), int(num_sweeps)
        counts = np.zeros((B, S), np.int64)
        rel = np.tile(np.eye(4), (B, S, 1, 1))
        ego = np.tile(np.eye(3), (B, 1, 1))
        for b, f in enumerate(frames):
            ns = len(f.sweep_offsets) - 1
            if ns > S:
                raise IndexError(f"frame {b}: {ns} sweeps but num_sweeps={S}")
            counts[b, :ns] = np.diff(f.sweep_offsets)
            rel[b, :ns] = pose.sweep_rel_transforms(f.current_pose, f.sweep_poses)      # dataset.py:285-339
            cp = np.asarray(f.current_pose, np.float64)
            ego[b] = pose.map_matrix_from_pose(cp[0], cp[1], cp[3:7])                   # utils.py:121-129
        off = np.zeros(B * S + 1, np.int64)
        off[1:] = np.cumsum(counts.reshape(-1))
        if off[-1] >= 2 ** 31:
            raise ValueError("more than 2^31 points in one batch")
        want_cols = (isinstance(points_dtype, str) and points_dtype == "av2") or \
                    (points_dtype is None and B > 0 and all(getattr(f, "columns", None) is not None for f in frames))
        if want_cols:
            # AV2 columns: frames that carry (x, y, z, intensity) column arrays (sweep_io.load_frame(layout="columns")) are
            # concatenated column by column -- no interleave; row frames are split (lossless or an error)
            parts = [f.columns if getattr(f, "columns", None) is not None else tuple(np.asarray(f.points4).reshape(-1, 4)[:, k] for k in range(4))
                     for f in frames]
            cat = [np.concatenate([p[k] for p in parts]) if B else np.zeros(0, np.float16 if k < 3 else np.uint8) for k in range(4)]
            ids = np.full(B, -1, np.int64) if map_ids is None else np.asarray(map_ids, np.int64)
            return FrameBatch(bev_ops.PointColumns.from_arrays(*cat), _pinned(off.astype(np.int32)), _pinned(rel.reshape(B * S, 16)),
                              _pinned(ego.reshape(B, 9)), ids, B, S, int(counts.max()) if B else 0)
        pts = np.concatenate([np.asarray(f.points4).reshape(-1, 4) for f in frames], 0) if B else np.zeros((0, 4), np.float32)
        if points_dtype is None:             # auto: keep fp16 when every frame already is (sweep_io.load_frame)
            points_dtype = np.float16 if (B and all(np.asarray(f.points4).dtype == np.float16 for f in frames)) else np.float32
        if points_dtype == np.float16:
            # lossless only if the values are fp16-representable (AV2 sweeps are stored as fp16 x,y,z + u8 intensity)
            half = pts.astype(np.float16)
            if not np.array_equal(half.astype(np.float32), pts.astype(np.float32), equal_nan=True):
                raise ValueError("points are not exactly representable in float16; pack them as float32")
            pts = half
        else:
            pts = pts.astype(np.float32)
        ids = np.full(B, -1, np.int64) if map_ids is None else np.asarray(map_ids, np.int64)
        return FrameBatch(_pinned(pts), _pinned(off.astype(np.int32)), _pinned(rel.reshape(B * S, 16)),
                          _pinned(ego.reshape(B, 9)), ids, B, S, int(counts.max()) if B else 0)


def partition_for_augmentation(frames: Sequence, aug: Sequence[_aug.AugParams]):
    """Stable partition of a batch into [frames that rotate / scale ..., frames that do not ...].

    Returns (frames, aug, order) with ``order[k]`` = original index of the frame now at position k.  A batch packed
    in this order lets ``BevGenerator.generate`` write the non-resampled frames straight into the output tensor
    (flip / dropout fused into generation) and run the gather kernel only over the frames that need it -- no identity
    copies.  Batch order carries no meaning for training; permute ``gt_list`` with ``order`` to keep items aligned."""
    idx = [i for i, p in enumerate(aug) if p.needs_gather] + [i for i, p in enumerate(aug) if not p.needs_gather]
    return [frames[i] for i in idx], [aug[i] for i in idx], idx


class BevGenerator:
    """Owns the device-side state (map table, workspaces) and turns FrameBatches into model inputs."""

    def __init__(self, map_table: Optional[map_ops.MapTable] = None, algo: int = bev_ops.ALGO_TILED,
                 device: Optional[torch.device] = None, fused_augment: bool = True, map_stream_priority: int = 0) -> None:
        self.device = torch.device(device) if device is not None else bev_ops._require_cuda()
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if map_table is not None and map_table.device != self.device:
            raise ValueError(f"map table lives on {map_table.device}, generator on {self.device}")
        self.map_table = map_table
        self.algo = algo
        # LiDAR frames that rotate / scale: built directly in augmented form from the binned point records (one write of
        # the grid) instead of generate -> gather (three passes over it).  Tiled algorithm only; same bits either way.
        self.fused_augment = fused_augment
        self._scratch = {}
        # ALGO_AUTO: TILED until the scenes overflow the tile bins (dense / skewed point clouds: BASELINE config 4), then
        # ATOMIC, re-probing with one TILED call every 32 batches; the overflow total comes back asynchronously (4 bytes)
        self._auto = {"algo": bev_ops.ALGO_TILED, "pending": None, "calls": 0}
        # the map rasterizer is independent of the LiDAR kernels and far from saturating the GPU (one warp per lane
        # segment): it runs on a side stream and is joined before the tensors are handed out
        # (measured on config 3: default priority 16 700 frames/s; a high-priority map stream delays the binning pass: 16 440)
        self._map_stream = torch.cuda.Stream(device=self.device, priority=map_stream_priority)

    def _buf(self, key, shape) -> torch.Tensor:
        t = self._scratch.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.float32, device=self.device)
            self._scratch[key] = t
        return t

    def upload(self, batch) -> dict:
        """Async H2D of the packed batch (pinned -> device) on the current stream.  A ``sweep_store.StoreBatch`` uploads only
        its per-frame tables: the points are the store's resident columns (the stream waits for the store's uploads)."""
        d = self.device
        store = getattr(batch, "store", None)
        if store is not None:
            if store.device != d:
                raise ValueError(f"sweep store lives on {store.device}, generator on {d}")
            B, S = batch.num_frames, batch.num_sweeps
            dt = store.upload_tables(batch.tables)               # one copy, on the store's copy stream
            dv = {name: dt[o:o + n * (8 if t == torch.float64 else 4)].view(t) for name, t, o, n in batch.layout}
            dv["rel_tf"], dv["ego_T"] = dv["rel_tf"].view(B * S, 16), dv["ego_T"].view(B, 9)
            dv["points4"] = store.columns
            store.ready(torch.cuda.current_stream(d))
            return dv
        return {"sweep_offsets": batch.sweep_offsets.to(d, non_blocking=True), "rel_tf": batch.rel_tf.to(d, non_blocking=True),
                "ego_T": batch.ego_T.to(d, non_blocking=True), "points4": batch.points4.to(d, non_blocking=True)}

    def generate(self, batch: FrameBatch, aug: Optional[Sequence[_aug.AugParams]] = None, *, dev: Optional[dict] = None,
                 with_map: bool = True, out_lidar: Optional[torch.Tensor] = None,
                 out_map: Optional[torch.Tensor] = None) -> dict:
        """Returns {"lidar_bev", "map_bev"} CUDA tensors.  ``aug`` = one AugParams per frame (train time):
        flip / dropout of frames that need no resampling are fused into generation (zero extra bytes); frames that
        rotate or scale are built directly in augmented form (LiDAR from the binned point records, map channels from
        the rasterizer's bit planes) -- or, with ``fused_augment=False`` / parameters outside the fused kernel's reach,
        generated into scratch and resampled by the dense gather kernel.  Same bits either way."""
        with torch.cuda.device(self.device):      # the generator's device, whatever the caller's current device is
            if self.algo != bev_ops.ALGO_AUTO:
                return self._generate(batch, aug, dev, with_map, out_lidar, out_map, self.algo)
            resampling = aug is not None and any(p.needs_gather for p in aug)
            algo, probe = self._auto_pick(resampling)
            out = self._generate(batch, aug, dev, with_map, out_lidar, out_map, algo)
            if probe and batch.num_points > 0:
                host, ev = bev_ops.overflow_total_async(batch.num_points, batch.num_frames, batch.num_sweeps, self.device)
                self._auto["pending"] = (host, ev, batch.num_points)
            return out

    def _auto_pick(self, resampling: bool):
        """ALGO_AUTO policy -> (algorithm of this call, probe its overflow total?).  Overflow fraction of the last probed TILED
        call > 25 % -> ATOMIC (config 4: ~90 %), < 10 % -> TILED (config 2: 0.6 %).  While TILED every 8th call is probed;
        while ATOMIC every 32nd call runs TILED to re-probe.  Batches with resampled frames always run TILED (the fused
        augmentation kernels read the tile bins) and are not probed (their workspace is laid out differently)."""
        st = self._auto
        st["calls"] += 1
        if st["pending"] is not None and st["pending"][1].query():
            host, _, n = st["pending"]
            frac = float(host.item()) / max(1, n)
            st["pending"] = None
            if frac > 0.25:
                st["algo"] = bev_ops.ALGO_ATOMIC
            elif frac < 0.10:
                st["algo"] = bev_ops.ALGO_TILED
        if resampling:
            return bev_ops.ALGO_TILED, False
        if st["algo"] == bev_ops.ALGO_TILED:
            return bev_ops.ALGO_TILED, st["calls"] % 8 == 1
        if st["calls"] % 32 == 0:
            return bev_ops.ALGO_TILED, True
        return st["algo"], False

    def _generate(self, batch, aug, dev, with_map, out_lidar, out_map, algo: int) -> dict:
        B, S = batch.num_frames, batch.num_sweeps
        dv = dev if dev is not None else self.upload(batch)
        flip = rects = cnt = None
        J = 0                                   # dense route: frames [0, J) go through the gather kernel
        packed = None                           # fused route: the packed per-frame parameters
        if aug is not None:
            if len(aug) != B:
                raise ValueError("one AugParams per frame")
            need = [p.needs_gather for p in aug]
            J = sum(need)
            if J and self.fused_augment and algo == bev_ops.ALGO_TILED:
                packed = _aug.pack_params(aug)
                if not _aug.fused_supported(aug, packed):
                    packed = None               # parameters outside the fused kernel's reach: dense route
            if packed is None:
                flip, rects, cnt = _aug.fusion_arrays(aug)
                if J and not all(need[:J]):
                    J = B                       # not partitioned (partition_for_augmentation): gather the whole batch
        lshape = (B, S * LIDAR_HEIGHT_CHANNELS, GRID_HEIGHT_PX, GRID_WIDTH_PX)
        mshape = (B, MAP_CHANNELS, GRID_HEIGHT_PX, GRID_WIDTH_PX)
        if with_map and (self.map_table is None or (batch.map_ids < 0).any()):
            raise ValueError("with_map=True needs a MapTable and a map id per frame")
        if out_lidar is None:
            out_lidar = torch.empty(lshape, dtype=torch.float32, device=self.device)
        if with_map and out_map is None:
            out_map = torch.empty(mshape, dtype=torch.float32, device=self.device)
        d_flip = None if flip is None else bev_ops._dev(flip, torch.uint8, self.device)
        d_rects = None if rects is None else bev_ops._dev(rects, torch.int32, self.device)
        d_cnt = None if cnt is None else bev_ops._dev(cnt, torch.int32, self.device)
        d_T = dv["ego_T"]
        ids = batch.map_ids

        store = getattr(batch, "store", None)
        if store is not None:
            store.check_resident(batch)          # (a batch generated a second time after its sweeps were evicted: KeyError)
            if dev is not None:
                store.ready(torch.cuda.current_stream(self.device))
        d_start = dv.get("sweep_start")          # sweep store form: first store row per sweep
        skw = {} if d_start is None else {"total_points": batch.total_points}

        def run(lo: int, hi: int, dst_l: torch.Tensor, dst_m: Optional[torch.Tensor]) -> None:
            """generate frames [lo, hi) into dst_l / dst_m (sub-ranges are pointer offsets: CSR offsets are absolute)"""
            sl = slice(lo, hi)
            bev_ops.lidar_bev_from_sweeps(dv["points4"], dv["sweep_offsets"][lo * S:hi * S + 1], dv["rel_tf"][lo * S:hi * S],
                                          hi - lo, S, max_sweep_points=batch.max_sweep_points,
                                          flip=None if d_flip is None else d_flip[sl],
                                          drop_rects=None if d_rects is None else d_rects[sl],
                                          drop_counts=None if d_cnt is None else d_cnt[sl], out=dst_l, algo=algo,
                                          sweep_start=None if d_start is None else d_start[lo * S:hi * S], **skw)
            if with_map:
                with torch.cuda.stream(self._map_stream):
                    map_ops.rasterize_maps(self.map_table, ids[sl], d_T[sl], flip=None if d_flip is None else d_flip[sl],
                                           drop_rects=None if d_rects is None else d_rects[sl],
                                           drop_counts=None if d_cnt is None else d_cnt[sl], out=dst_m)

        main = torch.cuda.current_stream()
        if with_map:
            self._map_stream.wait_stream(main)          # inputs (uploads, fusion arrays) are ready on the main stream
        if packed is not None:
            # LiDAR: one pass for the whole batch, whatever the frame order (no partition needed)
            _aug.lidar_bev_augmented(dv["points4"], dv["sweep_offsets"], dv["rel_tf"], B, S, aug,
                                     max_sweep_points=batch.max_sweep_points, out=out_lidar, packed=packed,
                                     sweep_start=d_start, **skw)
            if with_map:
                # map channels: resampled straight from the rasterizer's bit planes, on the side stream
                with torch.cuda.stream(self._map_stream):
                    map_ops.rasterize_maps_augmented(self.map_table, ids, d_T, aug, out=out_map, packed=packed)
        elif J == 0:
            run(0, B, out_lidar, out_map)
        else:
            gen_l = self._buf("gen_l", (J,) + lshape[1:])
            gen_m = self._buf("gen_m", (J,) + mshape[1:]) if with_map else None
            run(0, J, gen_l, gen_m)
            if J < B:
                run(J, B, out_lidar[J:], out_map[J:] if with_map else None)
            # flip / dropout of frames that need no resampling were fused above; identity for them if they are in [0, J)
            gp = [p if p.needs_gather else _aug.AugParams() for p in aug[:J]]
            _aug.augment_grid(gen_l, gp, out=out_lidar[:J])
            if with_map:
                main.wait_stream(self._map_stream)
                _aug.augment_grid(gen_m, gp, out=out_map[:J])
        if with_map:
            main.wait_stream(self._map_stream)
        if store is not None:
            store.consumed(main, batch)          # rows read by this call may be overwritten by uploads ordered after this point
        lid, mp = out_lidar, (out_map if with_map else None)
        return {"lidar_bev": lid, "map_bev": mp}


class PrefetchingGenerator:
    """Overlaps the host->device copy of batch i+1 (own copy stream) with the generation of batch i.

        pf = PrefetchingGenerator(gen)
        h = pf.submit(batch0)
        for nxt in batches[1:] + [None]:
            cur, h = h, (pf.submit(nxt) if nxt is not None else None)
            out = pf.run(cur, aug=...)
    """

    def __init__(self, gen: BevGenerator) -> None:
        self.gen = gen
        self.copy_stream = torch.cuda.Stream(device=gen.device)

    def submit(self, batch: FrameBatch):
        with torch.cuda.stream(self.copy_stream):
            dev = self.gen.upload(batch)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return batch, dev, ev

    def run(self, handle, aug: Optional[Sequence[_aug.AugParams]] = None, **kw) -> dict:
        batch, dev, ev = handle
        cur = torch.cuda.current_stream(self.gen.device)
        cur.wait_event(ev)
        store = getattr(batch, "store", None)
        for t in dev.values():
            if store is None or t is not store.columns:
                t.record_stream(cur)      # allocated on the copy stream, consumed on the compute stream
        return self.gen.generate(batch, aug, dev=dev, **kw)
