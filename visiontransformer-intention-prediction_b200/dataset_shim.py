"""Drop-in hooks for the reference's data layer (dataset.py).

1. ``install(dataset_module)`` -- rebinding, no source edits.  ``dataset.py:21-23`` does
   ``from utils import (load_ego_poses, transform_points, create_intentnet_lidar_bev,
   rasterize_map_ego_centric, prepare_gt_for_frame, augment_bev)``; those names live in the ``dataset``
   module's namespace, so replacing them there makes ``ArgoverseIntentNetDataset.__getitem__`` call the CUDA
   path with the reference's own signatures (numpy in / numpy out, dataset.py:340-353 untouched).
2. ``collate_fn`` -- same contract as dataset.py:137-150 (drops ``None`` items, stacks, returns ``None`` for an
   empty batch); works for CPU and CUDA tensors.
3. ``BevBatchLoader`` -- the fast path: whole batches generated on the device from raw sweeps
   (pipeline.BevGenerator), yielding the dict ``collate_fn`` would, already on the GPU, so the training loop's
   ``.to(DEVICE)`` (train_vit.py:154-155) is a no-op and the 2.75 GB/batch host stack + H2D copy disappears.
"""
from __future__ import annotations

import random
from typing import Callable, Iterator, List, Optional, Sequence

import numpy as np
import torch

from . import augment as _aug
from . import bev_ops, map_ops, pipeline

_REBOUND = ("transform_points", "create_intentnet_lidar_bev", "rasterize_map_ego_centric", "augment_bev")


def install(dataset_module, utils_module=None, defer_transform: bool = True) -> List[str]:
    """Rebind the hot-path names inside the reference's ``dataset`` (and optionally ``utils``) module.

    ``defer_transform``: ``dataset.transform_points`` -- called only at dataset.py:340, its results only handed to
    ``create_intentnet_lidar_bev`` at :347 -- returns ``bev_ops.DeferredPoints`` for float16 / float32 sweeps, so the ten
    per-sweep transforms and the binning of a sample run as one fused launch over the raw points (same bits; the
    array-like still turns into the float64 array for any other reader).  ``utils.transform_points`` stays eager."""
    impl = {"transform_points": bev_ops.transform_points,
            "create_intentnet_lidar_bev": bev_ops.create_intentnet_lidar_bev,
            "rasterize_map_ego_centric": map_ops.rasterize_map_ego_centric,
            "augment_bev": _aug.augment_bev}
    done = []
    for mod in (dataset_module, utils_module):
        if mod is None:
            continue
        for name in _REBOUND:
            if hasattr(mod, name):
                fn = impl[name]
                if name == "transform_points" and defer_transform and mod is dataset_module:
                    fn = bev_ops.transform_points_deferred
                setattr(mod, name, fn)
                done.append(f"{mod.__name__}.{name}")
        if mod is utils_module:
            for name, fn in (("random_flip_bev", _aug.random_flip_bev), ("random_rotate_bev", _aug.random_rotate_bev),
                             ("random_scale_bev", _aug.random_scale_bev), ("random_bev_dropout", _aug.random_bev_dropout),
                             ("world_to_bev_pixel", map_ops.world_to_bev_pixel)):
                if hasattr(mod, name):
                    setattr(mod, name, fn)
                    done.append(f"{mod.__name__}.{name}")
    return done


def collate_fn(batch: list) -> Optional[dict]:
    """dataset.py:137-150, device-agnostic."""
    batch = [item for item in batch if item is not None]
    if not batch:
        return None
    return {"lidar_bev": torch.stack([it["lidar_bev"] for it in batch]),
            "map_bev": torch.stack([it["map_bev"] for it in batch]),
            "gt_list": [it["gt"] for it in batch]}


class BevBatchLoader:
    """Iterates ``samples`` (objects with the synthetic.Frame fields plus optional ``gt`` dict of numpy arrays
    ``boxes_xywha`` (K,5) / ``intentions`` (K,)) in batches and yields device-resident model inputs.

    is_train=True applies the reference's augmentation: parameters are drawn per sample from Python's global
    ``random`` in the reference's order (one augment_bev worth of draws per sample, in batch order), GT is
    updated on the host (utils.py:401-414, :441-448, :478-479), grids on the device."""

    def __init__(self, samples: Sequence, batch_size: int, map_table: Optional[map_ops.MapTable] = None,
                 map_id_of: Optional[Callable] = None, is_train: bool = False, num_sweeps: int = 10,
                 algo: int = bev_ops.ALGO_TILED, drop_last: bool = False, keep_order: bool = True,
                 sweep_store=None) -> None:
        """``sweep_store`` (a ``sweep_store.SweepStore``): samples that carry ``sweep_keys`` (one hashable per sweep, e.g.
        (log id, timestamp)) have each sweep uploaded once and shared by every sample that names it -- the reference's
        samples are sliding windows that share 9 of their 10 sweeps (dataset.py:174-211)."""
        self.samples, self.bs, self.is_train, self.S = samples, int(batch_size), is_train, num_sweeps
        self.store = sweep_store
        self.map_id_of = map_id_of or (lambda s: 0)
        self.gen = pipeline.BevGenerator(map_table, algo=algo)
        self.with_map = map_table is not None
        self.drop_last = drop_last
        self.keep_order = keep_order     # False: within a batch, resampled frames come first (gt_list permuted alike)

    def _pack_from_store(self, chunk, ids):
        from . import sweep_store as ss
        frames = []
        for smp in chunk:
            if getattr(smp, "store_frame", None) is not None:        # the sample already names its sweeps in the store
                frames.append(smp.store_frame)
                continue
            off = np.asarray(smp.sweep_offsets)
            cols = getattr(smp, "columns", None)
            keys = list(smp.sweep_keys)
            for k, key in enumerate(keys):
                lo, hi = int(off[k]), int(off[k + 1])
                if hi <= lo:
                    keys[k] = None                                   # missing / empty sweep (utils.py:75-76)
                elif key not in self.store:
                    if cols is not None:
                        self.store.put(key, *(c[lo:hi] for c in cols))
                    else:                                            # row samples: split losslessly (or raise)
                        pc = bev_ops.PointColumns.from_points4(np.asarray(smp.points4)[lo:hi], pinned=False)
                        h, st, n = pc.buf.numpy(), pc.stride, pc.n
                        self.store.put(key, *(h[2 * j * st:2 * j * st + 2 * n].view(np.float16) for j in range(3)), h[6 * st:6 * st + n])
            frames.append(ss.StoreFrame(keys, np.asarray(smp.sweep_poses), np.asarray(smp.current_pose)))
        return ss.StoreBatch.pack(self.store, frames, self.S, map_ids=ids)

    def __len__(self) -> int:
        n = len(self.samples)
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def __iter__(self) -> Iterator[dict]:
        for i0 in range(0, len(self.samples), self.bs):
            chunk = [s for s in self.samples[i0:i0 + self.bs] if s is not None]
            if not chunk or (self.drop_last and len(chunk) < self.bs):
                continue
            aug = None
            gts = []
            if self.is_train:
                aug = [_aug.draw_aug_params(random) for _ in chunk]
                if not self.keep_order:      # frames that rotate / scale first: no identity copies (order is returned)
                    chunk, aug, _ = pipeline.partition_for_augmentation(chunk, aug)
            ids = [self.map_id_of(s) for s in chunk] if self.with_map else None
            if self.store is not None and all(getattr(s, "store_frame", None) is not None or
                                              getattr(s, "sweep_keys", None) is not None for s in chunk):
                fb = self._pack_from_store(chunk, ids)
            else:
                fb = pipeline.FrameBatch.pack(chunk, self.S, map_ids=ids)
            for k, s in enumerate(chunk):
                gt = getattr(s, "gt", None) or {"boxes_xywha": np.zeros((0, 5), np.float32), "intentions": np.zeros((0,), np.int64)}
                boxes = np.array(gt["boxes_xywha"], dtype=np.float32, copy=True)
                intents = np.array(gt["intentions"], dtype=np.int64, copy=True)
                if aug is not None:
                    _aug.apply_gt(aug[k], boxes, intents)
                gts.append({"boxes_xywha": torch.from_numpy(boxes).float(), "intentions": torch.from_numpy(intents).long()})
            # fresh output tensors per batch: the consumer may still hold the previous batch
            out = self.gen.generate(fb, aug, with_map=self.with_map)
            yield {"lidar_bev": out["lidar_bev"], "map_bev": out["map_bev"], "gt_list": gts}
