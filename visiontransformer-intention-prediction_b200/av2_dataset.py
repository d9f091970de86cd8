"""An AV2-format log directory -> device-resident BEV batches: the reference's ``ArgoverseIntentNetDataset`` + ``DataLoader`` +
``collate_fn`` (dataset.py:27-367) as one iterable, with the hot path on the GPU.

What is mirrored from the reference (so that the SAME directory yields the SAME samples in the SAME order):
  * which sub-directories count as scenarios: ``sensors/lidar/*.feather``, ``annotations.feather``, a non-empty ``map/`` with a
    ``log_map_archive_<log id>*.json`` (or any ``log_map_archive_*.json``), ``city_SE3_egovehicle.feather``, files of at least
    1 KiB (``ScenarioValidator._validate_scenario``, dataset.py:101-134)
  * samples = 10-sweep sliding windows over the sorted sweep timestamps, current timestamp = the newest sweep's
    (``_create_sequences``, dataset.py:174-211)
  * per sample: pose lookup, which sweeps count as missing, ``None`` samples (``__getitem__``, dataset.py:285-345; here
    ``sweep_io.load_frame``)
Ground truth: ``gt_of(sequence) -> {"boxes_xywha", "intentions"}``; the default, ``AnnotationGT``, is the host-side
mirror of ``prepare_gt_for_frame`` (utils.py:184-225) on the log's ``annotations_with_intent.feather`` (a handful of boxes
per frame; pandas / scipy exactly as the reference).  What is NOT here: everything after the batch (models, loss).

Sweeps are decoded as AV2 stores them (float16 x,y,z + uint8 intensity columns) and uploaded ONCE into a device-resident
``SweepStore``; every sample that names a sweep reads it there (the windows share 9 of their 10 sweeps).
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, Iterator, List, Optional

import numpy as np

from . import bev_ops, dataset_shim, map_ops, maps, pose, sweep_io, sweep_store
from .constants import LIDAR_SWEEPS, VEHICLE_CATEGORIES


def find_scenarios(data_dir, min_file_bytes: int = 1024) -> List[dict]:
    """dataset.py:40-134: the sub-directories of ``data_dir`` that are complete scenarios, in ``os.scandir`` order."""
    base = Path(data_dir)
    out = []
    if not base.is_dir():
        return out
    for entry in os.scandir(base):
        if not entry.is_dir():
            continue
        d = Path(entry.path)
        log_id = d.name
        lidar, ann, mp, ego = d / "sensors" / "lidar", d / "annotations.feather", d / "map", d / "city_SE3_egovehicle.feather"
        ok = True
        for p in (lidar, ann, mp, ego):
            if not p.exists() or (p.is_dir() and not any(p.iterdir())) or \
               (p.is_file() and min_file_bytes > 0 and p.stat().st_size < min_file_bytes):
                ok = False
        if not ok or not any(lidar.glob("*.feather")):
            continue
        files = list(mp.glob(f"log_map_archive_{log_id}*.json")) or list(mp.glob("log_map_archive_*.json"))
        if not files:
            continue
        out.append({"log_id": log_id, "log_dir": str(d), "map_json_path": str(files[0]), "annotations_path": str(ann)})
    return out


def make_sequences(scenarios: List[dict], num_sweeps: int = LIDAR_SWEEPS) -> List[dict]:
    """dataset.py:174-211: sliding windows of ``num_sweeps`` consecutive sweep files per log."""
    seqs = []
    for sc in scenarios:
        lidar = Path(sc["log_dir"]) / "sensors" / "lidar"
        try:
            ts = sorted(int(p.stem) for p in lidar.glob("*.feather"))
        except ValueError:
            continue                                                    # a non-numeric file name skips the log (dataset.py:206-207)
        for i in range(len(ts) - num_sweeps + 1):
            seqs.append({**sc, "current_ts_ns": ts[i + num_sweeps - 1], "sweep_ts_list": ts[i:i + num_sweeps]})
    return seqs


class AnnotationGT:
    """``gt_of`` callback: utils.py:184-225 (``prepare_gt_for_frame``) on ``<log>/annotations_with_intent.feather``
    (dataset.py:220-233) -- rows of the sample's timestamp whose category is a vehicle category and whose precomputed
    ``heuristic_intent`` is not -1 -> boxes ``[tx_m, ty_m, |width_m|, |length_m|, yaw]`` float32 (yaw =
    ``R.from_quat([qx,qy,qz,qw]).as_euler('xyz')[2]``) and intention ids int64, in file order; rows scipy rejects are
    skipped.  A log without the file yields no ground truth (the reference returns ``None`` for its samples).

    Library-version caveat: the reference hands scipy the OBJECT-dtype slice ``box_row[['qx','qy','qz','qw']].values``
    (utils.py:208).  scipy 1.13.1 (the reference's pin) converts it; scipy >= 1.14 raises ``ValueError`` ("Buffer dtype
    mismatch"), which the reference's own ``except (ValueError, KeyError)`` swallows -- so on a current scipy the
    reference silently drops EVERY box.  The default here converts the quaternion to float64 first (the pinned
    environment's behaviour); ``strict_reference=True`` makes the literal call and reproduces whatever the installed
    scipy does (this is what the dataset-level golden, generated in this container, contains)."""

    def __init__(self, strict_reference: bool = False) -> None:
        self._frames = {}
        self.strict_reference = strict_reference

    def __call__(self, seq: dict) -> Optional[dict]:
        from scipy.spatial.transform import Rotation as R
        df = self._frames.get(seq["log_id"], False)
        if df is False:
            path = Path(seq["log_dir"]) / "annotations_with_intent.feather"
            df = None
            if path.is_file():
                import pandas as pd
                df = pd.read_feather(path)
                if "heuristic_intent" not in df.columns:
                    df = None
            self._frames[seq["log_id"]] = df
        empty = {"boxes_xywha": np.zeros((0, 5), np.float32), "intentions": np.zeros((0,), np.int64)}
        if df is None:
            return empty
        rows = df[(df["timestamp_ns"] == seq["current_ts_ns"]) & (df["category"].isin(VEHICLE_CATEGORIES)) &
                  (df["heuristic_intent"] != -1)]
        boxes, intents = [], []
        for _, r in rows.iterrows():
            try:
                quat = r[["qx", "qy", "qz", "qw"]].values
                if not self.strict_reference:
                    quat = np.asarray(quat, dtype=np.float64)
                yaw = R.from_quat(quat).as_euler("xyz", degrees=False)[2]
                boxes.append([r["tx_m"], r["ty_m"], abs(r["width_m"]), abs(r["length_m"]), yaw])
                intents.append(int(r["heuristic_intent"]))
            except (ValueError, KeyError):
                continue
        if not boxes:
            return empty
        return {"boxes_xywha": np.asarray(boxes, np.float32), "intentions": np.asarray(intents, np.int64)}


class _Sample:
    """what Av2BevDataset hands to BevBatchLoader: ``store_frame`` (sweeps named in the store), ``map_index``, ``gt``"""
    store_frame = None
    map_index = 0
    gt = None
    sequence = None


class _LazySamples:
    """Sequence view BevBatchLoader slices: sample i is decoded when its batch is built."""

    def __init__(self, ds: "Av2BevDataset", order: List[int]) -> None:
        self.ds, self.order = ds, order

    def __len__(self) -> int:
        return len(self.order)

    def __getitem__(self, sl):
        idx = self.order[sl]
        return [self.ds.sample(i) for i in idx] if isinstance(sl, slice) else self.ds.sample(idx)


class Av2BevDataset:
    def __init__(self, data_dir, num_sweeps: int = LIDAR_SWEEPS, is_train: bool = False,
                 gt_of: Optional[Callable[[dict], dict]] = None, store_points: int = 64_000_000, device=None) -> None:
        self.num_sweeps, self.is_train = int(num_sweeps), is_train
        self.gt_of = gt_of if gt_of is not None else AnnotationGT()
        self.scenarios = find_scenarios(data_dir)
        if not self.scenarios:
            raise ValueError(f"No valid scenarios found in {data_dir}. Please check the path and data integrity.")   # dataset.py:166
        self.sequences = make_sequences(self.scenarios, self.num_sweeps)
        if not self.sequences:
            raise ValueError(f"Could not create any valid sequences from the provided scenarios in {data_dir}.")     # dataset.py:171
        self._poses = {}
        self._no_sweep = set()
        self._map_index = {sc["log_id"]: k for k, sc in enumerate(self.scenarios)}
        self.map_table = map_ops.MapTable([self._load_map(sc["map_json_path"]) for sc in self.scenarios], device=device)
        # AV2 stores float16 x,y,z + uint8 intensity; a directory with other dtypes (float32 exports) cannot live in the
        # column store and goes through the row path instead (every sample packs and uploads its own ten sweeps)
        self.columnar = all(self._probe_columnar(sc) for sc in self.scenarios)
        self.store = sweep_store.SweepStore(store_points, device=self.map_table.device) if self.columnar else None
        self._pose_frames = {}

    @staticmethod
    def _probe_columnar(sc: dict) -> bool:
        for p in sorted((Path(sc["log_dir"]) / "sensors" / "lidar").glob("*.feather")):
            cols = sweep_io.decode_sweep_columns(p, use_cache=False)
            if cols is not None:                                         # first readable, non-empty sweep decides
                return cols is not sweep_io.NOT_COLUMNAR
        return True

    @staticmethod
    def _load_map(path: str) -> maps.FlatMap:
        try:
            return maps.load_map_json(path)
        except Exception as e:  # noqa: BLE001 -- utils.py:114: unreadable map -> all-zero map + message
            print(f"Error loading map JSON {path}: {e}. Returning empty map.")
            return maps.parse_map_dict({})

    def __len__(self) -> int:
        return len(self.sequences)

    def _log_poses(self, seq: dict) -> dict:
        """timestamp -> float64 [7] (tx,ty,tz,qx,qy,qz,qw) of the log's FIRST pose row with that timestamp (what the
        reference's ``df[df.timestamp_ns == ts].iloc[0]`` picks, dataset.py:285-289,323-327), or ``None`` when scipy rejects
        the quaternion (dataset.py:293-296,330-334).  Built once per log instead of one pandas scan per lookup."""
        table = self._poses.get(seq["log_id"])
        if table is None:
            import pyarrow.feather as feather
            df = feather.read_feather(Path(seq["log_dir"]) / "city_SE3_egovehicle.feather")    # utils.py:22-25
            cols = np.stack([df[c].to_numpy(np.float64) for c in ("tx_m", "ty_m", "tz_m", "qx", "qy", "qz", "qw")], 1)
            table = {}
            for t, row in zip(df["timestamp_ns"].to_numpy().tolist(), cols):
                if t in table:
                    continue
                try:
                    pose.se3_from_pose(row[:3], row[3:7])
                    table[t] = row
                except ValueError:
                    table[t] = None
            self._poses[seq["log_id"]] = table
        return table

    def _resident(self, seq: dict, ts: int):
        """store key of sweep ``ts`` (decoded and uploaded on first use), or ``None`` for the reference's "skip this sweep"
        cases: file missing / empty table / unreadable (dataset.py:307-318)."""
        key = (seq["log_id"], int(ts))
        if key in self.store:
            return key
        if key in self._no_sweep:
            return None
        path = Path(seq["log_dir"]) / "sensors" / "lidar" / f"{ts}.feather"
        cols = sweep_io.decode_sweep_columns(path)
        if cols is sweep_io.NOT_COLUMNAR:                                 # other dtypes on disk: split the rows losslessly (or raise)
            rows = sweep_io.decode_sweep(path)
            pc = bev_ops.PointColumns.from_points4(rows.astype(np.float32), pinned=False)
            h, st, n = pc.buf.numpy(), pc.stride, pc.n
            cols = tuple(h[2 * j * st:2 * j * st + 2 * n].view(np.float16) for j in range(3)) + (h[6 * st:6 * st + n],)
        if cols is None:
            self._no_sweep.add(key)
            return None
        self.store.put(key, *cols)
        return key

    def sample(self, idx: int):
        """dataset.py:262-345 up to the hot path, or ``None`` where the reference returns ``None`` (no current pose row
        :286-287, invalid current quaternion :294-296, every sweep missing :344-345).  The sample names its sweeps in the
        store (``store_frame``); nothing is concatenated or re-uploaded."""
        seq = self.sequences[idx]
        if not self.columnar:
            return self._row_sample(seq)
        poses = self._log_poses(seq)
        cur = poses.get(seq["current_ts_ns"])
        if cur is None:
            if seq["current_ts_ns"] in poses:
                print(f"Warning: Invalid quaternion for ego pose at ts {seq['current_ts_ns']} in log {seq['log_id']}. Skipping item.")
            return None
        S = len(seq["sweep_ts_list"])
        keys: List[Optional[tuple]] = [None] * S
        sp = np.zeros((S, 7), np.float64)
        sp[:, 6] = 1.0                                                   # identity for missing sweeps (never used: no points)
        for k, ts in enumerate(seq["sweep_ts_list"]):
            prow = poses.get(ts)                                         # dataset.py:323-334: no / invalid pose row -> None sweep
            if prow is None:                                             # (checked before the file: such a sweep is never uploaded)
                continue
            key = self._resident(seq, ts)
            if key is None:
                continue
            keys[k], sp[k] = key, prow
        if all(k is None for k in keys):
            return None
        out = _Sample()
        out.store_frame = sweep_store.StoreFrame(keys, sp, cur)
        out.map_index = self._map_index[seq["log_id"]]
        out.sequence = seq
        out.gt = self.gt_of(seq) if self.gt_of is not None else None
        return out

    def _row_sample(self, seq: dict):
        """non-AV2 dtypes on disk: the sample carries its own points as rows (sweep_io.load_frame), no store"""
        df = self._pose_frames.get(seq["log_id"])
        if df is None:
            import pyarrow.feather as feather
            df = feather.read_feather(Path(seq["log_dir"]) / "city_SE3_egovehicle.feather")
            self._pose_frames[seq["log_id"]] = df
        fr = sweep_io.load_frame(seq["log_dir"], seq["sweep_ts_list"], seq["current_ts_ns"], df, layout="rows")
        if fr is None:
            return None
        fr.map_index = self._map_index[seq["log_id"]]
        fr.sequence = seq
        fr.gt = self.gt_of(seq) if self.gt_of is not None else None
        return fr

    def loader(self, batch_size: int = 8, shuffle: bool = False, rng: Optional[np.random.Generator] = None,
               drop_last: bool = False, decode_threads: int = 0, decode_ahead: int = 2) -> Iterator[dict]:
        """Batches like ``DataLoader(dataset, batch_size, shuffle, collate_fn=collate_fn)`` (train_vit.py:76-79), already on the
        device: ``{"lidar_bev", "map_bev", "gt_list"}``; ``None`` samples are dropped as collate_fn drops them.

        ``decode_threads`` > 0: the sweep files of the next ``decode_ahead`` batches are decoded (feather -> column arrays,
        into sweep_io's LRU) by a thread pool while the current batch is built and generated -- pyarrow releases the GIL, and
        after the kernels the decode is what bounds a from-disk loader.  Uploads and launches stay on the calling thread."""
        order = list(range(len(self)))
        if shuffle:
            (rng or np.random.default_rng()).shuffle(order)
        inner = dataset_shim.BevBatchLoader(_LazySamples(self, order), batch_size, self.map_table,
                                            map_id_of=lambda s: s.map_index, is_train=self.is_train,
                                            num_sweeps=self.num_sweeps, drop_last=drop_last, sweep_store=self.store)
        if decode_threads <= 0 or not self.columnar:
            return iter(inner)
        return self._prefetching(inner, order, batch_size, decode_threads, decode_ahead)

    def _prefetching(self, inner, order, batch_size, threads, ahead):
        from concurrent.futures import ThreadPoolExecutor
        asked = set()

        def request(pool, b):                                           # decode the not-yet-resident sweeps of batch b
            for i in order[b * batch_size:(b + 1) * batch_size]:
                seq = self.sequences[i]
                for ts in seq["sweep_ts_list"]:
                    key = (seq["log_id"], int(ts))
                    if key not in self.store and key not in asked:
                        asked.add(key)
                        pool.submit(sweep_io.decode_sweep_columns, Path(seq["log_dir"]) / "sensors" / "lidar" / f"{ts}.feather")

        with ThreadPoolExecutor(max_workers=threads) as pool:
            n_batches = (len(order) + batch_size - 1) // batch_size
            for b in range(min(ahead, n_batches)):
                request(pool, b)
            for b, batch in enumerate(inner):
                if b + ahead < n_batches:
                    request(pool, b + ahead)
                if len(asked) > 4096:
                    asked.clear()
                yield batch
