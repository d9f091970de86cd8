"""Dev timing FROM DISK: a synthetic log in AV2's on-disk layout (feather sweeps of float16 x,y,z + uint8 intensity, pose table,
map JSON) read by (a) the reference's numpy/OpenCV call sequence, one sample at a time as its Dataset does, and (b)
``Av2BevDataset`` (feather decode -> sweep store -> kernels), evaluation mode, consecutive samples.
usage: python tools/time_av2_dir.py [sweeps=90] [points=100000] [batch=8] [decode_threads=8]"""
import json, shutil, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, pandas as pd, torch
import bevgen_b200 as bg

n_sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 90
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
threads = int(sys.argv[4]) if len(sys.argv) > 4 else 8
root = Path(tempfile.mkdtemp(prefix="bevgen_av2_"))
log = root / "synthlog00"
(log / "sensors" / "lidar").mkdir(parents=True); (log / "map").mkdir()
sweeps, poses = bg.synthetic.sliding_log(16, n=n_points, seed=3)
poses = bg.synthetic._poses(n_sweeps)
poses[:, 0] += bg.synthetic.EGO_XY[0]; poses[:, 1] += bg.synthetic.EGO_XY[1]
ts = [315969904359876000 + 100000000 * k for k in range(n_sweeps)]
for k in range(n_sweeps):
    sw = sweeps[k % 16]
    pd.DataFrame({"x": sw[:, 0].astype(np.float16), "y": sw[:, 1].astype(np.float16), "z": sw[:, 2].astype(np.float16),
                  "intensity": sw[:, 3].astype(np.uint8), "laser_number": np.zeros(len(sw), np.uint8),
                  "offset_ns": np.zeros(len(sw), np.int32)}).to_feather(log / "sensors" / "lidar" / f"{ts[k]}.feather")
pd.DataFrame({"timestamp_ns": ts, "qw": poses[:, 6], "qx": poses[:, 3], "qy": poses[:, 4], "qz": poses[:, 5],
              "tx_m": poses[:, 0], "ty_m": poses[:, 1], "tz_m": poses[:, 2]}).to_feather(log / "city_SE3_egovehicle.feather")
md = bg.synthetic.synthetic_map()
(log / "map" / "log_map_archive_synthlog00.json").write_text(json.dumps(md))
pd.DataFrame({"timestamp_ns": ts * 20, "category": "REGULAR_VEHICLE", "pad": "x" * 40}).to_feather(log / "annotations.feather")
disk = sum(p.stat().st_size for p in (log / "sensors" / "lidar").iterdir()) / 1e6
print(f"synthetic AV2 log: {n_sweeps} sweeps x {n_points} points, {disk:.0f} MB of feather files, {n_sweeps - 9} samples")

# (a) the reference's call sequence from disk, a few samples (dataset.py:285-348 through oracle/numpy_port.py)
from oracle import numpy_port as npp
seqs = bg.av2_dataset.make_sequences(bg.av2_dataset.find_scenarios(root))
pose_df = pd.read_feather(log / "city_SE3_egovehicle.feather")
t0 = time.perf_counter()
n_ref = 3
for s in seqs[:n_ref]:
    cur = pose_df[pose_df["timestamp_ns"] == s["current_ts_ns"]].iloc[0]
    cur7 = np.array([cur[k] for k in ["tx_m", "ty_m", "tz_m", "qx", "qy", "qz", "qw"]])
    pl, il = [], []
    for t in s["sweep_ts_list"]:
        df = pd.read_feather(log / "sensors" / "lidar" / f"{t}.feather", columns=["x", "y", "z", "intensity"])
        row = pose_df[pose_df["timestamp_ns"] == t].iloc[0]
        p7 = np.array([row[k] for k in ["tx_m", "ty_m", "tz_m", "qx", "qy", "qz", "qw"]])
        pl.append(npp.transform_points(df[["x", "y", "z"]].values, npp.rel_transforms(cur7, [p7])[0]))
        il.append(df["intensity"].values.astype(np.float32))
    npp.lidar_bev(pl, il)
    npp.map_raster(s["map_json_path"], dict(zip(["tx_m", "ty_m", "tz_m", "qx", "qy", "qz", "qw"], cur7.tolist())))
ref_fps = n_ref / (time.perf_counter() - t0)
print(f"reference call sequence from disk (1 process): {ref_fps:6.2f} samples/s")

# (b) this repo, from the same directory
for thr in (0, threads):
    bg.sweep_io.clear_cache()
    ds = bg.Av2BevDataset(root, is_train=False, store_points=(n_sweeps + 2) * (n_points + 8))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for batch in ds.loader(batch_size=B, decode_threads=thr, decode_ahead=3):
        n += batch["lidar_bev"].shape[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"Av2BevDataset from disk, batch {B}, decode threads {thr}: {n / dt:8.1f} samples/s  ({n} samples, {ds.store.h2d_bytes / 1e6:.0f} MB uploaded, "
          f"{n / dt / ref_fps:6.0f} x the reference process)")
shutil.rmtree(root)
