"""Run one LiDAR algo (0 tiled, 1 atomic, 2 memset) a few times -- for ncu captures and fill-only experiments.
usage: python tools/prof_one.py algo batch [c2|c4|empty] [reps] [f32|av2]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bevgen_b200 as bg

algo = int(sys.argv[1]); B = int(sys.argv[2]); kind = sys.argv[3] if len(sys.argv) > 3 else "c2"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
fmt = sys.argv[5] if len(sys.argv) > 5 else "f32"          # f32 rows | av2 columns (float16 x,y,z + uint8 intensity)
S = 10
if kind == "empty":
    pts = np.zeros((0, 4), np.float32); off = np.zeros(B * S + 1, np.int32); tf = np.tile(np.eye(4), (B * S, 1, 1))
else:
    mk = {"c2": bg.synthetic.frame_c2, "c4": bg.synthetic.frame_c4}[kind]
    frames = [mk(i) for i in range(min(B, 2))]
    if fmt == "av2":
        frames = [bg.synthetic.quantize_f16(f) for f in frames]
    frames = [frames[i % len(frames)] for i in range(B)]
    pts = np.concatenate([f.points4 for f in frames])
    off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(f.sweep_offsets) for f in frames]))]).astype(np.int32)
    tf = np.concatenate([f.rel_tf for f in frames])
dev = torch.device("cuda")
d_pts = (bg.PointColumns.from_points4(pts) if fmt == "av2" else torch.from_numpy(pts)).to(dev); d_off = torch.from_numpy(off).to(dev); d_tf = torch.from_numpy(tf).to(dev)
out = torch.empty((B, 290, 400, 720), dtype=torch.float32, device=dev)
mx = int(np.diff(off).max())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    e0.record()
    bg.lidar_bev_from_sweeps(d_pts, d_off, d_tf, B, S, max_sweep_points=mx, out=out, algo=algo)
    e1.record(); torch.cuda.synchronize()
    print(f"algo {algo} B {B} {kind} rep {r}: {e0.elapsed_time(e1)*1e3:.1f} us  -> {B/e0.elapsed_time(e1)*1e3:.0f} fps")
