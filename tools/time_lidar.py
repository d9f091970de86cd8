"""Dev timing of the LiDAR path (not the bench contract): CUDA events, per-algo, vs a plain zero-fill."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bevgen_b200 as bg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kind = sys.argv[2] if len(sys.argv) > 2 else "c2"
S = 10
mk = {"c2": bg.synthetic.frame_c2, "c4": bg.synthetic.frame_c4}[kind]
frames = [mk(i) for i in range(min(B, 4))]
frames = [frames[i % len(frames)] for i in range(B)]
pts = np.concatenate([f.points4 for f in frames]); off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(f.sweep_offsets) for f in frames]))]).astype(np.int32)
tf = np.concatenate([f.rel_tf for f in frames])
dev = torch.device("cuda")
d_pts = torch.from_numpy(pts).to(dev); d_off = torch.from_numpy(off).to(dev); d_tf = torch.from_numpy(tf).to(dev)
out = torch.empty((B, 290, 400, 720), dtype=torch.float32, device=dev)
mx = int(np.diff(off).max())

def timeit(fn, n=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
res["zero_fill_ms"] = timeit(lambda: out.zero_())
for algo, name in [(0, "tiled"), (1, "atomic"), (2, "memset")]:
    res[name + "_ms"] = timeit(lambda: bg.lidar_bev_from_sweeps(d_pts, d_off, d_tf, B, S, max_sweep_points=mx, out=out, algo=algo))
bytes_frame = 290 * 400 * 720 * 4 + 16 * 1_000_000 + 1280
for k in list(res):
    ms = res[k]
    res[k.replace("_ms", "_fps")] = B / ms * 1e3
    res[k.replace("_ms", "_GBps")] = B * bytes_frame / ms / 1e6
res["B"] = B; res["kind"] = kind
print(json.dumps(res))
