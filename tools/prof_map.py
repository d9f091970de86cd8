"""Run the map rasterizer alone on a config-3 batch (for ncu captures): python tools/prof_map.py [B] [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bevgen_b200 as bg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
md = bg.synthetic.synthetic_map()
table = bg.MapTable([bg.maps.parse_map_dict(md)])
rng = np.random.default_rng(0)
Ts = np.stack([bg.pose.map_matrix_from_pose(1500 + rng.uniform(-5, 5), -700 + rng.uniform(-5, 5), np.array([0.0, 0.0, np.sin(0.3), np.cos(0.3)])) for _ in range(B)])
out = torch.empty((B, 9, 400, 720), dtype=torch.float32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    e0.record(); bg.rasterize_maps(table, [0] * B, Ts, out=out); e1.record(); torch.cuda.synchronize()
    print(f"map raster B {B} rep {r}: {e0.elapsed_time(e1) * 1e3:.1f} us, {int(out.sum().item())} pixels set")
