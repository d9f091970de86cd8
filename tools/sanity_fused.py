"""Tiny fused generation+augmentation run for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool racecheck python tools/sanity_fused.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bevgen_b200 as bg

S = 1
frames = [bg.synthetic.frame_c2(1, n=4_000, S=S), bg.synthetic.frame_c4(1, n=6_000, S=S)]     # sparse + clustered (slow path)
fb = bg.FrameBatch.pack(frames * 2, S)
dev = torch.device("cuda")
pts, off, tf = fb.points4.to(dev), fb.sweep_offsets.to(dev), fb.rel_tf.to(dev)
ps = [bg.AugParams(angle_deg=9.0, flip=True), bg.AugParams(scale=1.04, rects=[(10, 10, 30, 30)]),
      bg.AugParams(angle_deg=-12.0, scale=0.96), bg.AugParams(angle_deg=5.0, scale=1.02)]
out = bg.augment.lidar_bev_augmented(pts, off, tf, 4, S, ps, max_sweep_points=fb.max_sweep_points)
plain = bg.lidar_bev_from_sweeps(pts, off, tf, 4, S, max_sweep_points=fb.max_sweep_points)
want = bg.augment_grid(plain, ps)
torch.cuda.synchronize()
print("fused == dense:", bool(torch.equal(out, want)), "nonzeros", int((out != 0).sum()))
md = bg.synthetic.synthetic_map(seed=5, rows=4, cols=4, n_cross=3)
table = bg.MapTable([bg.maps.parse_map_dict(md)])
T3 = np.stack([bg.pose.map_matrix_from_pose(1500.0, -700.0, np.array([0, 0, 0.3, 0.95])) for _ in range(2)])
m = bg.rasterize_maps(table, [0, 0], T3)
torch.cuda.synchronize()
print("map nonzeros", int((m != 0).sum()))
