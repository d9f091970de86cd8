"""Tiny end-to-end run of every kernel (for compute-sanitizer): LiDAR x3 algos, fp16 input, map, augmentation."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bevgen_b200 as bg

md = bg.synthetic.synthetic_map(seed=2, rows=6, cols=6, n_cross=3)
table = bg.MapTable([bg.maps.parse_map_dict(md)])
frames = [bg.synthetic.quantize_f16(bg.synthetic.frame_c3(i, n=700, S=10, map_data=md)) for i in range(2)]
frames.append(bg.synthetic.frame_c4(0, n=900))
frames[2].map_data = md
aug = [bg.AugParams(flip=True, rects=[(5, 5, 30, 30)]), bg.AugParams(angle_deg=6.0, scale=0.97), bg.AugParams(scale=1.04)]
for algo in (0, 1, 2):
    gen = bg.BevGenerator(table, algo=algo)
    out = gen.generate(bg.FrameBatch.pack(frames, 10, map_ids=[0, 0, 0]), aug)
    torch.cuda.synchronize()
    print(algo, float(out["lidar_bev"].sum()), float(out["map_bev"].sum()))
fb16 = bg.FrameBatch.pack(frames[:2], 10, map_ids=[0, 0], points_dtype=np.float16)
print("f16", float(bg.BevGenerator(table).generate(fb16)["lidar_bev"].sum()))
pts = bg.transform_points(frames[0].sweep(0)[:, :3], frames[0].rel_tf[0])
print("tp", pts.shape, bg.world_to_bev_pixel(np.array([[1500.0, -700.0]]), np.eye(3)).tolist())
print("ok")
