"""Dev timing of generation fused with augmentation (k_tile_aug) per class vs the dense route, CUDA events."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bevgen_b200 as bg
from bevgen_b200 import _native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
only = sys.argv[2] if len(sys.argv) > 2 else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
kind = sys.argv[4] if len(sys.argv) > 4 else "c2"
mk = {"c2": bg.synthetic.frame_c2, "c4": bg.synthetic.frame_c4, "empty": lambda i: bg.synthetic.frame_c2(i, n=1)}[kind]
frames = [mk(i) for i in range(2)]
frames = [frames[i % 2] for i in range(B)]
batch = bg.FrameBatch.pack(frames, 10)
gen = bg.BevGenerator(None)
dv = gen.upload(batch)
out = torch.empty((B, 290, 400, 720), dtype=torch.float32, device="cuda")
classes = {"plain": bg.AugParams(), "rotate": bg.AugParams(angle_deg=9.0), "rotate15": bg.AugParams(angle_deg=-15.0), "scale_up": bg.AugParams(scale=1.03),
           "scale_down": bg.AugParams(scale=0.97), "rot+scale": bg.AugParams(angle_deg=9.0, scale=1.03, flip=True, rects=[(10, 10, 40, 40)])}
for name, p in classes.items():
    if only and name != only:
        continue
    ps = [p] * B
    for _ in range(2):
        bg.augment.lidar_bev_augmented(dv["points4"], dv["sweep_offsets"], dv["rel_tf"], B, 10, ps, max_sweep_points=batch.max_sweep_points, out=out)
    torch.cuda.synchronize()
    _native.timing_enable(True); _native.timing_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        bg.augment.lidar_bev_augmented(dv["points4"], dv["sweep_offsets"], dv["rel_tf"], B, 10, ps, max_sweep_points=batch.max_sweep_points, out=out)
    e1.record(); torch.cuda.synchronize()
    kt = _native.timing_collect(); _native.timing_enable(False)
    us = e0.elapsed_time(e1) / reps / B * 1e3
    ks = "  ".join(f"{k} {ms / reps / B * 1e3:.1f}" for k, (c, ms) in kt.items())
    print(f"{name:12s} {us:8.1f} us/frame  {290 * 400 * 720 * 4 / us / 1e6:6.2f} TB/s written   [{ks}]", flush=True)
