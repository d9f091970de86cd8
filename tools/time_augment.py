"""Dev timing of the augmentation kernel per class (CUDA events)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bevgen_b200 as bg

B, C = 4, 290
src = torch.zeros((B, C, 400, 720), device="cuda")
src.view(-1)[torch.randint(0, src.numel(), (2_000_000,), device="cuda")] = 100.0
out = torch.empty_like(src)
classes = {"identity": bg.AugParams(), "flip": bg.AugParams(flip=True), "rotate": bg.AugParams(angle_deg=9.0),
           "scale_up": bg.AugParams(scale=1.03), "scale_down": bg.AugParams(scale=0.97),
           "rot+scale (2 pass)": bg.AugParams(angle_deg=9.0, scale=1.03), "rot+scale (fused)": bg.AugParams(angle_deg=9.0, scale=1.03)}
for name, p in classes.items():
    two = "fused" not in name
    for _ in range(2):
        bg.augment_grid(src, [p] * B, out=out, two_pass=two)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        bg.augment_grid(src, [p] * B, out=out, two_pass=two)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 / B * 1e3
    print(f"{name:22s} {us:8.1f} us/frame   {2 * C * 400 * 720 * 4 / us / 1e6:6.2f} TB/s (1R+1W algorithmic)")
