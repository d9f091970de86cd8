"""Dev timing: per-kernel us/frame of the default LiDAR path as a function of the batch size (config-2 frames, AV2 columns).
Run once per setting of BEVGEN_L2_HINTS (unset = automatic: on for launches of <= 160 sweeps, 0 = off, 1 = on) to see what
the L2 residency hints of the record bins are worth:   BEVGEN_L2_HINTS=0 python tools/time_batch_sizes.py"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bevgen_b200 as bg
from bevgen_b200 import _native

print("BEVGEN_L2_HINTS =", os.environ.get("BEVGEN_L2_HINTS", "(auto)"))
for B in [int(a) for a in sys.argv[1:]] or (1, 2, 4, 8, 16, 32):
    frames = [bg.synthetic.quantize_f16(bg.synthetic.frame_c2(i)) for i in range(2)]
    frames = [frames[i % 2] for i in range(B)]
    batch = bg.FrameBatch.pack(frames, 10, points_dtype="av2")
    gen = bg.BevGenerator(None)
    dv = gen.upload(batch)
    out = torch.empty((B, 290, 400, 720), dtype=torch.float32, device="cuda")
    for _ in range(3):
        gen.generate(batch, dev=dv, with_map=False, out_lidar=out)
    torch.cuda.synchronize()
    _native.timing_enable(True); _native.timing_collect()
    reps = max(4, 64 // B)
    for _ in range(reps):
        gen.generate(batch, dev=dv, with_map=False, out_lidar=out)
    torch.cuda.synchronize()
    kt = _native.timing_collect(); _native.timing_enable(False)
    ks = "  ".join(f"{k} {ms / reps / B * 1e3:.2f}" for k, (c, ms) in kt.items())
    tot = sum(ms for c, ms in kt.values()) / reps / B * 1e3
    print(f"B {B:2d}: {tot:6.2f} us/frame  {1e6 / tot:7.0f} frames/s  [{ks}]", flush=True)
