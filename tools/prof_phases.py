"""Dev tool: per-phase cycle counts inside k_tile_aug (clock64 by one thread per CTA).  Needs a library built with
-DBEVGEN_AT_PROFILE=<thread id>:  nvcc <flags of _native.NVCC_FLAGS> -DBEVGEN_AT_PROFILE=0 -o /tmp/libprof.so csrc/*.cu
usage: python tools/prof_phases.py /tmp/libprof.so 0"""
import sys, ctypes as C
sys.path.insert(0, '.')
import torch, bevgen_b200 as bg
from bevgen_b200 import _native
L = C.CDLL(sys.argv[1])
import bevgen_b200._native as N
N._lib = None; N.LIB_PATH = __import__('pathlib').Path(sys.argv[1])
lib = N.lib()
B = 4
frames = [bg.synthetic.frame_c2(i) for i in range(2)]; frames = [frames[i % 2] for i in range(B)]
batch = bg.FrameBatch.pack(frames, 10); gen = bg.BevGenerator(None); dv = gen.upload(batch)
out = torch.empty((B, 290, 400, 720), dtype=torch.float32, device="cuda")
names = ["cp.async wait", "A", "barrier1", "prefetch issue", "B", "barrier2", "C", "barrier3"]
for name, p in {"rotate": bg.AugParams(angle_deg=9.0), "scale": bg.AugParams(scale=1.03), "rot+scale": bg.AugParams(angle_deg=9.0, scale=1.03)}.items():
    for _ in range(2):
        bg.augment.lidar_bev_augmented(dv["points4"], dv["sweep_offsets"], dv["rel_tf"], B, 10, [p] * B, max_sweep_points=batch.max_sweep_points, out=out)
    buf = (C.c_ulonglong * 16)()
    lib.bevgen_debug_at_profile(buf, 1)
    bg.augment.lidar_bev_augmented(dv["points4"], dv["sweep_offsets"], dv["rel_tf"], B, 10, [p] * B, max_sweep_points=batch.max_sweep_points, out=out)
    lib.bevgen_debug_at_profile(buf, 1)
    n = buf[8] * 29
    print(name, "cycles per channel iteration (thread %s):" % sys.argv[2], " ".join(f"{nm}={buf[k]/n:.0f}" for k, nm in enumerate(names)), "total=%.0f" % (sum(buf[:8]) / n))
    c = buf[8]
    print(f"   per CTA: set-up to 1st barrier (global loads) {buf[9]/c:.0f}, to 2nd (geometry, taps) {buf[10]/c:.0f}, to the loop (tables, clearing) {buf[11]/c:.0f}; "
          f"whole CTA {buf[12]/c:.0f} cycles -> set-up = {100*(buf[9]+buf[10]+buf[11])/buf[12]:.1f} %")
