"""Where the reference-signature path (tools/time_strict.py) spends a frame: per call, wall clock with a device synchronize."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bevgen_b200 as bg

f0 = bg.synthetic.frame_c2(0)
S = 10
il = [f0.sweep(s)[:, 3].copy() for s in range(S)]
sw = [f0.sweep(s)[:, :3] for s in range(S)]


def timed(fn, n=8):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


t_tp, _ = timed(lambda: [bg.transform_points(sw[s], f0.rel_tf[s]) for s in range(S)])
pl = [bg.transform_points(sw[s], f0.rel_tf[s]) for s in range(S)]
t_cr, _ = timed(lambda: bg.create_intentnet_lidar_bev(pl, il))
t_pack, _ = timed(lambda: bg.bev_ops._pack_sweeps(pl, il, S))
print(f"transform_points x {S}: {t_tp:.2f} ms   create_intentnet_lidar_bev: {t_cr:.2f} ms (of which host packing {t_pack:.2f} ms)")
