"""Dev timing: pipelined end-to-end loop of config 5 (pinned host frames -> H2D -> generation + augmentation), host-side
breakdown.  usage: python tools/time_e2e_c5.py [order]   order = submit_first (bench.py) | run_first"""
import random, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bevgen_b200 as bg

order = sys.argv[1] if len(sys.argv) > 1 else "submit_first"
B, S, steps = 16, 10, 20
md = bg.synthetic.synthetic_map()
frames = [bg.synthetic.frame_c3(i, map_data=md) for i in range(2)]; frames = [frames[i % 2] for i in range(B)]
aug = []
for i in range(B):
    random.seed(5000 + i); aug.append(bg.draw_aug_params(random))
frames, aug, _ = bg.partition_for_augmentation(frames, aug)
table = bg.MapTable([bg.maps.parse_map_dict(md)])
batch = bg.FrameBatch.pack(frames, S, map_ids=[0] * B)
gen = bg.BevGenerator(table); pf = bg.PrefetchingGenerator(gen)
outs = [(torch.empty((B, 290, 400, 720), device="cuda"), torch.empty((B, 9, 400, 720), device="cuda")) for _ in range(2)]
probe = torch.empty((B, 720)).pin_memory()
def loop(n):
    host_run = host_submit = 0.0
    h = pf.submit(batch)
    for i in range(n):
        ol, om = outs[i % 2]
        if order == "submit_first":
            t0 = time.perf_counter(); cur, h = h, (pf.submit(batch) if i + 1 < n else None); t1 = time.perf_counter()
            r = pf.run(cur, aug, out_lidar=ol, out_map=om); t2 = time.perf_counter()
            host_submit += t1 - t0; host_run += t2 - t1
        else:
            t0 = time.perf_counter(); r = pf.run(h, aug, out_lidar=ol, out_map=om); t1 = time.perf_counter()
            h = pf.submit(batch) if i + 1 < n else None; t2 = time.perf_counter()
            host_run += t1 - t0; host_submit += t2 - t1
        probe.copy_(r["lidar_bev"][:, 14, 200, :], non_blocking=True)
    return host_run / n, host_submit / n
loop(3); torch.cuda.synchronize()
t0 = time.perf_counter(); hr, hs = loop(steps); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{order}: {B * steps / dt:.0f} frames/s, {dt / steps * 1e3:.2f} ms/step; host time in run() {hr * 1e3:.2f} ms, in submit() {hs * 1e3:.2f} ms; H2D {batch.h2d_bytes / 1e6:.0f} MB/step")
