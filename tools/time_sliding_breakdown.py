"""Where a step of the sliding-window stream (bench.py's e2e.sliding_window) spends its time: host seconds in store.put /
StoreBatch.pack / generate, the kernels' own time (library event timing), and the step's wall clock.
usage: python tools/time_sliding_breakdown.py [batch=32] [steps=20] [store_steps=3]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bevgen_b200 as bg
from bevgen_b200 import _native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
store_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
S, n = 10, 100_000
dev = torch.device("cuda", 0)
pool_n = 64
sweeps, _ = bg.synthetic.sliding_log(pool_n, n=n, seed=7)
pool = [bg.PointColumns.from_points4(p) for p in sweeps]
def col(pc, k):
    b, st, m = pc.buf, pc.stride, pc.n
    return b[2 * k * st:2 * k * st + 2 * m].view(torch.float16) if k < 3 else b[6 * st:6 * st + m]
pool_cols = [tuple(col(pc, k) for k in range(4)) for pc in pool]
total = steps * 2 + 8
log_poses = bg.synthetic._poses(total * B + S)
store = bg.SweepStore(capacity_points=(store_steps * B + S) * (n + 8), device=dev)
def sample(i):
    cp = log_poses[i + S - 1]
    return bg.StoreFrame(list(range(i, i + S)), log_poses[i:i + S], cp, rel_tf=bg.pose.sweep_rel_transforms(cp, log_poses[i:i + S]),
                         ego_T=bg.pose.map_matrix_from_pose(cp[0], cp[1], cp[3:7]))
samples = [sample(i) for i in range(total * B)]
gen = bg.BevGenerator(None, algo=bg.ALGO_AUTO)
outs = [torch.empty((B, 290, 400, 720), dtype=torch.float32, device=dev) for _ in range(2)]
state = {"next": 0, "step": 0}
host = {"put": 0.0, "pack": 0.0, "generate": 0.0}
def step(o):
    i0 = state["step"] * B
    t0 = time.perf_counter()
    while state["next"] < i0 + B + S - 1:
        store.put(state["next"], *pool_cols[state["next"] % pool_n]); state["next"] += 1
    t1 = time.perf_counter()
    sb = bg.StoreBatch.pack(store, samples[i0:i0 + B], S)
    t2 = time.perf_counter()
    gen.generate(sb, None, with_map=False, out_lidar=o)
    t3 = time.perf_counter()
    host["put"] += t1 - t0; host["pack"] += t2 - t1; host["generate"] += t3 - t2
    state["step"] += 1
def run(k, lookahead=2):
    done = []
    for i in range(k):
        if len(done) >= lookahead:
            done.pop(0).synchronize()
        step(outs[i % 2])
        ev = torch.cuda.Event(); ev.record(); done.append(ev)
    torch.cuda.synchronize()
run(5)
for k in host: host[k] = 0.0
t0 = time.perf_counter(); run(steps); wall = (time.perf_counter() - t0) / steps * 1e3
print(f"B {B}, store of {store_steps} steps: {wall:.3f} ms/step wall -> {B / wall * 1e3:.0f} frames/s;  host per step: " +
      "  ".join(f"{k} {v / steps * 1e3:.3f} ms" for k, v in host.items()))
_native.timing_enable(True); _native.timing_collect()
for k in host: host[k] = 0.0
run(steps)
kt = _native.timing_collect(); _native.timing_enable(False)
print("kernels per step (event-timed): " + "  ".join(f"{k} {ms / steps:.4f}" for k, (c, ms) in kt.items()) +
      f"  sum {sum(ms for c, ms in kt.values()) / steps:.4f} ms")
