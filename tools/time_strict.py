"""Timing of the reference-signature path (numpy in -> fresh numpy out, one frame per call), as dataset.py:340-347 uses it."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bevgen_b200 as bg

f0 = bg.synthetic.frame_c2(0)
S = 10
il = [f0.sweep(s)[:, 3].copy() for s in range(S)]
# the arrays dataset.py:320 hands over: sweep_df[['x','y','z']].values -- float16, column-major (a pandas block transposed)
f0 = bg.synthetic.quantize_f16(f0)
raw = [np.asfortranarray(f0.sweep(s)[:, :3].astype(np.float16)) for s in range(S)]
def one():
    pl = [bg.transform_points(raw[s], f0.rel_tf[s]) for s in range(S)]
    return bg.create_intentnet_lidar_bev(pl, il)
def one_installed():
    """the same calls as dataset_shim.install() binds them in the reference's dataset module (transform deferred)"""
    tp = bg.bev_ops.transform_points_deferred
    return bg.create_intentnet_lidar_bev([tp(raw[s], f0.rel_tf[s]) for s in range(S)], il)
one(); one()
t0 = time.perf_counter()
keep = []
for i in range(8):
    keep.append(one())          # a batch of 8 results alive at once, as a DataLoader would hold them
dt = time.perf_counter() - t0
print(f"strict signatures: {8 / dt:.1f} frames/s ({dt / 8 * 1e3:.1f} ms per frame), 8 results alive")
del keep
t0 = time.perf_counter()
for i in range(8):
    r = one()
dt = time.perf_counter() - t0
print(f"strict signatures: {8 / dt:.1f} frames/s ({dt / 8 * 1e3:.1f} ms per frame), steady state (buffers recycled)")
one_installed(); one_installed()
t0 = time.perf_counter()
for i in range(8):
    r = one_installed()
dt = time.perf_counter() - t0
print(f"install() mode (deferred transform, fused launch): {8 / dt:.1f} frames/s ({dt / 8 * 1e3:.1f} ms per frame), steady state")
