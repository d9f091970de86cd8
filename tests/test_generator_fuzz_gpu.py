"""GPU: seeded differential fuzz of the whole generator against the oracle.

Every case draws a batch the way a loader could meet it -- batch size, ragged sweeps (empty ones, fewer than ten, clustered
ones that overflow the tile bins), point format (float32 rows / fp16 rows / AV2 columns / sweeps named in the device-resident
store), scatter algorithm, HD map on or off, augmentation parameters from the reference's own draw (utils.py:398-494,
all classes incl. dropout), fused or dense augmentation route -- runs ``BevGenerator.generate`` and compares every frame,
bit for bit, with oracle generation + the oracle's flip -> rotate -> scale -> dropout (utils.py:500-517)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

S = 10


@pytest.fixture(scope="module")
def bg():
    import bevgen_b200
    return bevgen_b200


@pytest.fixture(scope="module")
def co():
    from oracle import c_oracle
    return c_oracle


@pytest.fixture(scope="module")
def world(bg, co):
    md = bg.synthetic.synthetic_map(seed=11, rows=7, cols=6, n_cross=4)
    return {"md": md, "table": bg.MapTable([bg.maps.parse_map_dict(md)]), "flat": co.flatten_map(md)}


def _oracle_aug(co, g, p):
    out = g
    if p.flip:
        out = co.flip_w(out)
    if p.angle_deg is not None:
        out = co.rotate(out, p.angle_deg)
    if p.scale is not None:
        out = co.scale(out, p.scale)
    out = np.ascontiguousarray(out).copy()
    if p.rects:
        co.dropout(out, p.rects)
    return out


def _draw_frame(bg, rng, seed, md):
    """a frame with 1..10 loaded sweeps of 0..3000 points each; coordinates and intensities in AV2's value domain (float16
    coordinates, integer intensities 0..255) so that every point format holds the same values"""
    ns = int(rng.integers(1, S + 1))
    f = bg.synthetic.frame_c3(seed, n=3_000, S=ns, map_data=md)
    keep, off = [], [0]
    for s in range(ns):
        sw = f.sweep(s)
        kind = rng.integers(0, 6)
        n = 0 if kind == 0 else int(rng.integers(1, 3_001))
        sw = sw[:n].copy()
        if kind == 1 and n:                                             # a cluster: hundreds of points in a few voxels
            sw[:, :2] = sw[0, :2] + rng.normal(0, 0.15, (n, 2)).astype(np.float32)
        keep.append(sw)
        off.append(off[-1] + n)
    f.points4 = np.concatenate(keep).astype(np.float32) if off[-1] else np.zeros((0, 4), np.float32)
    f.sweep_offsets = np.asarray(off, np.int32)
    return bg.synthetic.quantize_f16(f)


@pytest.mark.parametrize("seed", range(12))
def test_generator_fuzz(bg, co, world, seed):
    rng = np.random.default_rng(9_000 + seed)
    B = int(rng.integers(1, 6))
    frames = [_draw_frame(bg, rng, 100 * seed + b, world["md"]) for b in range(B)]
    fmt = ["f32", "f16", "av2", "store"][seed % 4]
    algo = [bg.ALGO_TILED, bg.ALGO_ATOMIC, bg.ALGO_MEMSET_ATOMIC, bg.ALGO_AUTO][int(rng.integers(0, 4))]
    with_map = bool(rng.integers(0, 2))
    fused = bool(rng.integers(0, 4))                                    # the dense augmentation route one case in four
    aug = None
    if rng.integers(0, 4):
        aug = []
        for b in range(B):
            random.seed(31 * seed + b)
            aug.append(bg.draw_aug_params(random))                      # the reference's draw order and ranges
        if seed % 3 == 0:                                               # make sure the rare classes show up
            aug[0] = bg.AugParams(flip=True, angle_deg=-14.5, scale=1.049, rects=[(5, 5, 50, 50), (350, 670, 50, 50)])
    ids = [0] * B if with_map else None
    gen = bg.BevGenerator(world["table"] if with_map else None, algo=algo, fused_augment=fused)
    if fmt == "store":
        store = bg.SweepStore(capacity_points=B * S * 3_008 + 64)
        sf = []
        for b, f in enumerate(frames):
            keys = []
            for s in range(f.num_sweeps):
                sw = f.sweep(s)
                if len(sw) == 0 and rng.integers(0, 2):
                    keys.append(None)                                   # a missing sweep: no entry at all
                    continue
                store.put((b, s), sw[:, 0].astype(np.float16), sw[:, 1].astype(np.float16), sw[:, 2].astype(np.float16),
                          sw[:, 3].astype(np.uint8))
                keys.append((b, s))
            sf.append(bg.StoreFrame(keys, f.sweep_poses, f.current_pose))
        batch = bg.StoreBatch.pack(store, sf, S, map_ids=ids)
    else:
        batch = bg.FrameBatch.pack(frames, S, map_ids=ids, points_dtype={"f32": np.float32, "f16": np.float16, "av2": "av2"}[fmt])
    out = gen.generate(batch, aug, with_map=with_map)
    got_l = out["lidar_bev"].cpu().numpy()
    got_m = out["map_bev"].cpu().numpy() if with_map else None
    tag = (seed, fmt, algo, with_map, fused)
    assert got_l.any() and (got_m is None or got_m.any()), tag       # (not a comparison of empty grids)
    for b, f in enumerate(frames):
        rel = np.tile(np.eye(4), (S, 1, 1))
        rel[:f.num_sweeps] = f.rel_tf
        offs = np.full(S + 1, f.sweep_offsets[-1], np.int64)
        offs[:f.num_sweeps + 1] = f.sweep_offsets
        want = co.lidar_bev_fused(np.asarray(f.points4, np.float32), offs, rel, S)
        if aug is not None:
            want = _oracle_aug(co, want, aug[b])
        bad = got_l[b] != want
        assert not bad.any(), (tag, b, aug and aug[b], int(bad.sum()), np.argwhere(bad)[:4].tolist())
        if with_map:
            wm = co.map_raster_flat(world["flat"], bg.pose.map_matrix_from_pose(f.current_pose[0], f.current_pose[1], f.current_pose[3:7]))
            if aug is not None:
                wm = _oracle_aug(co, wm, aug[b])
            bad = got_m[b] != wm
            assert not bad.any(), (tag, b, aug and aug[b], int(bad.sum()), np.argwhere(bad)[:4].tolist())
