"""GPU: the device-resident sweep store (sliding-window streams: every sweep uploaded once, named by many frames) produces
the same bits as packing every frame's own ten sweeps, and as the oracle."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bg():
    import bevgen_b200
    return bevgen_b200


@pytest.fixture(scope="module")
def co():
    from oracle import c_oracle
    return c_oracle


def _put(bg, store, key, sw):
    store.put(key, sw[:, 0].astype(np.float16), sw[:, 1].astype(np.float16), sw[:, 2].astype(np.float16), sw[:, 3].astype(np.uint8))


def _store_frames(bg, store, sweeps, poses, firsts, S=10, log="L"):
    out = []
    for i in firsts:
        for k in range(i, i + S):
            _put(bg, store, (log, k), sweeps[k])
        out.append(bg.StoreFrame([(log, k) for k in range(i, i + S)], poses[i:i + S].copy(), poses[i + S - 1].copy()))
    return out


@pytest.mark.parametrize("algo", [0, 1])
def test_sliding_windows_from_store_equal_packed_frames_and_oracle(bg, co, algo):
    S = 10
    sweeps, poses = bg.synthetic.sliding_log(16, n=6_001, seed=1)        # odd length: unaligned column starts in the packed form
    firsts = [0, 1, 2, 5, 6]                                             # frames share up to 9 sweeps
    frames = [bg.synthetic.window_frame(sweeps, poses, i) for i in firsts]
    want = np.stack([co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, S) for f in frames])
    gen = bg.BevGenerator(algo=algo)
    packed = gen.generate(bg.FrameBatch.pack(frames, S, points_dtype="av2"), with_map=False)["lidar_bev"].clone()
    store = bg.SweepStore(capacity_points=16 * 6_008)
    sf = _store_frames(bg, store, sweeps, poses, firsts)
    assert len(store) == 16 and store.h2d_bytes == 16 * 6_001 * 7        # every sweep uploaded once
    batch = bg.StoreBatch.pack(store, sf, S)
    assert batch.h2d_bytes < 20_000 and batch.total_points == 5 * 10 * 6_001
    got = gen.generate(batch, with_map=False)["lidar_bev"]
    assert torch.equal(got, packed)
    assert np.array_equal(got.cpu().numpy(), want)
    # a frame with a missing sweep (None key) and one with fewer sweeps than slots
    sf2 = [bg.StoreFrame([("L", k) if k != 3 else None for k in range(10)], poses[:10].copy(), poses[9].copy()),
           bg.StoreFrame([("L", k) for k in range(4, 8)], poses[4:8].copy(), poses[7].copy())]
    got2 = gen.generate(bg.StoreBatch.pack(store, sf2, S), with_map=False)["lidar_bev"].cpu().numpy()
    f0 = bg.synthetic.window_frame(sweeps, poses, 0)
    w0 = co.lidar_bev_fused(f0.points4, f0.sweep_offsets, f0.rel_tf, S)
    w0[3 * 29:4 * 29] = 0
    assert np.array_equal(got2[0], w0)
    f1 = bg.synthetic.window_frame(sweeps, poses, 4, S=4)
    assert np.array_equal(got2[1], co.lidar_bev_fused(f1.points4, f1.sweep_offsets, f1.rel_tf, S))


def test_store_with_augmentation_and_map(bg, co):
    S = 10
    md = bg.synthetic.synthetic_map(seed=5, rows=6, cols=5, n_cross=3)
    table = bg.MapTable([bg.maps.parse_map_dict(md)])
    sweeps, poses = bg.synthetic.sliding_log(13, n=5_000, seed=2)
    poses[:, 0] += 1500.0; poses[:, 1] -= 700.0                          # onto the synthetic map
    firsts = [0, 1, 2, 3]
    frames = [bg.synthetic.window_frame(sweeps, poses, i) for i in firsts]
    aug = [bg.AugParams(angle_deg=6.0), bg.AugParams(flip=True, scale=1.04), bg.AugParams(angle_deg=-12.0, scale=0.96, rects=[(30, 40, 25, 35)]),
           bg.AugParams(flip=True, rects=[(200, 100, 40, 20)])]
    store = bg.SweepStore(capacity_points=13 * 5_000)
    sf = _store_frames(bg, store, sweeps, poses, firsts)
    for fused in (True, False):                                          # fused tile kernel and the dense route (sub-ranges)
        gen = bg.BevGenerator(table, fused_augment=fused)
        a = gen.generate(bg.FrameBatch.pack(frames, S, map_ids=[0] * 4, points_dtype="av2"), aug)
        la, ma = a["lidar_bev"].clone(), a["map_bev"].clone()
        b = gen.generate(bg.StoreBatch.pack(store, sf, S, map_ids=[0] * 4), aug)
        assert torch.equal(b["lidar_bev"], la) and torch.equal(b["map_bev"], ma), fused


def test_shared_dense_sweeps_keep_their_overflow_lists_apart(bg, co):
    """clustered sweeps overflow the tile bins; two frames that name the SAME sweeps must not share overflow-list slices"""
    S = 10
    sweeps, poses = bg.synthetic.sliding_log(11, n=20_000, seed=3, dense=True)
    frames = [bg.synthetic.window_frame(sweeps, poses, i) for i in (0, 1)]
    want = np.stack([co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, S) for f in frames])
    store = bg.SweepStore(capacity_points=11 * 20_000)
    sf = _store_frames(bg, store, sweeps, poses, (0, 1))
    gen = bg.BevGenerator()
    got = gen.generate(bg.StoreBatch.pack(store, sf, S), with_map=False)["lidar_bev"].cpu().numpy()
    assert np.array_equal(got, want)
    aug = [bg.AugParams(angle_deg=5.0, scale=1.02), bg.AugParams(scale=0.97)]
    ga = gen.generate(bg.StoreBatch.pack(store, sf, S), aug, with_map=False)["lidar_bev"].cpu().numpy()
    for k in range(2):
        w = want[k]
        if aug[k].angle_deg is not None: w = co.rotate(w, aug[k].angle_deg)
        w = co.scale(w, aug[k].scale)
        assert np.array_equal(ga[k], np.ascontiguousarray(w)), k


def test_ring_eviction_and_wraparound(bg, co):
    """a store smaller than the log: sweeps are evicted first-in-first-out as the window slides; results stay exact, and a
    batch that names an evicted sweep fails in pack()"""
    S = 10
    n = 3_000
    sweeps, poses = bg.synthetic.sliding_log(40, n=n, seed=4)
    store = bg.SweepStore(capacity_points=14 * 3_000 + 100)              # 14 sweeps fit
    gen = bg.BevGenerator()
    for i in range(0, 30, 3):                                            # windows i .. i+9, stride 3
        sf = _store_frames(bg, store, sweeps, poses, [i])
        got = gen.generate(bg.StoreBatch.pack(store, sf, S), with_map=False)["lidar_bev"][0].cpu().numpy()
        f = bg.synthetic.window_frame(sweeps, poses, i)
        assert np.array_equal(got, co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, S)), i
    assert store.evictions > 0 and len(store) <= 14
    assert ("L", 0) not in store
    with pytest.raises(KeyError):
        bg.StoreBatch.pack(store, [bg.StoreFrame([("L", 0)], poses[:1].copy(), poses[0].copy())], S)
    with pytest.raises(ValueError):
        store.put("f32", np.zeros(4, np.float32), np.zeros(4, np.float16), np.zeros(4, np.float16), np.zeros(4, np.uint8))
    with pytest.raises(ValueError):
        store.put("too big", *(np.zeros(10**6, np.float16),) * 3, np.zeros(10**6, np.uint8))


def test_batch_generated_again_after_eviction_is_an_error(bg, co):
    """a packed batch is pinned until its first generate; a second generate of the same batch is fine while its sweeps are
    still resident and a KeyError -- not a grid from other sweeps' points -- once an upload has taken their rows"""
    S, n = 10, 2_000
    sweeps, poses = bg.synthetic.sliding_log(30, n=n, seed=9)
    store = bg.SweepStore(capacity_points=12 * n)
    gen = bg.BevGenerator()
    batch = bg.StoreBatch.pack(store, _store_frames(bg, store, sweeps, poses, [0]), S)
    a = gen.generate(batch, with_map=False)["lidar_bev"].clone()
    b = gen.generate(batch, with_map=False)["lidar_bev"]                # still resident: same bits
    assert torch.equal(a, b)
    _store_frames(bg, store, sweeps, poses, [12])                        # uploads sweeps 12..21: evicts 0..9
    assert ("L", 0) not in store
    with pytest.raises(KeyError):
        gen.generate(batch, with_map=False)


def test_batch_loader_with_store_equals_loader_without(bg):
    """BevBatchLoader(sweep_store=...): sliding-window samples that carry sweep_keys share their sweeps through the store;
    same batches (train-time augmentation included) as packing every sample's own ten sweeps"""
    S = 10
    sweeps, poses = bg.synthetic.sliding_log(15, n=3_001, seed=9)
    samples = []
    for i in range(5):
        f = bg.synthetic.window_frame(sweeps, poses, i)
        f.sweep_keys = [("log", i + k) for k in range(S)]
        samples.append(f)
    store = bg.SweepStore(capacity_points=15 * 3_008)
    for train in (False, True):
        random.seed(11)
        a = list(bg.dataset_shim.BevBatchLoader(samples, 2, is_train=train))
        random.seed(11)
        b = list(bg.dataset_shim.BevBatchLoader(samples, 2, is_train=train, sweep_store=store))
        assert len(a) == len(b) == 3
        for x, y in zip(a, b):
            assert torch.equal(x["lidar_bev"], y["lidar_bev"])
    assert len(store) == 14 and store.h2d_bytes == 14 * 3_001 * 7        # 5 windows over 14 distinct sweeps, each sent once


def test_prefetching_generator_accepts_store_batches(bg, co):
    """PrefetchingGenerator.submit / run with StoreBatches: only the per-frame tables travel on the copy stream, the points
    are the store's resident columns; results as from a direct generate"""
    S = 10
    sweeps, poses = bg.synthetic.sliding_log(14, n=2_500, seed=12)
    store = bg.SweepStore(capacity_points=14 * 2_504)
    gen = bg.BevGenerator()
    pf = bg.PrefetchingGenerator(gen)
    handles, wants = [], []
    for i in (0, 2, 4):
        sf = _store_frames(bg, store, sweeps, poses, [i])
        handles.append(pf.submit(bg.StoreBatch.pack(store, sf, S)))
        f = bg.synthetic.window_frame(sweeps, poses, i)
        wants.append(co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, S))
    for h, w in zip(handles, wants):
        out = pf.run(h, with_map=False)["lidar_bev"][0]
        assert np.array_equal(out.cpu().numpy(), w)
    assert not store._pins                                             # every batch was consumed: nothing stays pinned
