"""CPU: host-side logic of the product (no kernels): map pre-parse, batch packing, parameter packing, sharding."""
import ctypes
import random

import numpy as np
import pytest


@pytest.fixture(scope="module")
def bg():
    import bevgen_b200
    return bevgen_b200


def test_flat_map_skips_unusable_entries(bg):
    md = {"lane_segments": {
        "a": {"left_lane_boundary": [{"x": 1.0, "y": 2.0}, "junk", {"x": 3.0}, {"x": 4.0, "y": 5.0, "z": 9}],
              "right_lane_boundary": [], "is_intersection": True, "lane_type": "BUS",
              "left_lane_mark_type": "SOLID_YELLOW", "right_lane_mark_type": "NONE"},
        "b": {"left_lane_boundary": [None], "right_lane_boundary": [{"x": 0.0, "y": 0.0}, {"x": 1.0, "y": 1.0}]}},
        "pedestrian_crossings": {"c": {"polygon": [{"x": 0, "y": 0}, {"x": 1, "y": 0}, {"x": 1, "y": 1}]}, "d": {"edge1": []}}}
    fm = bg.maps.parse_map_dict(md)
    assert fm.verts.tolist() == [[1, 2], [4, 5], [0, 0], [1, 1], [0, 0], [1, 0], [1, 1]]
    assert fm.pl_off.tolist() == [0, 2, 4, 7]
    assert fm.lanes.tolist() == [[0, -1, 1 | 2 | (8 << 4), 0], [-1, 1, 0, 0]]
    assert fm.cws.tolist() == [2]
    assert fm.num_prims == 3 and fm.max_polyline_len() == 3
    assert bg.maps.parse_map_dict({}).num_prims == 0
    too_long = {"lane_segments": {"a": {"left_lane_boundary": [{"x": i, "y": 0} for i in range(1025)]}}}
    with pytest.raises(ValueError):
        bg.maps.parse_map_dict(too_long)


def test_frame_batch_pack(bg):
    frames = [bg.synthetic.frame_c2(0, n=50, S=10), bg.synthetic.frame_c1(0, n=33), bg.synthetic.frame_c2(1, n=7, S=4)]
    fb = bg.FrameBatch.pack(frames, 10)
    off = fb.sweep_offsets.numpy()
    assert fb.num_frames == 3 and fb.num_sweeps == 10 and off.shape == (31,)
    assert fb.num_points == 500 + 33 + 28 and fb.max_sweep_points == 50
    assert np.diff(off).tolist() == [50] * 10 + [33] + [0] * 9 + [7] * 4 + [0] * 6
    rel = fb.rel_tf.numpy().reshape(3, 10, 4, 4)
    assert np.array_equal(rel[0], frames[0].rel_tf) and np.array_equal(rel[2, :4], frames[2].rel_tf)
    assert np.array_equal(rel[1, 5], np.eye(4))
    with pytest.raises(IndexError):
        bg.FrameBatch.pack(frames, 3)


def test_aug_param_struct_layout(bg):
    from bevgen_b200 import _native
    assert ctypes.sizeof(_native.AugParams) == 152          # 6*8 + 8 + 4*4 + 20*4 (include/bevgen.h)
    p = bg.AugParams(flip=True, angle_deg=5.0, scale=0.97, rects=[(1, 2, 20, 30), (3, 4, 25, 35)])
    raw = bg.augment.pack_params([bg.AugParams(), p])
    assert raw.shape == (304,)
    s = _native.AugParams.from_buffer_copy(raw[152:].tobytes())
    assert (s.do_flip, s.do_rotate, s.do_scale, s.n_rects) == (1, 1, 1, 2)
    assert list(s.rects[:8]) == [1, 2, 20, 30, 3, 4, 25, 35] and s.scale == 0.97
    mi = bg.augment.invert_affine_2x3(bg.augment.rotation_matrix_2d(5.0)).reshape(6)
    assert list(s.minv) == mi.tolist()
    z = _native.AugParams.from_buffer_copy(raw[:152].tobytes())
    assert (z.do_flip, z.do_rotate, z.do_scale, z.n_rects) == (0, 0, 0, 0)
    flip, rects, cnt = bg.augment.fusion_arrays([bg.AugParams(flip=True, rects=[(5, 6, 20, 20)]), p])
    assert flip.tolist() == [1, 0] and cnt.tolist() == [1, 0] and rects[0, 0].tolist() == [5, 6, 20, 20]


def test_draw_aug_params_ranges(bg):
    random.seed(99)
    seen = set()
    for _ in range(400):
        p = bg.draw_aug_params(random)
        seen.add((p.flip, p.angle_deg is not None, p.scale is not None, bool(p.rects)))
        assert p.angle_deg is None or -15.0 <= p.angle_deg <= 15.0
        assert p.scale is None or 0.95 <= p.scale <= 1.05
        assert len(p.rects) <= 5
        for y0, x0, ph, pw in p.rects:
            assert 20 <= ph <= 50 and 20 <= pw <= 50 and 0 <= y0 <= 400 - ph and 0 <= x0 <= 720 - pw
    assert len(seen) == 16


def test_shard_range(bg):
    from bevgen_b200 import sharding
    for n in (0, 1, 7, 256):
        for w in (1, 2, 3, 8):
            got = [i for r in range(w) for i in sharding.shard_range(n, r, w)]
            assert got == list(range(n))
            sizes = [len(sharding.shard_range(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.gather_stats({"a": 1.5}) == {"a": [1.5]}


def test_no_cuda_means_loud_failure(bg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU"):
        bg.create_intentnet_lidar_bev([np.zeros((3, 3))], [np.ones(3, np.float32)])
    with pytest.raises(RuntimeError):
        bg.transform_points(np.zeros((3, 3), np.float32), np.eye(4))


def test_pinned_pool_never_aliases_live_results(monkeypatch):
    """The reference-signature functions return fresh arrays (utils.py:70, :146, :182); ours are backed by pooled pinned
    buffers that may only be recycled once the array, its views and its torch.from_numpy aliases are all gone."""
    import gc
    import torch
    from bevgen_b200 import bev_ops
    pool = bev_ops._PinnedPool()

    def take(n):                                   # no CUDA here: pageable stand-ins for the pinned buffers
        lst = pool._free.get(n)
        if lst:
            pool._idle_bytes -= n
            return lst.pop()
        return torch.empty(n, dtype=torch.uint8)

    class _Stream:
        def synchronize(self):
            pass

    monkeypatch.setattr(pool, "take", take)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    t = torch.arange(1 << 19, dtype=torch.float32)
    a, b = pool.to_numpy(t), pool.to_numpy(t + 1)
    assert a.ctypes.data != b.ctypes.data and a[5] == 5 and b[5] == 6      # two live results: two buffers
    view, ptr = a[10:20], a.ctypes.data
    del a
    gc.collect()
    assert pool._idle_bytes == 0                   # a view keeps the buffer out of the pool
    alias = torch.from_numpy(view)                 # what dataset.py:355 does
    del view
    gc.collect()
    assert pool._idle_bytes == 0
    del alias
    gc.collect()
    assert pool._idle_bytes == t.numel() * 4       # now it is back
    c = pool.to_numpy(t * 2)
    assert c.ctypes.data == ptr and c[3] == 6 and b[5] == 6               # recycled; the live result is untouched


def test_offset_validation(bg):
    """Host-side CSR checks: offsets must be ordered and inside the point array; max_sweep_points is a sizing
    hint that is corrected upward, never trusted below the largest sweep."""
    from bevgen_b200.bev_ops import _check_offsets
    assert _check_offsets(np.array([0, 5, 5, 12]), 12, None) == 7
    assert _check_offsets(np.array([0, 5, 5, 12]), 12, 3) == 7          # too small a hint is corrected
    assert _check_offsets(np.array([0, 5, 5, 12]), 12, 100) == 100
    assert _check_offsets(np.zeros(0), 0, None) == 0
    for bad in ([0, 6, 5, 12], [-1, 5, 12], [0, 5, 13]):
        with pytest.raises(ValueError):
            _check_offsets(np.array(bad), 12, None)


def test_sweep_cache_is_byte_bounded():
    from bevgen_b200.sweep_io import _SweepCache
    c = _SweepCache(max_bytes=10 * 1024)
    blk = lambda: np.zeros((512, 4), np.float16)      # 4 KiB + overhead
    c.put("a", blk()); c.put("b", blk())
    assert len(c) == 2 and c.nbytes <= c.max_bytes
    assert c.get("a")[0]                               # touch a -> b is now the eviction candidate
    c.put("c", blk())
    assert len(c) == 2 and c.get("a")[0] and c.get("c")[0] and not c.get("b")[0]
    c.put("none", None)                                # "missing sweep" results are cached too (tiny)
    assert c.get("none") == (True, None)
    c.put("huge", np.zeros((4096, 4), np.float32))     # larger than the whole budget: not cached, nothing evicted for it
    assert not c.get("huge")[0] and c.nbytes <= c.max_bytes
    c.resize(0)
    assert len(c) == 0 and c.nbytes == 0
    c.put("a", blk())
    assert len(c) == 0


def test_map_cache_is_entry_bounded(bg, tmp_path):
    from bevgen_b200.maps import LruDict
    d = LruDict(2)
    d.store("a", 1); d.store("b", 2)
    assert d.lookup("a") == 1                    # a is now the most recent
    d.store("c", 3)
    assert list(d) == ["a", "c"] and d.lookup("b") is None
    md = bg.synthetic.synthetic_map(seed=1, rows=2, cols=2, n_cross=1)
    p = bg.synthetic.write_map_json(md, tmp_path / "m.json")
    bg.maps.clear_cache()
    a = bg.maps.load_map_json(p)
    assert bg.maps.load_map_json(p) is a and bg.maps.load_map_json(p, use_cache=False) is not a


def test_sweep_store_ring_allocation_and_pins():
    """the store's row allocator (host logic only: no device buffer is touched): first-in-first-out eviction across the wrap,
    8-row alignment, pinned sweeps are never overwritten"""
    import collections
    from bevgen_b200.sweep_store import SweepStore
    s = SweepStore.__new__(SweepStore)
    s.capacity, s._entries, s._head, s._epoch, s._events = 64, collections.OrderedDict(), 0, 0, {}
    s._last_use, s._pins, s._wait_epoch, s.evictions = {}, collections.Counter(), 0, 0

    def put(k, n):
        r = s._alloc(max(8, (n + 7) // 8 * 8))
        s._entries[k] = (r, n)
        return r

    assert [put(0, 20), put(1, 20)] == [0, 24] and s._head == 48
    assert put(2, 20) == 0 and list(s._entries) == [1, 2]           # wrapped: sweep 0 evicted, sweep 1 (rows 24..47) still there
    s._last_use[1] = 5
    assert put(3, 20) == 24 and list(s._entries) == [2, 3] and s._wait_epoch == 5   # evicting 1 -> wait for its last epoch
    s._pin([2])
    with pytest.raises(RuntimeError):
        put(4, 20)                                                   # would overwrite pinned sweep 2
    assert list(s._entries) == [2, 3] and s._wait_epoch == 5        # (a failed allocation keeps the pending wait)
    s._unpin([2])
    assert put(4, 20) == 0 and 2 not in s._entries and not s._pins
    with pytest.raises(ValueError):
        s._alloc(72)
    # the event an overwrite waits for: the epoch's own, the oldest kept one once it has been pruned, none if never read
    s._events = {7: "e7", 9: "e9"}
    assert s._wait_event(9) == "e9" and s._wait_event(3) == "e7" and s._wait_event(0) is None
    # a batch whose sweeps have moved / gone since it was packed is refused
    from bevgen_b200.sweep_store import StoreBatch
    b = StoreBatch.__new__(StoreBatch)
    b.rows = {k: s._entries[k] for k in s._entries}
    s.check_resident(b)
    b.rows = {2: (0, 20)}
    with pytest.raises(KeyError):
        s.check_resident(b)


def test_sweep_store_ring_never_overlaps_live_sweeps():
    """random put sequences against the allocator: live sweeps never share rows, rows stay 8-aligned and inside the ring"""
    import collections
    from hypothesis import given, settings, strategies as st
    from bevgen_b200.sweep_store import SweepStore

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(min_value=0, max_value=70), min_size=1, max_size=60), st.integers(min_value=80, max_value=400))
    def run(sizes, cap):
        s = SweepStore.__new__(SweepStore)
        s.capacity, s._entries, s._head, s._epoch, s._events = (cap + 7) // 8 * 8, collections.OrderedDict(), 0, 0, {}
        s._last_use, s._pins, s._wait_epoch, s.evictions = {}, collections.Counter(), 0, 0
        for k, n in enumerate(sizes):
            need = max(8, (n + 7) // 8 * 8)
            r = s._alloc(need)
            s._entries[k] = (r, n)
            assert r % 8 == 0 and 0 <= r and r + need <= s.capacity
            spans = sorted((row, row + max(8, (m + 7) // 8 * 8)) for row, m in s._entries.values())
            assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])), spans
            assert k in s._entries

    run()


def test_uint8_intensities_fit_the_four_byte_bin_record():
    """csrc/lidar.cu band_rec_pack / band_rec_unpack: AV2 columns carry uint8 intensities; as float32 (dataset.py:321) every
    one of them has its low 16 bits clear, so `cell | (bits >> 16) << 13` keeps cell (13 bits: 10 rows x 720 columns) and
    intensity exactly, fits 32 bits, and orders like the float (the tile kernel takes the max of the unpacked bits)"""
    v = np.arange(256, dtype=np.float32)
    bits = v.view(np.uint32)
    assert not (bits & 0xFFFF).any()
    cell = np.uint32(10 * 720 - 1)
    assert int(cell) < 1 << 13
    w = cell | ((bits >> np.uint32(16)) << np.uint32(13))
    assert int(w.astype(np.uint64).max()) < 1 << 32
    assert np.array_equal(w & np.uint32(8191), np.full(256, cell)) and np.array_equal((w >> np.uint32(13)) << np.uint32(16), bits)
    assert (np.diff(bits.astype(np.int64)) > 0).all()                # unsigned order of the bits == order of the values


def test_store_batch_tables_are_one_aligned_buffer(bg):
    """StoreBatch.pack (host only): the four per-frame tables are views of ONE buffer (one upload), float64 views 8-byte
    aligned, contents as the separate arrays would be; missing sweeps are empty ranges at row 0"""
    import collections
    from bevgen_b200.sweep_store import SweepStore, StoreBatch, StoreFrame
    s = SweepStore.__new__(SweepStore)
    s._entries = collections.OrderedDict({("L", 0): (0, 11), ("L", 1): (16, 5), ("L", 2): (24, 0)})
    s._pins = collections.Counter()
    poses = bg.synthetic._poses(3)
    frames = [StoreFrame([("L", 0), None, ("L", 1)], poses, poses[2]), StoreFrame([("L", 2), ("L", 1)], poses[:2], poses[1])]
    b = StoreBatch.pack(s, frames, 4, map_ids=[0, 1])
    assert b.tables.dtype.is_floating_point is False and b.tables.numel() == b.h2d_bytes
    for name, dt, off, n in b.layout:
        v = getattr(b, name).reshape(-1)
        assert v.data_ptr() == b.tables.data_ptr() + off and v.numel() == n and off % v.element_size() == 0, name
    assert b.sweep_start.tolist() == [0, 0, 16, 0, 24, 16, 0, 0]
    assert b.sweep_offsets.tolist() == [0, 11, 11, 16, 16, 16, 21, 21, 21]
    assert b.total_points == 21 and b.max_sweep_points == 11 and set(b.keys) == {("L", 0), ("L", 1), ("L", 2)}
    want = bg.pose.sweep_rel_transforms(poses[2], poses)
    assert np.array_equal(b.rel_tf.numpy()[:3].reshape(3, 4, 4), want) and np.array_equal(b.rel_tf.numpy()[3].reshape(4, 4), np.eye(4))
    assert np.array_equal(b.ego_T.numpy()[1].reshape(3, 3), bg.pose.map_matrix_from_pose(poses[1][0], poses[1][1], poses[1][3:7]))
    assert s._pins[("L", 1)] == 1                                    # pinned until generated / dropped
    b._release()
    assert not s._pins
    with pytest.raises(IndexError):
        StoreBatch.pack(s, [StoreFrame([("L", 0)] * 5, np.tile(poses[:1], (5, 1)), poses[0])], 4)
    with pytest.raises(KeyError):
        StoreBatch.pack(s, [StoreFrame([("L", 9)], poses[:1], poses[0])], 4)


def test_deferred_points_read_like_the_transformed_array(bg, monkeypatch):
    """bev_ops.DeferredPoints (what install() makes dataset.transform_points return): shape / len / dtype without any work;
    every other way of reading it -- np.asarray, indexing, iteration, arithmetic, ndarray attributes -- yields the array
    transform_points returns, computed once (the CUDA call is stubbed here with the reference formula, utils.py:31-33)"""
    calls = []

    def ref_transform(points, T):
        calls.append(1)
        p = np.asarray(points, np.float64)
        return (np.asarray(T) @ np.c_[p[:, :3], np.ones(len(p))].T).T[:, :3]

    monkeypatch.setattr(bg.bev_ops, "transform_points", ref_transform)
    raw = np.asfortranarray(np.random.default_rng(0).uniform(-50, 50, (17, 3)).astype(np.float16))
    T = np.eye(4); T[:3, 3] = [1.0, -2.0, 0.5]; T[0, 1] = 0.25
    d = bg.bev_ops.transform_points_deferred(raw, T)
    assert isinstance(d, bg.bev_ops.DeferredPoints) and d.shape == (17, 3) and len(d) == 17 and d.ndim == 2 and d.dtype == np.float64
    assert not calls                                                  # nothing computed yet
    raw_before = raw.copy()
    raw[:] = 0                                                        # the caller's array may change: the snapshot does not
    want = ref_transform(raw_before, T); calls.clear()
    assert np.array_equal(np.asarray(d), want) and np.array_equal(d[3:5, :2], want[3:5, :2])
    assert np.array_equal(d * 2.0, want * 2.0) and np.array_equal(1.0 + d, 1.0 + want) and np.array_equal(-d, -want)
    assert np.array_equal(d.T, want.T) and d.mean() == want.mean() and np.array_equal(np.stack(list(d)), want)
    assert np.array_equal(np.asarray(d, dtype=np.float32), want.astype(np.float32)) and (d == want).all()
    assert len(calls) == 1                                            # materialised once
    # inputs that are not float16 / float32 sweeps are transformed right away
    assert isinstance(bg.bev_ops.transform_points_deferred(raw_before.astype(np.float64), T), np.ndarray)
    with pytest.raises(ValueError):
        bg.bev_ops.transform_points_deferred(raw_before, np.eye(3))
