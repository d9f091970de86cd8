"""GPU parity: CUDA LiDAR path (through the C ABI) vs the C oracle, bit-exact.

Reference rows covered: transform_points (utils.py:27-33), create_intentnet_lidar_bev
(utils.py:62-106), flip / dropout fusion (utils.py:399, :496).
"""
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


def _oracle_frames(co, frames, S):
    return np.stack([co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, S) for f in frames])


def _batch(frames, S):
    pts = np.concatenate([f.points4 for f in frames], 0)
    offs = [np.zeros(1, np.int64)]
    base = 0
    for f in frames:
        o = np.zeros(S, np.int64)
        o[:f.num_sweeps] = np.diff(f.sweep_offsets)
        offs.append(base + np.cumsum(o))
        base += f.num_points
    off = np.concatenate(offs).astype(np.int32)
    tf = np.zeros((len(frames), S, 4, 4))
    tf[:] = np.eye(4)
    for b, f in enumerate(frames):
        tf[b, :f.num_sweeps] = f.rel_tf
    return pts, off, tf.reshape(-1, 4, 4)


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_fused_single_frame_matches_oracle(bg, co, algo):
    f = bg.synthetic.frame_c2(0, n=20_000)
    want = co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, 10)
    got = bg.lidar_bev_from_sweeps(f.points4, f.sweep_offsets, f.rel_tf, 1, 10, algo=algo)[0].cpu().numpy()
    assert got.shape == (290, 400, 720) and got.dtype == np.float32
    assert np.array_equal(got, want)
    assert np.count_nonzero(got) > 50_000


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_fused_ragged_batch(bg, co, algo):
    S = 10
    frames = [bg.synthetic.frame_c2(1, n=5_000), bg.synthetic.frame_c1(0, n=7_001),
              bg.synthetic.frame_c4(0, n=3_000), bg.synthetic.frame_c2(2, n=1, S=3)]
    pts, off, tf = _batch(frames, S)
    got = bg.lidar_bev_from_sweeps(pts, off, tf, len(frames), S, algo=algo).cpu().numpy()
    want = _oracle_frames(co, frames, S)
    assert np.array_equal(got, want)
    # frame_c1 has one loaded sweep: trailing 9*29 channels are zero (utils.py:71-73)
    assert not got[1, 29:].any()


def test_empty_batch_is_all_zero(bg):
    S = 10
    off = np.zeros(2 * S + 1, np.int32)
    got = bg.lidar_bev_from_sweeps(np.zeros((0, 4), np.float32), off, np.tile(np.eye(4), (2 * S, 1, 1)), 2, S)
    assert got.shape == (2, 290, 400, 720) and not got.any().item()


def test_contention_cluster(bg, co):
    """config-4 style: all points in a handful of voxels (warp-aggregated atomics path)."""
    f = bg.synthetic.frame_c4(3, n=60_000)
    want = co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, 10)
    for algo in (0, 1, 2):
        got = bg.lidar_bev_from_sweeps(f.points4, f.sweep_offsets, f.rel_tf, 1, 10, algo=algo)[0].cpu().numpy()
        assert np.array_equal(got, want)
    assert np.count_nonzero(want) < 60_000


def test_bin_edges_and_non_finite(bg, co):
    """points exactly on voxel / height-bin edges, on the grid border, NaN / inf coordinates."""
    k = np.arange(-5, 406)
    xs = (60.0 - 0.2 * k).astype(np.float32)                 # row edges incl. just outside
    ys = (-72.0 + 0.2 * np.arange(-5, 726)).astype(np.float32)
    zs = (-2.0 + 0.2 * np.arange(-2, 32)).astype(np.float32)  # z-bin edges, Z_MIN, Z_MAX
    rng = np.random.default_rng(7)
    n = 40_000
    p = np.empty((n, 4), np.float32)
    p[:, 0] = rng.choice(xs, n); p[:, 1] = rng.choice(ys, n); p[:, 2] = rng.choice(zs, n)
    p[:, 3] = rng.integers(1, 256, n)
    p[:50, 0] = np.nan; p[50:100, 1] = np.inf; p[100:150, 2] = -np.inf; p[150:200, 0] = 1e30
    p[200:260, 3] = 0.0                                       # zero intensity: no-op
    p[260:300, 2] = [np.nextafter(np.float32(3.8), np.float32(0))] * 40
    off = np.array([0, n], np.int32)
    tf = np.eye(4)[None]
    want = co.lidar_bev_fused(p, off, tf, 1)
    for algo in (0, 1, 2):
        got = bg.lidar_bev_from_sweeps(p, off, tf, 1, 1, algo=algo)[0].cpu().numpy()
        assert np.array_equal(got, want)


def test_reference_signature_f64_lists(bg, co):
    """create_intentnet_lidar_bev(points_list, intensity_list): None / empty / short lists (utils.py:71-76)."""
    f = bg.synthetic.frame_c2(5, n=8_000)
    pl, il = [], []
    for s in range(f.num_sweeps):
        sw = f.sweep(s)
        pl.append(co.transform_points(sw[:, :3], f.rel_tf[s])); il.append(sw[:, 3].copy())
    pl[2] = None; il[2] = None
    pl[4] = np.zeros((0, 3)); il[4] = np.zeros((0,), np.float32)
    pl, il = pl[:7], il[:7]                                   # fewer sweeps than expected
    want = co.lidar_bev(pl, il, 10)
    got = bg.create_intentnet_lidar_bev(pl, il)
    assert isinstance(got, np.ndarray) and got.dtype == np.float32 and got.flags.c_contiguous
    assert np.array_equal(got, want)
    assert not got[2 * 29:3 * 29].any() and not got[7 * 29:].any()
    got5 = bg.create_intentnet_lidar_bev(pl[:5], il[:5], 5)
    assert got5.shape == (145, 400, 720) and np.array_equal(got5, want[:145])
    with pytest.raises(IndexError):
        bg.create_intentnet_lidar_bev(pl, il, 3)
    t = bg.create_intentnet_lidar_bev(pl, il, as_tensor=True)
    assert t.is_cuda and torch.equal(t.cpu(), torch.from_numpy(want))


def test_transform_points(bg, co):
    f = bg.synthetic.frame_c2(6, n=30_000, S=2)
    for dt in (np.float32, np.float64, np.float16):
        pts = f.sweep(1)[:, :3].astype(dt)
        got = bg.transform_points(pts, f.rel_tf[0])
        want = co.transform_points(pts.astype(np.float32) if dt == np.float16 else pts, f.rel_tf[0])
        assert got.dtype == np.float64 and got.shape == (30_000, 3)
        assert np.array_equal(got, want)
    e = bg.transform_points(np.zeros((0, 3), np.float32), np.eye(4))
    assert e.shape == (0, 3) and e.dtype == np.float32


def test_flip_and_dropout_fusion(bg, co):
    S = 4
    frames = [bg.synthetic.frame_c2(10 + i, n=6_000, S=S) for i in range(3)]
    pts, off, tf = _batch(frames, S)
    flip = np.array([1, 0, 1], np.uint8)
    rects = np.zeros((3, 5, 4), np.int32)
    cnt = np.array([2, 5, 0], np.int32)
    rects[0, :2] = [[100, 300, 50, 50], [0, 0, 20, 33]]
    rects[1, :5] = [[380, 700, 50, 50], [150, 340, 25, 45], [10, 10, 20, 20], [200, 0, 50, 20], [399, 719, 20, 20]]
    want = _oracle_frames(co, frames, S)
    for algo in (0, 1, 2):
        got = bg.lidar_bev_from_sweeps(pts, off, tf, 3, S, flip=flip, drop_rects=rects, drop_counts=cnt, algo=algo).cpu().numpy()
        for b in range(3):
            w = co.flip_w(want[b]) if flip[b] else want[b].copy()
            co.dropout(w, rects[b, :cnt[b]])
            assert np.array_equal(got[b], w), (algo, b)


def test_full_size_frame_and_properties(bg, co):
    """config-2 frame at full size (10 x 100k points): oracle equality, permutation invariance,
    batch-vs-single equivalence, tiled == atomic == memset."""
    f = bg.synthetic.frame_c2(0)
    want = co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, 10)
    got = bg.lidar_bev_from_sweeps(f.points4, f.sweep_offsets, f.rel_tf, 1, 10)
    assert torch.equal(got[0].cpu(), torch.from_numpy(want))
    # permutation inside every sweep does not change a max-scatter
    rng = np.random.default_rng(0)
    p2 = f.points4.copy()
    for s in range(10):
        a, b = f.sweep_offsets[s], f.sweep_offsets[s + 1]
        p2[a:b] = p2[a:b][rng.permutation(b - a)]
    got2 = bg.lidar_bev_from_sweeps(p2, f.sweep_offsets, f.rel_tf, 1, 10)
    assert torch.equal(got, got2)
    g = bg.synthetic.frame_c4(0)
    pts, off, tf = _batch([f, g], 10)
    both = bg.lidar_bev_from_sweeps(pts, off, tf, 2, 10, algo=0)
    assert torch.equal(both[0], got[0])
    for algo in (1, 2):
        other = bg.lidar_bev_from_sweeps(pts, off, tf, 2, 10, algo=algo)
        assert torch.equal(both, other), algo
    assert torch.equal(both[1].cpu(), torch.from_numpy(co.lidar_bev_fused(g.points4, g.sweep_offsets, g.rel_tf, 10)))


def test_fp16_packed_points(bg, co):
    """AV2-precision input (4 x fp16, 8 B/point): identical to the float32 path / oracle on the same values."""
    frames = [bg.synthetic.quantize_f16(bg.synthetic.frame_c2(70 + i, n=4_000)) for i in range(2)]
    want = _oracle_frames(co, frames, 10)
    fb = bg.FrameBatch.pack(frames, 10, points_dtype=np.float16)
    assert fb.points4.dtype == torch.float16 and fb.h2d_bytes < 0.6 * bg.FrameBatch.pack(frames, 10).h2d_bytes
    for algo in (0, 1):
        got = bg.BevGenerator(algo=algo).generate(fb, with_map=False)["lidar_bev"].cpu().numpy()
        assert np.array_equal(got, want), algo
    with pytest.raises(ValueError):
        bg.FrameBatch.pack([bg.synthetic.frame_c2(3, n=100)], 10, points_dtype=np.float16)   # not fp16-representable


def test_reference_signature_results_are_independent_arrays(bg):
    """create_intentnet_lidar_bev returns a new array per call (utils.py:70): results backed by pooled pinned memory must
    not alias while alive, and must be ordinary writable numpy arrays"""
    f = bg.synthetic.frame_c1(0, n=3_000)
    pts = [bg.transform_points(f.sweep(0)[:, :3], f.rel_tf[0])]
    inten = [f.sweep(0)[:, 3].copy()]
    a = bg.create_intentnet_lidar_bev(pts, inten)
    b = bg.create_intentnet_lidar_bev(pts, inten)
    assert a.ctypes.data != b.ctypes.data and np.array_equal(a, b)
    assert a.flags.writeable and a.flags.c_contiguous and a.dtype == np.float32
    a[0, 0, 0] = 123.0
    assert b[0, 0, 0] != 123.0
    t = torch.from_numpy(a).float()                 # dataset.py:355: no copy for float32
    del a
    c = bg.create_intentnet_lidar_bev(pts, inten)
    assert t[0, 0, 0] == 123.0 and c[0, 0, 0] != 123.0


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_av2_column_points(bg, co, algo):
    """AV2's on-disk columns (float16 x,y,z + uint8 intensity, 7 B/point, dataset.py:311-321): identical to the oracle on
    the same values; sweeps of odd lengths start at unaligned rows (the kernels read aligned 8-row vectors and mask),
    empty and missing sweeps included."""
    S = 10
    frames = [bg.synthetic.quantize_f16(f) for f in
              (bg.synthetic.frame_c2(80, n=4_001), bg.synthetic.frame_c1(2, n=777), bg.synthetic.frame_c4(1, n=3_003),
               bg.synthetic.frame_c2(81, n=1, S=3), bg.synthetic.frame_c2(82, n=9_000))]
    want = _oracle_frames(co, frames, S)
    fb = bg.FrameBatch.pack(frames, S, points_dtype="av2")
    assert fb.points_format == "av2" and fb.point_bytes == 7 and fb.num_points == sum(f.num_points for f in frames)
    assert fb.h2d_bytes < 0.46 * bg.FrameBatch.pack(frames, S).h2d_bytes
    got = bg.BevGenerator(algo=algo).generate(fb, with_map=False)["lidar_bev"].cpu().numpy()
    assert np.array_equal(got, want)
    # direct call on host columns; sub-range of frames (absolute CSR offsets, pointer offset of the offsets only)
    pc = fb.points4
    off = fb.sweep_offsets.numpy()
    got2 = bg.lidar_bev_from_sweeps(pc, off[2 * S:5 * S + 1], fb.rel_tf.numpy()[2 * S:5 * S], 3, S, algo=algo).cpu().numpy()
    assert np.array_equal(got2, want[2:5])
    with pytest.raises(ValueError):
        bg.FrameBatch.pack([bg.synthetic.frame_c2(3, n=100)], S, points_dtype="av2")         # not fp16-representable
    f = bg.synthetic.quantize_f16(bg.synthetic.frame_c2(3, n=100))
    f.points4[5, 3] = 0.5
    with pytest.raises(ValueError):
        bg.FrameBatch.pack([f], S, points_dtype="av2")                                       # intensity not a uint8 value


def test_av2_columns_non_finite_and_extreme_values(bg, co):
    """float16 columns can hold NaN, +-inf and +-65504: such points are masked exactly as the reference masks them
    (utils.py:83-87 on the float64 values), bin-edge coordinates that ARE float16 values land in the reference's voxel"""
    rng = np.random.default_rng(5)
    n = 4096
    p = np.empty((n, 4), np.float32)
    p[:, 0] = rng.choice((60.0 - 0.25 * np.arange(-8, 330)).astype(np.float16).astype(np.float32), n)     # multiples of 0.25: exact in fp16
    p[:, 1] = rng.choice((-72.0 + 0.25 * np.arange(-8, 590)).astype(np.float16).astype(np.float32), n)
    p[:, 2] = rng.choice((-2.0 + 0.125 * np.arange(-4, 52)).astype(np.float16).astype(np.float32), n)
    p[:, 3] = rng.integers(0, 256, n)
    p[:16, 0] = np.nan; p[16:32, 1] = np.inf; p[32:48, 2] = -np.inf; p[48:64, 0] = 65504.0; p[64:80, 1] = -65504.0
    off = np.array([0, n], np.int32)
    want = co.lidar_bev_fused(p, off, np.eye(4)[None], 1)
    assert np.count_nonzero(want) > 1000
    pc = bg.PointColumns.from_points4(p)
    for algo in (0, 1):
        got = bg.lidar_bev_from_sweeps(pc, off, np.eye(4)[None], 1, 1, algo=algo)[0].cpu().numpy()
        assert np.array_equal(got, want), algo
