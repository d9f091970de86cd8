"""GPU parity: generation FUSED with augmentation (bevgen_lidar_bev_augmented, csrc/lidar_aug_tile.cuh) vs the C oracle.

The fused kernel builds the rotated / scaled grid directly from the binned point records; the expected result is the
oracle's create_intentnet_lidar_bev (utils.py:62-106) followed by the oracle's flip -> rotate -> scale -> dropout
(utils.py:508-511).  Bit-exact in every case (the oracle's rotate equals cv2.warpAffine bit for bit and its scale equals
cv2.resize with IPP off; tests/test_oracle_vs_reference.py).
"""
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


def _pack(bg, frames, S, dtype=None):
    batch = bg.FrameBatch.pack(frames, S, points_dtype=dtype)
    dev = torch.device("cuda")
    return batch, batch.points4.to(dev), batch.sweep_offsets.to(dev), batch.rel_tf.to(dev)


def _check(bg, co, frames, params, S=10, dtype=None):
    batch, pts, off, tf = _pack(bg, frames, S, dtype)
    got = bg.augment.lidar_bev_augmented(pts, off, tf, len(frames), S, params, max_sweep_points=batch.max_sweep_points)
    got = got.cpu().numpy()
    for b, (f, p) in enumerate(zip(frames, params)):
        rel = np.tile(np.eye(4), (S, 1, 1))
        rel[:f.num_sweeps] = f.rel_tf
        offs = np.zeros(S + 1, np.int64)
        offs[1:f.num_sweeps + 1] = f.sweep_offsets[1:]
        offs[f.num_sweeps + 1:] = f.sweep_offsets[-1]
        pts4 = np.asarray(f.points4, np.float32)
        want = _oracle_aug(co, co.lidar_bev_fused(pts4, offs, rel, S), p)
        bad = got[b] != want
        assert not bad.any(), (b, p, int(bad.sum()), np.argwhere(bad)[:5].tolist(), float(np.abs(got[b] - want).max()))
    return got


CASES = [
    dict(angle_deg=7.3), dict(angle_deg=-15.0), dict(angle_deg=15.0), dict(angle_deg=1e-9), dict(flip=True, angle_deg=-3.25),
    dict(scale=0.95), dict(scale=1.05), dict(scale=1.0), dict(scale=1.0009), dict(scale=0.9991), dict(flip=True, scale=1.0312),
    dict(angle_deg=11.0, scale=0.9731), dict(angle_deg=15.0, scale=0.95), dict(angle_deg=-15.0, scale=1.05),
    dict(flip=True, angle_deg=-9.5, scale=1.0417, rects=[(150, 300, 50, 50), (0, 0, 30, 30), (360, 680, 40, 40)]),
    dict(angle_deg=4.0, rects=[(100, 300, 40, 40)]), dict(scale=0.97, rects=[(10, 10, 50, 50)]),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_fused_matches_generate_then_augment(bg, co, case):
    p = bg.AugParams(**CASES[case])
    f = bg.synthetic.frame_c2(case, n=30_000, S=2)
    _check(bg, co, [f], [p], S=2)


def test_mixed_batch_any_order(bg, co):
    """classes interleaved (no partition): plain / flip+dropout frames take the tile-store route inside the same call"""
    random.seed(4321)
    ps = [bg.draw_aug_params(random) for _ in range(7)]
    ps[2] = bg.AugParams()
    ps[4] = bg.AugParams(flip=True, rects=[(200, 100, 33, 47)])
    ps[5] = bg.AugParams(angle_deg=-12.0, scale=1.02, flip=True)
    frames = [bg.synthetic.frame_c2(10 + i, n=8_000, S=3) for i in range(7)]
    frames[3] = bg.synthetic.frame_c2(99, n=5_000, S=1)          # missing sweeps: trailing channels stay zero
    got = _check(bg, co, frames, ps, S=3)
    assert not got[3, 29:].any()


def test_fp16_points(bg, co):
    f = bg.synthetic.quantize_f16(bg.synthetic.frame_c2(5, n=20_000, S=2))
    _check(bg, co, [f, f], [bg.AugParams(angle_deg=9.0, scale=1.04), bg.AugParams(scale=0.96, flip=True)], S=2, dtype=np.float16)


@pytest.mark.parametrize("params", [dict(angle_deg=-6.5), dict(scale=1.03), dict(angle_deg=13.0, scale=0.96, flip=True)])
def test_dense_scene_overflowing_bins(bg, co, params):
    """config-4 style clusters: the fixed-capacity tile bins overflow, the kernel consumes the per-sweep overflow lists"""
    f = bg.synthetic.frame_c4(2, n=40_000, S=2)
    _check(bg, co, [f], [bg.AugParams(**params)], S=2)


def test_overflowing_row_bands_with_sparse_tiles(bg, co):
    """a thin dense line: the 10-row band bins overflow into the sweep's overflow list while every output tile still holds
    few records -- the overflow records are re-binned per tile and take the fast path"""
    import dataclasses
    rng = np.random.default_rng(11)
    n = 3000
    pts = np.stack([rng.uniform(9.9, 10.3, n), rng.uniform(-72, 72, n), rng.uniform(0.0, 0.15, n),
                    rng.integers(1, 256, n)], 1).astype(np.float32)
    base = bg.synthetic.frame_c1(0, n=10)
    f = dataclasses.replace(base, points4=pts, sweep_offsets=np.array([0, n], np.int64))
    for p in (bg.AugParams(angle_deg=-11.0), bg.AugParams(scale=1.04, flip=True), bg.AugParams(scale=0.96, angle_deg=6.0)):
        _check(bg, co, [f], [p], S=1)


def test_grid_aligned_points_and_borders(bg, co):
    """points on voxel boundaries and along the grid borders (taps that fall outside the grid are zeros)"""
    rng = np.random.default_rng(7)
    xs = np.concatenate([np.linspace(-20, 60, 401), rng.uniform(59.0, 60.0, 400), rng.uniform(-20.0, -19.0, 400)])
    ys = np.concatenate([np.linspace(-72, 72, 401), rng.uniform(71.0, 72.0, 400), rng.uniform(-72.0, -71.0, 400)])
    n = xs.size
    pts = np.stack([rng.permutation(xs), ys, rng.uniform(-2, 3.8, n), rng.integers(1, 256, n)], 1).astype(np.float32)
    import dataclasses
    base = bg.synthetic.frame_c1(0, n=10)
    f = dataclasses.replace(base, points4=pts, sweep_offsets=np.array([0, n], np.int64))
    for p in (bg.AugParams(angle_deg=15.0), bg.AugParams(scale=0.95), bg.AugParams(scale=1.05, angle_deg=-15.0)):
        _check(bg, co, [f], [p], S=1)


def test_frames_without_points_are_all_zero(bg):
    """empty / missing sweeps (utils.py:75-76): every channel stays zero whatever the augmentation"""
    import dataclasses
    base = bg.synthetic.frame_c1(0, n=10)
    empty = dataclasses.replace(base, points4=np.zeros((0, 4), np.float32), sweep_offsets=np.array([0, 0], np.int64))
    batch, pts, off, tf = _pack(bg, [empty, empty], 2)
    ps = [bg.AugParams(angle_deg=12.0, scale=0.97), bg.AugParams(scale=1.05, flip=True)]
    got = bg.augment.lidar_bev_augmented(pts, off, tf, 2, 2, ps, max_sweep_points=batch.max_sweep_points)
    assert got.shape == (2, 58, 400, 720) and not got.any().item()


def test_unsupported_parameters_are_refused_not_approximated(bg):
    assert not bg.augment.fused_supported([bg.AugParams(angle_deg=40.0)])
    assert not bg.augment.fused_supported([bg.AugParams(scale=0.6)])
    f = bg.synthetic.frame_c2(0, n=1_000, S=1)
    batch, pts, off, tf = _pack(bg, [f], 1)
    with pytest.raises(RuntimeError):
        bg.augment.lidar_bev_augmented(pts, off, tf, 1, 1, [bg.AugParams(angle_deg=40.0)], max_sweep_points=batch.max_sweep_points)


def test_generator_routes_through_fused_kernel_and_matches_dense_route(bg):
    md = bg.synthetic.synthetic_map(seed=11, rows=6, cols=5, n_cross=3)
    frames = [bg.synthetic.frame_c3(i, n=6_000, S=10, map_data=md) for i in range(4)]
    table = bg.MapTable([bg.maps.parse_map_dict(md)])
    aug = [bg.AugParams(scale=1.02), bg.AugParams(flip=True), bg.AugParams(angle_deg=5.0, rects=[(50, 60, 30, 30)]),
           bg.AugParams(angle_deg=-8.0, scale=0.97, flip=True)]
    batch = bg.FrameBatch.pack(frames, 10, map_ids=[0] * 4)
    a = bg.BevGenerator(table, fused_augment=True).generate(batch, aug)
    b = bg.BevGenerator(table, fused_augment=False).generate(batch, aug)
    assert torch.equal(a["lidar_bev"], b["lidar_bev"])
    assert torch.equal(a["map_bev"], b["map_bev"])


def test_every_parameter_set_the_host_accepts_runs_and_matches_the_dense_route(bg):
    """the host screening (bevgen_aug_fused_supported) is conservative: whatever it lets through must fit the kernel's
    shared-memory tiles (the kernel traps otherwise) and give the dense route's bits -- swept past the reference's ranges"""
    f = bg.synthetic.frame_c2(3, n=20_000, S=1)
    fb = bg.FrameBatch.pack([f], 1)
    dev = torch.device("cuda")
    pts, off, tf = fb.points4.to(dev), fb.sweep_offsets.to(dev), fb.rel_tf.to(dev)
    plain = bg.lidar_bev_from_sweeps(pts, off, tf, 1, 1, max_sweep_points=fb.max_sweep_points)
    accepted = refused = 0
    for angle in (None, 15.2, -15.6, 16.0, -16.4, 17.0, 19.0, -25.0):
        for scale in (None, 0.90, 0.93, 0.94, 0.949, 1.051, 1.06, 1.08, 1.15):
            if angle is None and scale is None:
                continue
            p = bg.AugParams(angle_deg=angle, scale=scale, flip=(accepted % 2 == 0))
            if not bg.augment.fused_supported([p]):
                refused += 1
                continue
            accepted += 1
            got = bg.augment.lidar_bev_augmented(pts, off, tf, 1, 1, [p], max_sweep_points=fb.max_sweep_points)
            want = bg.augment_grid(plain, [p])
            assert torch.equal(got, want), (angle, scale)
    torch.cuda.synchronize()
    assert accepted >= 8 and refused >= 8, (accepted, refused)


def test_repeated_runs_are_bit_identical_at_full_size(bg):
    """functional race check (compute-sanitizer is not available on the pool): 25 back-to-back runs of the fused path on
    BASELINE-size frames, sparse and clustered, must give one and the same grid -- a missing barrier or a racy shared-memory
    clear shows up as a run-to-run difference"""
    frames = [bg.synthetic.frame_c2(200), bg.synthetic.frame_c4(200), bg.synthetic.frame_c2(201), bg.synthetic.frame_c2(202)]
    aug = [bg.AugParams(angle_deg=14.0, flip=True), bg.AugParams(angle_deg=-7.0, scale=1.03),
           bg.AugParams(scale=0.955, rects=[(100, 100, 50, 50)]), bg.AugParams(angle_deg=3.0, scale=0.99)]
    fb = bg.FrameBatch.pack(frames, 10)
    dev = torch.device("cuda")
    pts, off, tf = fb.points4.to(dev), fb.sweep_offsets.to(dev), fb.rel_tf.to(dev)
    first = bg.augment.lidar_bev_augmented(pts, off, tf, 4, 10, aug, max_sweep_points=fb.max_sweep_points).clone()
    out = torch.empty_like(first)
    for _ in range(25):
        bg.augment.lidar_bev_augmented(pts, off, tf, 4, 10, aug, max_sweep_points=fb.max_sweep_points, out=out)
        assert torch.equal(out, first)
    dense = bg.augment_grid(bg.lidar_bev_from_sweeps(pts, off, tf, 4, 10, max_sweep_points=fb.max_sweep_points), aug)
    assert torch.equal(first, dense)


def test_fused_augmentation_on_av2_columns(bg, co):
    """generation fused with rotate / scale from AV2 column input == the float32 row input on the same values"""
    frames = [bg.synthetic.quantize_f16(bg.synthetic.frame_c2(300 + i, n=6_001)) for i in range(4)]
    aug = [bg.AugParams(angle_deg=7.5), bg.AugParams(flip=True, scale=0.97), bg.AugParams(angle_deg=-11.0, scale=1.04, rects=[(10, 20, 30, 40)]),
           bg.AugParams(flip=True, rects=[(100, 300, 25, 45)])]
    rows = bg.FrameBatch.pack(frames, 10)
    cols = bg.FrameBatch.pack(frames, 10, points_dtype="av2")
    gen = bg.BevGenerator()
    a = gen.generate(rows, aug, with_map=False)["lidar_bev"].clone()
    b = gen.generate(cols, aug, with_map=False)["lidar_bev"]
    assert torch.equal(a, b)
    f = frames[2]
    g = co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, 10)
    g = np.ascontiguousarray(co.scale(co.rotate(g, -11.0), 1.04)).copy()
    co.dropout(g, [(10, 20, 30, 40)])
    assert np.array_equal(b[2].cpu().numpy(), g)
