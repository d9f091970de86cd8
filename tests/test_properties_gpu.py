"""GPU property tests (hypothesis): size-independent invariants of the CUDA path."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bg():
    import bevgen_b200
    return bevgen_b200


def _frame(seed, n, S):
    import bevgen_b200 as bg
    return bg.synthetic.frame_c2(seed, n=n, S=S)


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10_000), n=st.integers(1, 3000), S=st.integers(1, 10), algo=st.sampled_from([0, 1, 2]))
def test_lidar_matches_oracle_and_is_order_free(bg, seed, n, S, algo):
    from oracle import c_oracle as co
    f = _frame(seed, n, S)
    want = co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, S)
    got = bg.lidar_bev_from_sweeps(f.points4, f.sweep_offsets, f.rel_tf, 1, S, algo=algo)[0]
    assert torch.equal(got.cpu(), torch.from_numpy(want))
    # max-scatter: any permutation of the points of a sweep gives the same grid
    rng = np.random.default_rng(seed)
    p2 = f.points4.copy()
    for s in range(S):
        a, b = f.sweep_offsets[s], f.sweep_offsets[s + 1]
        p2[a:b] = p2[a:b][rng.permutation(b - a)]
    again = bg.lidar_bev_from_sweeps(p2, f.sweep_offsets, f.rel_tf, 1, S, algo=algo)[0]
    assert torch.equal(got, again)
    # idempotence / linearity of max: feeding every point twice changes nothing
    dup = np.concatenate([f.sweep(s) for s in range(S) for _ in (0, 1)])
    off2 = np.concatenate([[0], np.cumsum(np.repeat(np.diff(f.sweep_offsets), 1) * 2)]).astype(np.int32)
    twice = bg.lidar_bev_from_sweeps(dup, off2, f.rel_tf, 1, S, algo=algo)[0]
    assert torch.equal(got, twice)


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 1000), angle=st.floats(-15, 15), scale=st.floats(0.95, 1.05), flip=st.booleans())
def test_augment_composition_matches_oracle(bg, seed, angle, scale, flip):
    from oracle import c_oracle as co
    rng = np.random.default_rng(seed)
    g = np.zeros((2, 400, 720), np.float32)
    idx = rng.integers(0, g.size, 3000)
    g.reshape(-1)[idx] = rng.integers(1, 256, idx.size)
    p = bg.AugParams(flip=flip, angle_deg=angle, scale=scale)
    got = bg.augment_grid(torch.from_numpy(g).cuda()[None], [p])[0].cpu().numpy()
    want = co.scale(co.rotate(co.flip_w(g) if flip else g, angle), scale)
    assert np.array_equal(got, want)
    # double flip is the identity
    once = bg.augment_grid(torch.from_numpy(g).cuda()[None], [bg.AugParams(flip=True)])
    back = bg.augment_grid(once, [bg.AugParams(flip=True)])[0].cpu().numpy()
    assert np.array_equal(back, g)


def test_batch_equals_singles_and_shards(bg):
    """a batch is the concatenation of its frames; sharding the batch over ranks is the same computation"""
    from bevgen_b200 import sharding
    frames = [bg.synthetic.frame_c2(900 + i, n=1500, S=10) for i in range(5)]
    fb = bg.FrameBatch.pack(frames, 10)
    gen = bg.BevGenerator()
    whole = gen.generate(fb, with_map=False)["lidar_bev"].clone()
    for w in (2, 3):
        parts = []
        for r in range(w):
            sub = sharding.shard_frames(frames, r, w)
            parts.append(gen.generate(bg.FrameBatch.pack(sub, 10), with_map=False)["lidar_bev"].clone())
        assert torch.equal(torch.cat(parts), whole)


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10_000), n=st.integers(1, 4000), S=st.integers(1, 4), angle=st.one_of(st.none(), st.floats(-15, 15)),
       scale=st.one_of(st.none(), st.floats(0.95, 1.05)), flip=st.booleans())
def test_fused_generation_plus_augmentation_matches_oracle_and_is_order_free(bg, seed, n, S, angle, scale, flip):
    """bevgen_lidar_bev_augmented == oracle generation followed by the oracle's flip -> rotate -> scale, for any draw in
    the reference's parameter ranges; permuting the points of a sweep changes nothing (max is order-free)"""
    from oracle import c_oracle as co
    f = _frame(seed, n, S)
    p = bg.AugParams(flip=flip, angle_deg=angle, scale=scale)
    g = co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, S)
    if flip: g = co.flip_w(g)
    if angle is not None: g = co.rotate(g, angle)
    if scale is not None: g = co.scale(g, scale)
    dev = torch.device("cuda")
    fb = bg.FrameBatch.pack([f], S)
    args = (fb.sweep_offsets.to(dev), fb.rel_tf.to(dev), 1, S, [p])
    got = bg.augment.lidar_bev_augmented(fb.points4.to(dev), *args, max_sweep_points=fb.max_sweep_points)[0]
    assert torch.equal(got.cpu(), torch.from_numpy(np.ascontiguousarray(g)))
    rng = np.random.default_rng(seed)
    p2 = f.points4.copy()
    for s in range(S):
        a, b = f.sweep_offsets[s], f.sweep_offsets[s + 1]
        p2[a:b] = p2[a:b][rng.permutation(b - a)]
    again = bg.augment.lidar_bev_augmented(torch.from_numpy(p2).to(dev), *args, max_sweep_points=fb.max_sweep_points)[0]
    assert torch.equal(got, again)


def test_fused_equals_dense_route_at_full_size(bg):
    """BASELINE.json's frame size (10 sweeps x 100k points): building the augmented grid from the point records gives
    the same bits as generating the grid and resampling it (size-independent property; the oracle takes minutes here)"""
    frames = [bg.synthetic.frame_c2(70 + i) for i in range(3)]
    aug = [bg.AugParams(angle_deg=-13.7, flip=True), bg.AugParams(scale=1.047, rects=[(30, 600, 45, 50)]),
           bg.AugParams(angle_deg=8.1, scale=0.953)]
    fb = bg.FrameBatch.pack(frames, 10)
    a = bg.BevGenerator(fused_augment=True).generate(fb, aug, with_map=False)["lidar_bev"]
    b = bg.BevGenerator(fused_augment=False).generate(fb, aug, with_map=False)["lidar_bev"]
    assert torch.equal(a, b)
    assert int((a != 0).sum()) > 3_000_000
