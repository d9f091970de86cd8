"""GPU parity: CUDA map rasterizer (through the C ABI) vs the C oracle, bit-exact.
Rows: world_to_bev_pixel utils.py:47-60, rasterize_map_ego_centric utils.py:108-182 (cv2.fillPoly / polylines)."""
import json

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


def _poses(n, seed):
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        tx, ty = 1500 + rng.uniform(-90, 90), -700 + rng.uniform(-90, 90)
        q = R.from_euler("xyz", [0.01, -0.02, rng.uniform(-3.1, 3.1)]).as_quat()
        out.append((tx, ty, q))
    return out


def test_world_to_bev_pixel(bg, co):
    rng = np.random.default_rng(3)
    T = bg.pose.ego_centric_matrix(np.array([1500.0, -700.0]), 0.6)
    xy = np.stack([1500 + rng.uniform(-100, 100, 5000), -700 + rng.uniform(-100, 100, 5000)], 1)
    # half-integer pixel coordinates (round-half-even) with the identity matrix
    half = np.stack([(np.arange(-10, 10) + 0.5) * 0.2 * -1 + 60.0, (np.arange(-10, 10) + 0.5) * 0.2], 1)
    for pts, M in ((xy, T), (half, np.eye(3))):
        got = bg.world_to_bev_pixel(pts, M)
        assert got.dtype == np.int64 or got.dtype == int
        assert np.array_equal(got, co.world_to_bev_pixel(pts, M))
    assert bg.world_to_bev_pixel(np.zeros((0, 2)), T).shape == (0, 2)


def test_synthetic_map_batch_matches_oracle(bg, co):
    md = bg.synthetic.synthetic_map()
    fm = bg.maps.parse_map_dict(md)
    table = bg.MapTable([fm])
    poses = _poses(6, 11)
    Ts = np.stack([bg.pose.map_matrix_from_pose(tx, ty, q) for tx, ty, q in poses])
    got = bg.rasterize_maps(table, [0] * len(poses), Ts).cpu().numpy()
    flat = co.flatten_map(md)
    for b in range(len(poses)):
        want = co.map_raster_flat(flat, Ts[b])
        assert np.array_equal(got[b], want), (b, np.argwhere(got[b] != want)[:5])
    assert got.sum() > 10_000 and set(np.unique(got)) <= {0.0, 1.0}


def test_two_maps_in_one_table_and_fusions(bg, co):
    md_a = bg.synthetic.synthetic_map(seed=1, rows=12, cols=9, n_cross=6)
    md_b = bg.synthetic.synthetic_map(seed=2, rows=7, cols=15, n_cross=0)
    md_b["lane_segments"]["0"]["left_lane_boundary"] = []              # boundary missing
    md_b["lane_segments"]["1"]["right_lane_boundary"] = [{"x": 1.0}, "junk", None]   # unusable vertices (utils.py:134)
    table = bg.MapTable([bg.maps.parse_map_dict(md_a), bg.maps.parse_map_dict(md_b)])
    poses = _poses(4, 5)
    Ts = np.stack([bg.pose.map_matrix_from_pose(tx, ty, q) for tx, ty, q in poses])
    ids = [0, 1, 1, 0]
    flip = np.array([1, 0, 1, 0], np.uint8)
    rects = np.zeros((4, 5, 4), np.int32)
    cnt = np.array([0, 3, 1, 0], np.int32)
    rects[1, :3] = [[100, 300, 50, 50], [0, 0, 25, 33], [390, 700, 50, 50]]
    rects[2, 0] = [150, 330, 40, 60]
    got = bg.rasterize_maps(table, ids, Ts, flip=flip, drop_rects=rects, drop_counts=cnt).cpu().numpy()
    flats = [co.flatten_map(md_a), co.flatten_map(md_b)]
    for b in range(4):
        w = co.map_raster_flat(flats[ids[b]], Ts[b])
        if flip[b]:
            w = co.flip_w(w)
        co.dropout(w, rects[b, :cnt[b]])
        assert np.array_equal(got[b], w), b


def test_stress_polygons(bg, co):
    """self-intersecting / long / degenerate boundaries incl. > 32 vertices per boundary and long chords."""
    rng = np.random.default_rng(9)
    lanes = {}
    for i in range(60):
        n = int(rng.integers(2, 90))
        if i % 3 == 0:
            c = rng.uniform([-10, -60], [50, 60]); l = c + rng.uniform(-12, 12, (n, 2)); r = c + rng.uniform(-12, 12, (n, 2))
        elif i % 3 == 1:
            l = rng.uniform([-40, -100], [80, 100], (n, 2)); r = rng.uniform([-40, -100], [80, 100], (n, 2))   # partly off-grid
        else:
            c = rng.uniform([-10, -60], [50, 60]); l = c + rng.uniform(-0.5, 0.5, (n, 2)); r = c + rng.uniform(-0.5, 0.5, (n, 2))
        lanes[str(i)] = {"is_intersection": bool(i % 2), "lane_type": "BUS" if i % 5 == 0 else "VEHICLE",
                         "left_lane_boundary": [{"x": float(a), "y": float(b)} for a, b in l],
                         "right_lane_boundary": [{"x": float(a), "y": float(b)} for a, b in r],
                         "left_lane_mark_type": "SOLID_YELLOW", "right_lane_mark_type": "DASHED_WHITE"}
    cws = {str(k): {"polygon": [{"x": float(a), "y": float(b)} for a, b in rng.uniform([-20, -72], [60, 72], (int(rng.integers(3, 50)), 2))]}
           for k in range(15)}
    md = {"lane_segments": lanes, "pedestrian_crossings": cws}
    table = bg.MapTable([bg.maps.parse_map_dict(md)])
    T = np.eye(3)[None]
    got = bg.rasterize_maps(table, [0], T)[0].cpu().numpy()
    want = co.map_raster_flat(co.flatten_map(md), np.eye(3))
    assert np.array_equal(got, want), [int((got[c] != want[c]).sum()) for c in range(9)]


def test_reference_signature_and_error_conventions(bg, co, tmp_path, capsys):
    md = bg.synthetic.synthetic_map(seed=4, rows=10, cols=10, n_cross=5)
    path = bg.synthetic.write_map_json(md, tmp_path / "map.json")
    f = bg.synthetic.frame_c3(0, n=10, S=2, map_data=md)
    got = bg.rasterize_map_ego_centric(path, f.pose_series())
    T3 = bg.pose.map_matrix_from_pose(f.current_pose[0], f.current_pose[1], f.current_pose[3:7])
    assert got.dtype == np.float32 and got.shape == (9, 400, 720)
    assert np.array_equal(got, co.map_raster_flat(co.flatten_map(md), T3))
    # unreadable JSON -> zeros + message (utils.py:111-116)
    z = bg.rasterize_map_ego_centric(str(tmp_path / "missing.json"), f.pose_series())
    assert not z.any() and "Returning empty map" in capsys.readouterr().out
    # invalid quaternion -> zeros + warning (utils.py:123-127)
    bad = f.pose_series(); bad["qx"] = bad["qy"] = bad["qz"] = bad["qw"] = 0.0
    z = bg.rasterize_map_ego_centric(path, bad)
    assert not z.any() and "Invalid ego quaternion" in capsys.readouterr().out
    # empty map dict
    p2 = tmp_path / "empty.json"; p2.write_text(json.dumps({}))
    assert not bg.rasterize_map_ego_centric(str(p2), f.pose_series()).any()


def test_augmented_maps_from_bit_planes_match_raster_then_augment_and_oracle(bg, co):
    """bevgen_map_raster_augmented: the 9 map channels resampled straight from the rasterizer's bit planes == rasterize,
    then the dense augmentation kernel == the oracle's flip -> rotate -> scale -> dropout of the oracle's map"""
    md = bg.synthetic.synthetic_map(seed=21, rows=8, cols=7, n_cross=5)
    table = bg.MapTable([bg.maps.parse_map_dict(md)])
    poses = _poses(6, 5)
    Ts = np.stack([bg.pose.map_matrix_from_pose(tx, ty, q) for tx, ty, q in poses])
    ps = [bg.AugParams(), bg.AugParams(flip=True, rects=[(30, 40, 50, 45)]), bg.AugParams(angle_deg=-14.0, flip=True),
          bg.AugParams(scale=0.951, rects=[(0, 0, 25, 25), (370, 680, 30, 40)]), bg.AugParams(angle_deg=9.5, scale=1.049),
          bg.AugParams(angle_deg=3.0, scale=0.97, flip=True, rects=[(100, 100, 50, 50)])]
    got = bg.map_ops.rasterize_maps_augmented(table, [0] * 6, Ts, ps)
    dense = bg.augment_grid(bg.rasterize_maps(table, [0] * 6, Ts), ps)
    assert torch.equal(got, dense)
    flat = co.flatten_map(md)
    got = got.cpu().numpy()
    for b, p in enumerate(ps):
        g = co.map_raster_flat(flat, Ts[b])
        if p.flip: g = co.flip_w(g)
        if p.angle_deg is not None: g = co.rotate(g, p.angle_deg)
        if p.scale is not None: g = co.scale(g, p.scale)
        g = np.ascontiguousarray(g).copy()
        if p.rects: co.dropout(g, p.rects)
        assert np.array_equal(got[b], g), (b, p)
