"""GPU integration: the batched device path equals the per-sample reference-signature path + collate, and the
reference's own dataset module can be rebound to the CUDA functions (no source edits)."""
import random
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bg():
    import bevgen_b200
    return bevgen_b200


def _samples(bg, md, n):
    out = []
    for i in range(n):
        f = bg.synthetic.frame_c3(40 + i, n=2_000, S=10, map_data=md)
        f.gt = {"boxes_xywha": np.array([[5.0 + i, -2.0, 2.0, 4.5, 0.2]], np.float32), "intentions": np.array([1 + (i % 4)])}
        out.append(f)
    return out


def test_batch_loader_equals_per_sample_path(bg, tmp_path):
    from oracle import c_oracle as co
    md = bg.synthetic.synthetic_map(seed=3, rows=9, cols=9, n_cross=4)
    path = bg.synthetic.write_map_json(md, tmp_path / "m.json")
    samples = _samples(bg, md, 5)
    table = bg.MapTable([bg.maps.parse_map_dict(md)])
    # eval mode: device batches == per-sample numpy functions stacked by collate_fn
    loader = bg.dataset_shim.BevBatchLoader(samples, 2, table, is_train=False)
    assert len(loader) == 3
    batches = list(loader)
    items = []
    for f in samples:
        pl = [bg.transform_points(f.sweep(s)[:, :3], f.rel_tf[s]) for s in range(10)]
        il = [f.sweep(s)[:, 3].copy() for s in range(10)]
        items.append({"lidar_bev": torch.from_numpy(bg.create_intentnet_lidar_bev(pl, il)).float(),
                      "map_bev": torch.from_numpy(bg.rasterize_map_ego_centric(path, f.pose_series())).float(), "gt": f.gt})
    items.insert(1, None)                                          # failed sample: dropped by collate (dataset.py:142)
    ref = bg.dataset_shim.collate_fn(items)
    got_l = torch.cat([b["lidar_bev"] for b in batches]).cpu()
    got_m = torch.cat([b["map_bev"] for b in batches]).cpu()
    assert got_l.shape == (5, 290, 400, 720) and got_l.is_contiguous()
    assert torch.equal(got_l, ref["lidar_bev"]) and torch.equal(got_m, ref["map_bev"])
    assert bg.dataset_shim.collate_fn([None, None]) is None
    # train mode: same draws as a per-sample augment_bev sequence with the same seed
    random.seed(77)
    tb = list(bg.dataset_shim.BevBatchLoader(samples, 5, table, is_train=True))[0]
    random.seed(77)
    for k, f in enumerate(samples):
        lo, mo, gt = bg.augment_bev(ref["lidar_bev"][k].numpy(), ref["map_bev"][k].numpy(),
                                    {"boxes_xywha": torch.from_numpy(f.gt["boxes_xywha"].copy()), "intentions": torch.from_numpy(f.gt["intentions"].copy())})
        assert np.array_equal(tb["lidar_bev"][k].cpu().numpy(), lo), k
        assert np.array_equal(tb["map_bev"][k].cpu().numpy(), mo), k
        assert torch.equal(tb["gt_list"][k]["boxes_xywha"], gt["boxes_xywha"]) and torch.equal(tb["gt_list"][k]["intentions"], gt["intentions"])


def test_install_rebinds_reference_style_module(bg):
    """a stand-in for the reference's `dataset` module namespace (from utils import ...)"""
    ds = types.ModuleType("dataset")
    for n in ("transform_points", "create_intentnet_lidar_bev", "rasterize_map_ego_centric", "augment_bev", "prepare_gt_for_frame"):
        setattr(ds, n, lambda *a, **k: (_ for _ in ()).throw(AssertionError("CPU path called")))
    done = bg.dataset_shim.install(ds)
    assert set(done) == {"dataset.transform_points", "dataset.create_intentnet_lidar_bev",
                         "dataset.rasterize_map_ego_centric", "dataset.augment_bev"}
    f = bg.synthetic.frame_c1(0, n=500)
    pts = ds.transform_points(f.sweep(0)[:, :3], f.rel_tf[0])
    bev = ds.create_intentnet_lidar_bev([pts], [f.sweep(0)[:, 3]])
    assert bev.shape == (290, 400, 720) and bev.any()


def test_deferred_transform_equals_eager_path_and_oracle(bg):
    """install()'s default: dataset.transform_points hands back DeferredPoints and create_intentnet_lidar_bev fuses the ten
    transforms with the binning -- same grid as the eager float64 route and as the oracle, same skip rules and errors; the
    array-like still reads as the transformed float64 array"""
    from oracle import c_oracle as co
    ds = types.ModuleType("dataset")
    for n in ("transform_points", "create_intentnet_lidar_bev"):
        setattr(ds, n, None)
    assert "dataset.transform_points" in bg.dataset_shim.install(ds)
    assert ds.transform_points is bg.bev_ops.transform_points_deferred
    f = bg.synthetic.quantize_f16(bg.synthetic.frame_c2(3, n=4_000))
    S = 10
    raw = [f.sweep(s)[:, :3].astype(np.float16) for s in range(S)]           # what dataset.py:320 reads from the feather file
    inten = [f.sweep(s)[:, 3].copy() for s in range(S)]
    raw = [np.asfortranarray(r) if s % 2 else r for s, r in enumerate(raw)]   # pandas' .values is column-major; plain arrays are not
    raw[6] = raw[6].astype(np.float32)
    lazy = [ds.transform_points(raw[s], f.rel_tf[s]) for s in range(S)]
    assert lazy[1].raw.flags.f_contiguous and lazy[0].raw.flags.c_contiguous and lazy[6].raw.dtype == np.float32
    eager = [bg.transform_points(raw[s], f.rel_tf[s]) for s in range(S)]
    assert all(isinstance(p, bg.bev_ops.DeferredPoints) and p.shape == e.shape and len(p) == len(e) and p.dtype == e.dtype
               for p, e in zip(lazy, eager))
    want = co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, S)
    got = ds.create_intentnet_lidar_bev(lazy, inten)
    assert got.dtype == np.float32 and got.shape == (290, 400, 720)
    assert np.array_equal(got, want) and np.array_equal(got, bg.create_intentnet_lidar_bev(eager, inten))
    assert all(p._value is None for p in lazy)                              # the fused route never materialised them
    assert np.array_equal(np.asarray(lazy[2]), eager[2]) and np.array_equal(lazy[2][5:9], eager[2][5:9])
    # missing sweeps, an empty one, fewer sweeps than expected (utils.py:71-76)
    pl, il = list(lazy[:7]), list(inten[:7])
    pl[1], il[1] = None, None
    pl[4], il[4] = ds.transform_points(raw[4][:0], f.rel_tf[4]), inten[4][:0]
    assert isinstance(pl[4], np.ndarray) and pl[4].shape == (0, 3)          # utils.py:29-30
    pe = [None if p is None else np.asarray(p) for p in pl]
    assert np.array_equal(ds.create_intentnet_lidar_bev(pl, il), bg.create_intentnet_lidar_bev(pe, il))
    # a mix of deferred and plain arrays takes the eager route (still exact)
    mixed = [eager[0]] + list(lazy[1:])
    assert np.array_equal(ds.create_intentnet_lidar_bev(mixed, inten), want)
    # float64 input is transformed right away; errors as the reference's
    assert isinstance(ds.transform_points(raw[0].astype(np.float64), f.rel_tf[0]), np.ndarray)
    with pytest.raises(IndexError):
        ds.create_intentnet_lidar_bev(lazy, inten, 9)                       # utils.py:102
    with pytest.raises(IndexError):
        ds.create_intentnet_lidar_bev(lazy[:1], [inten[0][:-1]])
    # the caller's array may change after the call: the deferred points keep their own copy
    r0 = raw[0].copy()
    p0 = ds.transform_points(r0, f.rel_tf[0])
    r0[:] = 0
    assert np.array_equal(np.asarray(p0), eager[0])


def test_partitioned_augmentation_equals_in_order_path(bg):
    """resampled frames first (no identity copies) == the in-order path, up to the returned permutation"""
    md = bg.synthetic.synthetic_map(seed=3, rows=9, cols=9, n_cross=4)
    table = bg.MapTable([bg.maps.parse_map_dict(md)])
    frames = _samples(bg, md, 6)
    aug = [bg.AugParams(flip=True), bg.AugParams(angle_deg=5.0), bg.AugParams(), bg.AugParams(scale=1.03, rects=[(10, 20, 30, 40)]),
           bg.AugParams(rects=[(100, 200, 50, 50)]), bg.AugParams(flip=True, angle_deg=-7.0, scale=0.97)]
    gen = bg.BevGenerator(table)
    ref = gen.generate(bg.FrameBatch.pack(frames, 10, map_ids=[0] * 6), aug)
    ref_l, ref_m = ref["lidar_bev"].clone(), ref["map_bev"].clone()
    f2, a2, order = bg.partition_for_augmentation(frames, aug)
    assert order == [1, 3, 5, 0, 2, 4]
    got = gen.generate(bg.FrameBatch.pack(f2, 10, map_ids=[0] * 6), a2)
    for k, i in enumerate(order):
        assert torch.equal(got["lidar_bev"][k], ref_l[i]), (k, i)
        assert torch.equal(got["map_bev"][k], ref_m[i]), (k, i)
    # loader with keep_order=False keeps grids and GT aligned
    random.seed(5)
    a = list(bg.dataset_shim.BevBatchLoader(frames, 6, table, is_train=True, keep_order=True))[0]
    random.seed(5)
    b = list(bg.dataset_shim.BevBatchLoader(frames, 6, table, is_train=True, keep_order=False))[0]
    used = set()
    for k in range(6):
        match = [i for i in range(6) if i not in used and torch.equal(b["lidar_bev"][k], a["lidar_bev"][i])
                 and torch.equal(b["gt_list"][k]["boxes_xywha"], a["gt_list"][i]["boxes_xywha"])]
        assert match, k
        used.add(match[0])
