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
