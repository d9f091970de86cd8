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
