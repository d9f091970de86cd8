"""Generates tests/golden/fused_chain.npz: the reference's OWN outputs for the whole LiDAR chain on small seeded frames,

    transform_points (utils.py:27-33) -> create_intentnet_lidar_bev (utils.py:62-106) -> augment_bev (utils.py:500-517),

run from the UNMODIFIED reference (loaded read-only from /root/reference, see oracle/ref_loader.py) with Python's global
``random`` seeded per case, so that the CUDA path that builds the augmented grid straight from the point records
(bevgen_lidar_bev_augmented) is pinned against the reference itself and not only against the oracle.  cv2.resize runs with
IPP disabled, like the other fixtures.  Build container only:

    python -m oracle.make_golden_chain
"""
from __future__ import annotations

import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.make_golden import GOLD, pref, sparse  # noqa: E402


def main() -> None:
    import cv2
    import torch
    from oracle import ref_loader
    import bevgen_b200 as bg                      # only the pure-host synthetic generators / parameter draw are used

    u = ref_loader.utils()
    cv2.ipp.setUseIPP(False)
    S = 2
    f = bg.synthetic.frame_c2(frame_idx=31, n=2500, S=S)
    pl = [u.transform_points(f.sweep(s)[:, :3], f.rel_tf[s]) for s in range(S)]
    il = [f.sweep(s)[:, 3].astype(np.float32) for s in range(S)]
    bev = u.create_intentnet_lidar_bev(pl, il, S)
    mapb = np.zeros((1, 400, 720), np.float32)
    out = {"points4": f.points4, "sweep_offsets": f.sweep_offsets, "rel_tf": f.rel_tf}
    seeds, kinds = [], []
    for sd in range(200):
        random.seed(sd)
        p = bg.draw_aug_params(random)            # same draws, same order as augment_bev (tests/test_host_logic.py)
        kind = (p.angle_deg is not None, p.scale is not None, p.flip, bool(p.rects))
        if not (kind[0] or kind[1]) or kinds.count(kind) >= 1:
            continue                              # one case per (rotate, scale, flip, dropout) combination that resamples
        kinds.append(kind); seeds.append(sd)
        random.seed(sd)
        gt = {"boxes_xywha": torch.zeros((0, 5)), "intentions": torch.zeros((0,), dtype=torch.long)}
        lo, _, _ = u.augment_bev(bev, mapb, gt)
        out.update(pref(sparse(lo), f"s{sd}"))
    out["seeds"] = np.asarray(seeds)
    np.savez_compressed(GOLD / "fused_chain.npz", **out)
    print("seeds", seeds, "kinds", kinds, f"{(GOLD / 'fused_chain.npz').stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
