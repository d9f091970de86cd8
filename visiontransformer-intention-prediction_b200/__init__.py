"""bevgen_b200 -- B200-native (sm_100a) BEV input generation for IntentNet / IntentNet-ViT.

Drop-in for the one hot path of Nadeem202020/VisionTransformer-Intention-Prediction: the
functions ``dataset.py`` imports from ``utils.py`` (reference dataset.py:21-23).  See DESIGN.md.
Import as ``bevgen_b200`` (alias module at the repo root).
"""
from . import constants, maps, pose, synthetic          # noqa: F401  (pure host modules)
from . import _native                                    # noqa: F401
from .bev_ops import (ALGO_ATOMIC, ALGO_TILED, create_intentnet_lidar_bev, lidar_bev_from_ego_points,   # noqa: F401
                      lidar_bev_from_sweeps, transform_points)

__all__ = ["constants", "maps", "pose", "synthetic", "transform_points", "create_intentnet_lidar_bev",
           "lidar_bev_from_sweeps", "lidar_bev_from_ego_points", "ALGO_TILED", "ALGO_ATOMIC"]
