"""bevgen_b200 -- B200-native (sm_100a) BEV input generation for IntentNet / IntentNet-ViT.

Drop-in for the one hot path of Nadeem202020/VisionTransformer-Intention-Prediction: the
functions ``dataset.py`` imports from ``utils.py`` (reference dataset.py:21-23).  See DESIGN.md.
Import as ``bevgen_b200`` (alias module at the repo root).
"""
from . import constants, maps, pose, synthetic          # noqa: F401  (pure host modules)
from . import _native                                    # noqa: F401
from .bev_ops import (ALGO_ATOMIC, ALGO_AUTO, ALGO_MEMSET_ATOMIC, ALGO_TILED, PointColumns, create_intentnet_lidar_bev,   # noqa: F401
                      lidar_bev_from_ego_points, lidar_bev_from_sweeps, transform_points)

from . import augment, map_ops                           # noqa: F401,E402
from .augment import (AugParams, augment_bev, augment_grid, draw_aug_params, random_bev_dropout,   # noqa: F401,E402
                      random_flip_bev, random_rotate_bev, random_scale_bev)
from . import dataset_shim, sharding, sweep_io                     # noqa: F401,E402
from . import pipeline                                   # noqa: F401,E402
from .pipeline import BevGenerator, FrameBatch, PrefetchingGenerator, partition_for_augmentation           # noqa: F401,E402
from . import av2_dataset, sweep_store                    # noqa: F401,E402
from .av2_dataset import Av2BevDataset                     # noqa: F401,E402
from .sweep_store import StoreBatch, StoreFrame, SweepStore   # noqa: F401,E402
from .map_ops import MapTable, rasterize_map_ego_centric, rasterize_maps, world_to_bev_pixel   # noqa: F401,E402

__all__ = ["augment", "map_ops", "AugParams", "augment_bev", "augment_grid", "draw_aug_params", "random_bev_dropout",
           "random_flip_bev", "random_rotate_bev", "random_scale_bev", "MapTable", "rasterize_map_ego_centric",
           "rasterize_maps", "world_to_bev_pixel", "ALGO_MEMSET_ATOMIC",  "pipeline", "BevGenerator", "FrameBatch", "PrefetchingGenerator", "partition_for_augmentation", "constants", "maps", "pose", "synthetic", "transform_points", "create_intentnet_lidar_bev",
           "lidar_bev_from_sweeps", "lidar_bev_from_ego_points", "ALGO_TILED", "ALGO_ATOMIC", "ALGO_AUTO", "PointColumns", "sweep_store", "SweepStore", "StoreFrame", "StoreBatch", "av2_dataset", "Av2BevDataset"]
