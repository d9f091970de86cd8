"""BEV grid constants of the path (reference constants.py:28-47), mirrored for the host side.

The CUDA library bakes the same values in as ``constexpr`` (csrc/bevgen_common.cuh); the
loader (``_native.py``) asserts at load time that both sides agree.
"""
GRID_HEIGHT_PX, GRID_WIDTH_PX = 400, 720          # constants.py:28
VOXEL_SIZE_M = 0.2                                # constants.py:29
BEV_PIXEL_OFFSET_X = GRID_WIDTH_PX / 2.0          # constants.py:38 -> 360.0
BEV_PIXEL_OFFSET_Y = GRID_HEIGHT_PX * 3.0 / 4.0   # constants.py:39 -> 300.0
Z_MIN, Z_MAX = -2.0, 3.8                          # constants.py:42
LIDAR_HEIGHT_CHANNELS = 29                        # constants.py:43
LIDAR_SWEEPS = 10                                 # constants.py:44
LIDAR_TOTAL_CHANNELS = LIDAR_HEIGHT_CHANNELS * LIDAR_SWEEPS   # constants.py:45 -> 290
MAP_CHANNELS = 9                                  # constants.py:47

# intention ids touched by the flip augmentation's label swap (constants.py:50-60, utils.py:406-411)
INTENTIONS_MAP = {
    "KEEP_LANE": 0, "TURN_LEFT": 1, "TURN_RIGHT": 2, "LEFT_CHANGE_LANE": 3,
    "RIGHT_CHANGE_LANE": 4, "STOPPING_STOPPED": 5, "PARKED": 6, "OTHER": 7,
}

# annotation categories kept as ground truth (constants.py:80-84; used by the Dataset mirror's default GT, utils.py:190-194)
VEHICLE_CATEGORIES = {
    "REGULAR_VEHICLE", "LARGE_VEHICLE", "BUS", "BOX_TRUCK", "TRUCK", "MOTORCYCLE", "SCHOOL_BUS", "ARTICULATED_BUS",
    "VEHICULAR_TRAILER", "TRUCK_CAB", "BICYCLE", "BICYCLIST", "MOTORCYCLIST",
}

LIDAR_FRAME_BYTES = LIDAR_TOTAL_CHANNELS * GRID_HEIGHT_PX * GRID_WIDTH_PX * 4   # 334 080 000
MAP_FRAME_BYTES = MAP_CHANNELS * GRID_HEIGHT_PX * GRID_WIDTH_PX * 4             # 10 368 000
