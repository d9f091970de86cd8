"""Host-side pose preparation.

These matrices are tiny (1 + S poses per frame) and stay on the host, produced by the
*same* numpy / scipy calls the reference makes, so the float64 bits handed to the kernels
equal the reference's:

* ``ego_se3_world`` / ``sweep_rel_transforms``  -- dataset.py:285-301 and dataset.py:323-339
* ``ego_yaw`` / ``ego_centric_matrix``          -- utils.py:121-129 and utils.py:35-45
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as R


def se3_from_pose(t_xyz, q_xyzw) -> np.ndarray:
    """world_SE3_ego 4x4 float64 from a translation and a scalar-last quaternion
    (dataset.py:293-300 / :331-338).  Raises ValueError on an invalid quaternion, as scipy does."""
    m = np.eye(4)
    m[:3, :3] = R.from_quat(q_xyzw).as_matrix()
    m[:3, 3] = [t_xyz[0], t_xyz[1], t_xyz[2]]
    return m


def ego_se3_world(current_pose7) -> np.ndarray:
    """inverse of the current ego pose (dataset.py:301).  pose7 = (tx,ty,tz,qx,qy,qz,qw)."""
    p = np.asarray(current_pose7, dtype=np.float64)
    return np.linalg.inv(se3_from_pose(p[:3], p[3:7]))


def sweep_rel_transforms(current_pose7, sweep_poses7) -> np.ndarray:
    """rel_tf[s] = ego_SE3_world @ sweep_s_SE3 (dataset.py:339); float64 [S,4,4]."""
    inv_cur = ego_se3_world(current_pose7)
    sp = np.asarray(sweep_poses7, dtype=np.float64).reshape(-1, 7)
    out = np.empty((sp.shape[0], 4, 4), np.float64)
    for s in range(sp.shape[0]):
        out[s] = inv_cur @ se3_from_pose(sp[s, :3], sp[s, 3:7])
    return out


def ego_yaw(q_xyzw) -> float:
    """utils.py:124 -- z Euler angle of the ego quaternion (extrinsic xyz)."""
    return R.from_quat(q_xyzw).as_euler("xyz")[2]


def ego_centric_matrix(ego_xy, yaw) -> np.ndarray:
    """utils.py:35-45 -- 3x3 float64 world -> ego-centric BEV frame (same numpy calls)."""
    c, s = np.cos(-yaw), np.sin(-yaw)
    rot = np.array([[c, -s], [s, c]])
    t = -rot @ np.asarray(ego_xy)
    m = np.eye(3)
    m[:2, :2] = rot
    m[:2, 2] = t
    return m


def map_matrix_from_pose(tx, ty, q_xyzw) -> np.ndarray:
    """utils.py:121-129: the matrix rasterize_map_ego_centric builds from the ego pose row."""
    return ego_centric_matrix(np.array([tx, ty]), ego_yaw(q_xyzw))
