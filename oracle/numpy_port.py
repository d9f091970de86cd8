"""numpy / OpenCV port of the reference's BEV generation -- the TIMED CPU BASELINE (test infrastructure).

Restates reference utils.py:27-33, :47-60, :62-106, :108-182, :394-498 and the pose prep of
dataset.py:285-301 / :323-339 with the SAME library calls in the same order (``@`` through
OpenBLAS, ``np.maximum.at`` per height bin, ``json.load`` per sample, ``cv2.fillPoly`` /
``cv2.polylines`` per lane, ``cv2.warpAffine`` / ``cv2.resize`` per channel), so its cost profile
is the reference's.  bench.py times it (``cpu_baseline.kind = "port"`` and ``--impl reference``);
tests check it against the C oracle and, in the build container, against the reference itself.

Third-party arithmetic it leans on (absent from the reference repo): numpy 2.x (pinned 2.0.2 by the
reference's requirements.txt:35, 2.3.5 here), opencv-python (pinned 4.11.0.86, 4.13.0 here),
scipy (pinned 1.13.1, 1.18.1 here).
"""
from __future__ import annotations

import json
import random

import numpy as np
from scipy.spatial.transform import Rotation as R

H, W, VOX, OFF_X, OFF_Y = 400, 720, 0.2, 360.0, 300.0
Z_LO, Z_HI, NZ, NMAP = -2.0, 3.8, 29, 9
MARK_CH = {"DASHED_WHITE": 6, "SOLID_WHITE": 7, "SOLID_YELLOW": 8}


def _cv2():
    import cv2
    return cv2


# ---------------------------------------------------------------- pose prep (dataset.py:285-301, :323-339)
def rel_transforms(current_pose7, sweep_poses7):
    cur = np.eye(4)
    cur[:3, :3] = R.from_quat(current_pose7[3:7]).as_matrix()
    cur[:3, 3] = current_pose7[:3]
    inv = np.linalg.inv(cur)
    out = []
    for p in sweep_poses7:
        m = np.eye(4)
        m[:3, :3] = R.from_quat(p[3:7]).as_matrix()
        m[:3, 3] = p[:3]
        out.append(inv @ m)
    return out


# ---------------------------------------------------------------- utils.py:27-33
def transform_points(points, T):
    if points.shape[0] == 0:
        return np.empty((0, 3), dtype=points.dtype)
    homo = np.hstack((points[:, :3], np.ones((points.shape[0], 1))))
    return (T @ homo.T).T[:, :3]


# ---------------------------------------------------------------- utils.py:62-106
def lidar_bev(points_list, intensity_list, num_sweeps=10):
    bev = np.zeros((NZ * num_sweeps, H, W), dtype=np.float32)
    for s in range(min(len(points_list), len(intensity_list))):
        pts, inten = points_list[s], intensity_list[s]
        if pts is None or inten is None or pts.shape[0] == 0:
            continue
        xe, ye, ze = pts[:, 0], pts[:, 1], pts[:, 2]
        col = np.floor(OFF_X + ye / VOX).astype(int)
        row = np.floor(OFF_Y - xe / VOX).astype(int)
        keep = (col >= 0) & (col < W) & (row >= 0) & (row < H) & (ze >= Z_LO) & (ze < Z_HI)
        col, row, zk, ik = col[keep], row[keep], ze[keep], inten[keep]
        if len(col) == 0:
            continue
        zb = np.clip(np.floor((zk - Z_LO) / (Z_HI - Z_LO) * NZ).astype(int), 0, NZ - 1)
        for h in range(NZ):
            sel = zb == h
            if np.any(sel):
                np.maximum.at(bev[s * NZ + h], (row[sel], col[sel]), ik[sel])
    return bev


# ---------------------------------------------------------------- utils.py:35-60, :108-182
def ego_matrix(xy, yaw):
    c, s = np.cos(-yaw), np.sin(-yaw)
    rot = np.array([[c, -s], [s, c]])
    T = np.eye(3)
    T[:2, :2] = rot
    T[:2, 2] = -rot @ xy
    return T


def to_pixels(xy, T):
    if xy.shape[0] == 0:
        return np.empty((0, 2), dtype=int)
    ego = (T @ np.hstack([xy, np.ones((xy.shape[0], 1))]).T).T[:, :2]
    px = OFF_X + ego[:, 1] / VOX
    py = OFF_Y - ego[:, 0] / VOX
    return np.round(np.vstack([px, py]).T).astype(int)


def map_raster(map_json_path, pose):
    """pose: mapping with tx_m, ty_m, qx, qy, qz, qw.  Parses the JSON per call, like the reference."""
    cv2 = _cv2()
    try:
        with open(map_json_path, "r") as fh:
            data = json.load(fh)
    except Exception:
        return np.zeros((NMAP, H, W), np.float32)
    try:
        yaw = R.from_quat([pose["qx"], pose["qy"], pose["qz"], pose["qw"]]).as_euler("xyz")[2]
    except ValueError:
        return np.zeros((NMAP, H, W), np.float32)
    T = ego_matrix(np.array([pose["tx_m"], pose["ty_m"]]), yaw)

    def pixels(lst):
        if not lst:
            return np.empty((0, 2), dtype=int)
        ok = [p for p in lst if isinstance(p, dict) and "x" in p and "y" in p]
        if not ok:
            return np.empty((0, 2), dtype=int)
        px = to_pixels(np.array([[p["x"], p["y"]] for p in ok]), T)
        m = (px[:, 0] >= 0) & (px[:, 0] < W) & (px[:, 1] >= 0) & (px[:, 1] < H)
        return px[m]

    img = np.zeros((NMAP, H, W), np.uint8)
    for lane in data.get("lane_segments", {}).values():
        lp, rp = pixels(lane.get("left_lane_boundary", [])), pixels(lane.get("right_lane_boundary", []))
        if len(lp) > 1 and len(rp) > 1:
            poly = np.vstack([lp, np.flipud(rp)])
            if poly.shape[0] >= 3:
                cv2.fillPoly(img[0], [poly.reshape(-1, 1, 2)], color=1)
                if lane.get("is_intersection", False):
                    cv2.fillPoly(img[4], [poly.reshape(-1, 1, 2)], color=1)
                if lane.get("lane_type") == "BUS":
                    cv2.fillPoly(img[5], [poly.reshape(-1, 1, 2)], color=1)
        if len(lp) > 1:
            cv2.polylines(img[1], [lp.reshape(-1, 1, 2)], isClosed=False, color=1, thickness=1)
        if len(rp) > 1:
            cv2.polylines(img[2], [rp.reshape(-1, 1, 2)], isClosed=False, color=1, thickness=1)
        lm, rm = lane.get("left_lane_mark_type", ""), lane.get("right_lane_mark_type", "")
        if lm in MARK_CH and len(lp) > 1:
            cv2.polylines(img[MARK_CH[lm]], [lp.reshape(-1, 1, 2)], False, 1, 1)
        if rm in MARK_CH and len(rp) > 1:
            cv2.polylines(img[MARK_CH[rm]], [rp.reshape(-1, 1, 2)], False, 1, 1)
    for cw in data.get("pedestrian_crossings", {}).values():
        poly = cw.get("polygon", [])
        if poly:
            px = pixels(poly)
            if len(px) >= 3:
                cv2.fillPoly(img[3], [px.reshape(-1, 1, 2)], color=1)
    return img.astype(np.float32)


# ---------------------------------------------------------------- utils.py:394-517 (image part)
def augment(lidar, mapb, flip=False, angle_deg=None, scale=None, rects=()):
    """Applies already-drawn decisions the way utils.py:500-517 does (copies first, utils.py:506)."""
    cv2 = _cv2()
    lidar, mapb = lidar.copy(), mapb.copy()
    if flip:
        lidar = np.ascontiguousarray(np.flip(lidar, axis=2))
        mapb = np.ascontiguousarray(np.flip(mapb, axis=2))
    if angle_deg is not None:
        M = cv2.getRotationMatrix2D((W / 2.0, H / 2.0), angle_deg, 1.0)
        rot = lambda g: np.stack([cv2.warpAffine(g[i], M, (W, H), flags=cv2.INTER_LINEAR,
                                                 borderMode=cv2.BORDER_CONSTANT, borderValue=0)
                                  for i in range(g.shape[0])], axis=0)
        lidar, mapb = rot(lidar), rot(mapb)
    if scale is not None:
        nh, nw = int(H * scale), int(W * scale)

        def scl(g):
            outs = []
            for i in range(g.shape[0]):
                r = cv2.resize(g[i], (nw, nh), interpolation=cv2.INTER_LINEAR)
                fin = np.zeros((H, W), dtype=g.dtype)
                if scale > 1.0:
                    h0, w0 = (nh - H) // 2, (nw - W) // 2
                    fin = r[h0:h0 + H, w0:w0 + W]
                else:
                    h0, w0 = (H - nh) // 2, (W - nw) // 2
                    fin[h0:h0 + nh, w0:w0 + nw] = r
                outs.append(fin)
            return np.stack(outs, axis=0)
        lidar, mapb = scl(lidar), scl(mapb)
    for (y0, x0, ph, pw) in rects:
        lidar[:, y0:y0 + ph, x0:x0 + pw] = 0.0
        mapb[:, y0:y0 + ph, x0:x0 + pw] = 0.0
    return lidar, mapb


# ---------------------------------------------------------------- one whole frame, as __getitem__ does it
def frame_lidar(points4, sweep_offsets, sweep_poses7, current_pose7, num_sweeps=10):
    """dataset.py:285-347: pose prep, per-sweep transform_points, create_intentnet_lidar_bev."""
    tfs = rel_transforms(current_pose7, sweep_poses7)
    pl, il = [], []
    for s in range(len(sweep_offsets) - 1):
        sw = points4[sweep_offsets[s]:sweep_offsets[s + 1]]
        if sw.shape[0] == 0:
            pl.append(None); il.append(None)
            continue
        pl.append(transform_points(sw[:, :3], tfs[s]))
        il.append(sw[:, 3].astype(np.float32))
    return lidar_bev(pl, il, num_sweeps)


def set_threads(n: int) -> None:
    """1 = pin OpenCV (and, via env set before numpy import, OpenBLAS) to one thread; 0 = library defaults."""
    _cv2().setNumThreads(n)
