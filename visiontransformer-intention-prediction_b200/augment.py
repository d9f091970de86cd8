"""Train-time BEV augmentation (reference utils.py:394-517) on the GPU.

Host responsibilities (Python, as in the reference):
  * draw the random parameters from Python's global ``random`` in the reference's exact call
    order (utils.py:398, 421-422, 455-456, 488-494) so a seeded run makes the same decisions;
  * the handful of GT-box / label updates (utils.py:401-414, 441-448, 478-479);
  * cv2.getRotationMatrix2D + warpAffine's matrix inversion in float64 (libm cos/sin).
Device: one fused gather kernel per batch (csrc/augment.cu) for flip -> rotate -> scale -> dropout.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native
from .bev_ops import PointColumns, _dev, _device_of, _ptr, _require_cuda, _stream_ptr, to_numpy_fresh, workspace
from .constants import GRID_HEIGHT_PX, GRID_WIDTH_PX, INTENTIONS_MAP

_CV_PI = 3.1415926535897932384626433832795


@dataclasses.dataclass
class AugParams:
    """One frame's augmentation decisions (what the reference draws inside its four random_* calls)."""
    flip: bool = False
    angle_deg: Optional[float] = None          # None = no rotation
    scale: Optional[float] = None              # None = no scaling
    rects: List[Tuple[int, int, int, int]] = dataclasses.field(default_factory=list)   # (y0, x0, ph, pw)

    @property
    def needs_gather(self) -> bool:
        """True when a resampling pass is needed; flip / dropout alone can be fused into generation."""
        return self.angle_deg is not None or self.scale is not None


def draw_aug_params(rng=random, angle_range_deg=(-15.0, 15.0), scale_range=(0.95, 1.05), dropout_prob=0.1,
                    patch_size_range=(20, 50), num_patches_range=(1, 5)) -> AugParams:
    """Consumes ``rng`` exactly like augment_bev does (utils.py:508-511 -> :398, :421-422, :455-456, :488-494)."""
    p = AugParams()
    p.flip = rng.random() < 0.5
    if rng.random() < 0.5:
        p.angle_deg = rng.uniform(angle_range_deg[0], angle_range_deg[1])
    if rng.random() < 0.5:
        p.scale = rng.uniform(scale_range[0], scale_range[1])
    if rng.random() < dropout_prob:
        for _ in range(rng.randint(num_patches_range[0], num_patches_range[1])):
            ph = rng.randint(patch_size_range[0], patch_size_range[1])
            pw = rng.randint(patch_size_range[0], patch_size_range[1])
            y0 = rng.randint(0, max(0, GRID_HEIGHT_PX - ph))
            x0 = rng.randint(0, max(0, GRID_WIDTH_PX - pw))
            p.rects.append((y0, x0, ph, pw))
    return p


# ------------------------------------------------------------------------------- matrices (host, float64)
def rotation_matrix_2d(angle_deg: float) -> np.ndarray:
    """cv2.getRotationMatrix2D((360.0, 200.0), angle_deg, 1.0) (utils.py:425-428) restated in float64."""
    cx, cy = GRID_WIDTH_PX / 2.0, GRID_HEIGHT_PX / 2.0
    a = angle_deg * (_CV_PI / 180.0)
    alpha, beta = math.cos(a) * 1.0, math.sin(a) * 1.0
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy],
                     [-beta, alpha, beta * cx + (1 - alpha) * cy]], dtype=np.float64)


def invert_affine_2x3(M: np.ndarray) -> np.ndarray:
    """The inversion cv2.warpAffine performs on the forward matrix (no WARP_INVERSE_MAP), same op order."""
    m = [float(v) for v in np.asarray(M, np.float64).reshape(6)]
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return np.array(m, np.float64).reshape(2, 3)


def pack_params(params: Sequence[AugParams]) -> np.ndarray:
    """-> uint8 array holding ``len(params)`` C structs ``bevgen_aug_params`` (include/bevgen.h)."""
    arr = (_native.AugParams * len(params))()
    for i, p in enumerate(params):
        a = arr[i]
        if p.angle_deg is not None:
            mi = invert_affine_2x3(rotation_matrix_2d(p.angle_deg)).reshape(6)
            for k in range(6):
                a.minv[k] = mi[k]
            a.do_rotate = 1
        if p.scale is not None:
            a.scale = float(p.scale)
            a.do_scale = 1
        a.do_flip = 1 if p.flip else 0
        if len(p.rects) > 5:
            raise ValueError("at most 5 dropout rectangles (utils.py:485)")
        a.n_rects = len(p.rects)
        for j, (y0, x0, ph, pw) in enumerate(p.rects):
            a.rects[4 * j:4 * j + 4] = [int(y0), int(x0), int(ph), int(pw)]
    return np.frombuffer(bytes(arr), dtype=np.uint8).copy()


def fusion_arrays(params: Sequence[AugParams]):
    """(flip uint8 [B], rects int32 [B,5,4], counts int32 [B]) for the generator-fused flip/dropout of frames
    that need no resampling; frames that do get flip=0 / no rects here (the gather kernel applies them)."""
    B = len(params)
    flip = np.zeros(B, np.uint8)
    rects = np.zeros((B, 5, 4), np.int32)
    cnt = np.zeros(B, np.int32)
    for b, p in enumerate(params):
        if p.needs_gather:
            continue
        flip[b] = 1 if p.flip else 0
        cnt[b] = len(p.rects)
        for j, r in enumerate(p.rects):
            rects[b, j] = r
    return flip, rects, cnt


# ------------------------------------------------------------------------------- device entry
def augment_grid(src: torch.Tensor, params: Sequence[AugParams], out: Optional[torch.Tensor] = None,
                 two_pass: bool = True) -> torch.Tensor:
    """Apply per-frame ``params`` to ``src`` [B,C,400,720] (float32 CUDA); returns a new tensor (dst != src)."""
    _require_cuda()
    if src.dim() != 4 or src.shape[2:] != (GRID_HEIGHT_PX, GRID_WIDTH_PX) or src.dtype != torch.float32 or not src.is_cuda:
        raise ValueError("src must be a float32 CUDA tensor [B,C,400,720]")
    src = src.contiguous()
    B, Cc = int(src.shape[0]), int(src.shape[1])
    if len(params) != B:
        raise ValueError("one AugParams per frame")
    if out is None:
        out = torch.empty_like(src)
    elif (out.shape != src.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != src.device
          or out.data_ptr() == src.data_ptr()):
        raise ValueError("out must be a distinct contiguous float32 tensor of the same shape on the same device")
    hparams = pack_params(params)
    with torch.cuda.device(src.device):
        dparams = _dev(hparams, torch.uint8, src.device)
        # scratch for the frames that rotate AND scale (two 4-tap passes instead of one 16-tap gather)
        n_both = sum(1 for p in params if p.angle_deg is not None and p.scale is not None)
        ws, wbytes = None, 0
        if n_both and two_pass:
            wbytes = n_both * Cc * GRID_HEIGHT_PX * GRID_WIDTH_PX * 4
            ws = workspace.get(("augment", Cc), wbytes, src.device)
        rc = _native.lib().bevgen_augment(_ptr(src), _ptr(out), B, Cc, _ptr(dparams), C.c_void_p(hparams.ctypes.data),
                                          _ptr(ws), wbytes, _stream_ptr())
    _native.check(rc, "bevgen_augment")
    return out


def fused_supported(params: Sequence[AugParams], packed: Optional[np.ndarray] = None) -> bool:
    """True when every frame's parameters fit the fused generation+augmentation tile kernel (always true for the
    reference's own draws: |angle| <= 15 deg, 0.95 <= scale <= 1.05).  ``packed`` = pack_params(params) if at hand."""
    hp = pack_params(params) if packed is None else packed
    return bool(_native.lib().bevgen_aug_fused_supported(C.c_void_p(hp.ctypes.data), len(params)))


def lidar_bev_augmented(points4: torch.Tensor, sweep_offsets: torch.Tensor, rel_tf: torch.Tensor, num_frames: int,
                        num_sweeps: int, params: Sequence[AugParams], *, max_sweep_points: int,
                        out: Optional[torch.Tensor] = None, packed: Optional[np.ndarray] = None, sweep_start=None,
                        total_points: Optional[int] = None) -> torch.Tensor:
    """create_intentnet_lidar_bev (utils.py:62-106) + the image part of augment_bev (utils.py:508-511) for a batch, every
    output byte written once: frames that rotate / scale are built directly in augmented form from the binned point
    records (csrc/lidar_aug_tile.cuh) instead of generate -> read back -> resample -> write.  Device tensors in
    (points4 float32 or float16 [N,4], sweep_offsets int32 [B*S+1], rel_tf float64 [B*S,16]); returns float32
    [B, S*29, 400, 720].  Raises if a frame's parameters are unsupported (see ``fused_supported``)."""
    _require_cuda()
    B, S = int(num_frames), int(num_sweeps)
    if len(params) != B:
        raise ValueError("one AugParams per frame")
    cols = isinstance(points4, PointColumns)
    if not points4.is_cuda or (not cols and (points4.dim() != 2 or points4.shape[1] != 4
                                             or points4.dtype not in (torch.float32, torch.float16))):
        raise ValueError("points4 must be a CUDA tensor [N,4] float32 or float16, or a device-resident PointColumns")
    device = points4.device
    shape = (B, S * 29, GRID_HEIGHT_PX, GRID_WIDTH_PX)
    with torch.cuda.device(device):
        pts = points4 if cols else points4.contiguous()
        off = _dev(sweep_offsets, torch.int32, device)
        tf = _dev(rel_tf, torch.float64, device).reshape(-1, 16)
        if off.numel() != B * S + 1 or tf.shape[0] != B * S:
            raise ValueError("sweep_offsets must have B*S+1 entries and rel_tf B*S matrices")
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != device:
            raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape} on {device}")
        hparams = pack_params(params) if packed is None else packed
        dparams = _dev(hparams, torch.uint8, device)
        L = _native.lib()
        n = pts.n if cols else int(pts.shape[0])
        start = None
        if sweep_start is not None:            # sweep store form (bev_ops.lidar_bev_from_sweeps)
            if not cols or total_points is None:
                raise ValueError("sweep_start needs PointColumns and total_points")
            start = _dev(sweep_start, torch.int32, device)
            store_rows, n = n, int(total_points)
        wbytes = int(L.bevgen_lidar_aug_workspace_bytes(n, B, S, sum(1 for p in params if p.needs_gather)))
        ws = workspace.get("lidar", wbytes, device)
        if start is not None:
            rc = L.bevgen_lidar_bev_store_augmented_cols(*pts.ptrs(), store_rows, _ptr(start), _ptr(off), _ptr(tf), B, S, n,
                                                         int(max_sweep_points), _ptr(dparams), C.c_void_p(hparams.ctypes.data),
                                                         _ptr(out), _ptr(ws), wbytes, _stream_ptr())
        elif cols:
            rc = L.bevgen_lidar_bev_augmented_cols(*pts.ptrs(), _ptr(off), _ptr(tf), B, S, n, int(max_sweep_points),
                                                   _ptr(dparams), C.c_void_p(hparams.ctypes.data), _ptr(out), _ptr(ws), wbytes,
                                                   _stream_ptr())
        else:
            rc = L.bevgen_lidar_bev_augmented(_ptr(pts), 1 if pts.dtype == torch.float16 else 0, _ptr(off), _ptr(tf), B, S, n,
                                              int(max_sweep_points), _ptr(dparams), C.c_void_p(hparams.ctypes.data), _ptr(out),
                                              _ptr(ws), wbytes, _stream_ptr())
    _native.check(rc, "bevgen_lidar_bev_augmented")
    return out


# ------------------------------------------------------------------------------- GT (host, reference semantics)
def _wrap(a: np.ndarray) -> np.ndarray:
    return np.arctan2(np.sin(a), np.cos(a))


_FLIP_SWAP = {INTENTIONS_MAP["TURN_LEFT"]: INTENTIONS_MAP["TURN_RIGHT"],
              INTENTIONS_MAP["TURN_RIGHT"]: INTENTIONS_MAP["TURN_LEFT"],
              INTENTIONS_MAP["LEFT_CHANGE_LANE"]: INTENTIONS_MAP["RIGHT_CHANGE_LANE"],
              INTENTIONS_MAP["RIGHT_CHANGE_LANE"]: INTENTIONS_MAP["LEFT_CHANGE_LANE"]}


def apply_gt(p: AugParams, boxes: np.ndarray, intentions: np.ndarray):
    """GT side of utils.py:401-414 (flip), :441-448 (rotate about the ego origin) and :478-479 (scale),
    in place on ``boxes`` (K,5: cx,cy,w,l,yaw) and ``intentions`` (K,), in the reference's order."""
    if p.flip:
        if boxes.shape[0] > 0:
            boxes[:, 1] *= -1
            boxes[:, 4] *= -1
            boxes[:, 4] = _wrap(boxes[:, 4])
        if intentions.shape[0] > 0:
            orig = intentions.copy()
            for old, new in _FLIP_SWAP.items():
                intentions[orig == old] = new
    if p.angle_deg is not None and boxes.shape[0] > 0:
        ang = np.radians(p.angle_deg)
        cx, cy = boxes[:, 0].copy(), boxes[:, 1].copy()
        ca, sa = np.cos(ang), np.sin(ang)
        boxes[:, 0], boxes[:, 1] = cx * ca - cy * sa, cx * sa + cy * ca
        boxes[:, 4] += ang
        boxes[:, 4] = _wrap(boxes[:, 4])
    if p.scale is not None and boxes.shape[0] > 0:
        boxes[:, :4] *= p.scale
    return boxes, intentions


# ------------------------------------------------------------------------------- reference signatures (numpy)
def _apply_np(lidar_bev: np.ndarray, map_bev: np.ndarray, p: AugParams):
    dev = _device_of(lidar_bev, map_bev)
    lo = to_numpy_fresh(augment_grid(_dev(lidar_bev, torch.float32, dev)[None], [p])[0])
    mo = to_numpy_fresh(augment_grid(_dev(map_bev, torch.float32, dev)[None], [p])[0])
    return lo, mo


def random_flip_bev(lidar_bev, map_bev, gt_boxes_xywha, gt_intentions):
    """Drop-in for utils.py:394-415."""
    p = AugParams(flip=random.random() < 0.5)
    if p.flip:
        lidar_bev, map_bev = _apply_np(lidar_bev, map_bev, p)
        apply_gt(p, gt_boxes_xywha, gt_intentions)
    return lidar_bev, map_bev, gt_boxes_xywha, gt_intentions


def random_rotate_bev(lidar_bev, map_bev, gt_boxes_xywha, angle_range_deg=(-15.0, 15.0)):
    """Drop-in for utils.py:417-449."""
    if random.random() < 0.5:
        p = AugParams(angle_deg=random.uniform(angle_range_deg[0], angle_range_deg[1]))
        lidar_bev, map_bev = _apply_np(lidar_bev, map_bev, p)
        apply_gt(p, gt_boxes_xywha, np.zeros((0,), np.int64))
    return lidar_bev, map_bev, gt_boxes_xywha


def random_scale_bev(lidar_bev, map_bev, gt_boxes_xywha, scale_range=(0.95, 1.05)):
    """Drop-in for utils.py:451-480."""
    if random.random() < 0.5:
        p = AugParams(scale=random.uniform(scale_range[0], scale_range[1]))
        lidar_bev, map_bev = _apply_np(lidar_bev, map_bev, p)
        apply_gt(p, gt_boxes_xywha, np.zeros((0,), np.int64))
    return lidar_bev, map_bev, gt_boxes_xywha


def random_bev_dropout(lidar_bev, map_bev, dropout_prob=0.1, patch_size_range=(20, 50), num_patches_range=(1, 5)):
    """Drop-in for utils.py:482-498 (mutates and returns its inputs, like the reference)."""
    if random.random() < dropout_prob:
        p = AugParams()
        for _ in range(random.randint(num_patches_range[0], num_patches_range[1])):
            ph = random.randint(patch_size_range[0], patch_size_range[1])
            pw = random.randint(patch_size_range[0], patch_size_range[1])
            y0 = random.randint(0, max(0, GRID_HEIGHT_PX - ph))
            x0 = random.randint(0, max(0, GRID_WIDTH_PX - pw))
            p.rects.append((y0, x0, ph, pw))
        if isinstance(lidar_bev, torch.Tensor):
            lidar_bev.copy_(augment_grid(lidar_bev[None], [p])[0])
            map_bev.copy_(augment_grid(map_bev[None], [p])[0])
        else:
            lo, mo = _apply_np(lidar_bev, map_bev, p)
            lidar_bev[...] = lo          # in place, as utils.py:496-497
            map_bev[...] = mo
    return lidar_bev, map_bev


def augment_bev(lidar_bev, map_bev, gt_dict):
    """Drop-in for utils.py:500-517: numpy (or CUDA tensor) grids in, same kind out, GT dict of tensors.

    Inputs are never modified (the reference copies at utils.py:503-506).  With CUDA tensors in, the
    result stays on the device (no 344 MB host round trip)."""
    gb, gi = gt_dict["boxes_xywha"], gt_dict["intentions"]
    boxes = gb.clone().numpy() if isinstance(gb, torch.Tensor) else gb.copy()
    intents = gi.clone().numpy() if isinstance(gi, torch.Tensor) else gi.copy()
    p = draw_aug_params(random)
    apply_gt(p, boxes, intents)
    on_device = isinstance(lidar_bev, torch.Tensor)
    dev = _device_of(lidar_bev, map_bev)
    ld = lidar_bev if on_device else _dev(lidar_bev, torch.float32, dev)
    md = map_bev if isinstance(map_bev, torch.Tensor) else _dev(map_bev, torch.float32, dev)
    lo = augment_grid(ld[None], [p])[0]
    mo = augment_grid(md[None], [p])[0]
    if not on_device:
        lo, mo = to_numpy_fresh(lo), to_numpy_fresh(mo)
    gt = {"boxes_xywha": torch.from_numpy(np.asarray(boxes)).float(), "intentions": torch.from_numpy(np.asarray(intents)).long()}
    return lo, mo, gt
