"""CPU: the C-ABI library loads and exports every symbol include/bevgen.h declares (no compute)."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported():
    import bevgen_b200
    from bevgen_b200 import _native
    _native.build()
    L = _native.lib()
    header = (ROOT / "include" / "bevgen.h").read_text()
    declared = set(re.findall(r"\b(bevgen_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_native.PROTOTYPES), declared ^ set(_native.PROTOTYPES)
    for name in declared:
        assert hasattr(L, name), name
    assert L.bevgen_abi_version() == 2
    assert L.bevgen_error_string(-2).decode().startswith("bevgen")
    assert L.bevgen_lidar_workspace_bytes(1000, 1, 10) > 16_000


def test_argument_validation_without_a_gpu():
    """Error behaviour of the C ABI that is decided before any CUDA call: these run on a CPU-only box."""
    import ctypes as C
    from bevgen_b200 import _native
    _native.build()
    L = _native.lib()
    INVALID, WORKSPACE = -1, -2
    one = C.c_void_p(16)            # a non-null, 16-byte aligned dummy pointer (never dereferenced on these paths)
    null = C.c_void_p(0)
    # LiDAR: no frames / null outputs / missing points / unknown algorithm / missing workspace for the tiled algorithm
    f = L.bevgen_lidar_bev_fused
    assert f(one, one, one, 0, 10, 0, 0, null, null, null, one, null, 0, 1, null) == INVALID
    assert f(one, one, one, 1, 10, 0, 0, null, null, null, null, null, 0, 1, null) == INVALID
    assert f(null, one, one, 1, 10, 5, 5, null, null, null, one, null, 0, 1, null) == INVALID
    assert f(one, one, one, 1, 10, 0, 0, null, null, null, one, null, 0, 99, null) == INVALID
    assert f(one, one, one, 1, 10, 0, 0, null, one, null, one, null, 0, 1, null) == INVALID      # rects without counts
    assert f(one, one, one, 1, 10, 0, 0, null, null, null, one, null, 0, 0, null) == WORKSPACE
    assert f(one, one, one, 1, 10, 0, 0, null, null, null, C.c_void_p(20), one, 1 << 40, 0, null) == INVALID   # misaligned out
    assert L.bevgen_lidar_bev_fused_f16(C.c_void_p(12), one, one, 1, 10, 5, 5, null, null, null, one, null, 0, 1, null) == INVALID
    assert L.bevgen_lidar_bev_ego_f64(null, null, one, 1, 10, 5, 5, null, null, null, one, null, 0, 1, null) == INVALID
    assert L.bevgen_lidar_workspace_bytes(-1, 1, 10) == 0 and L.bevgen_lidar_workspace_bytes(10, 0, 10) == 0
    need = L.bevgen_lidar_workspace_bytes(1_000_000, 1, 10)
    assert need >= 11_600 * 128 * 8 + 1_000_000 * 16        # tile bins + overflow list
    # map: no frames, missing workspace
    m = L.bevgen_map_raster
    assert m(one, one, one, one, one, one, 0, 1, 8, null, null, null, one, one, 1 << 30, null) == INVALID
    assert m(one, one, one, one, one, one, 1, 1, 8, null, null, null, one, null, 0, null) == WORKSPACE
    assert L.bevgen_map_workspace_bytes(2) == 2 * 9 * 400 * 24 * 4 and L.bevgen_map_workspace_bytes(0) == 0
    # transform / pixels: bad stride, empty input is a no-op
    T = (C.c_double * 16)(*([0.0] * 16))
    assert L.bevgen_transform_points_f32(one, 5, 2, T, one, null) == INVALID
    assert L.bevgen_transform_points_f64(null, 0, 3, T, null, null) == 0
    assert L.bevgen_world_to_bev_pixel(null, 0, T, null, null) == 0
    assert L.bevgen_world_to_bev_pixel(null, 3, T, null, null) == INVALID
    # augmentation: in-place is refused
    assert L.bevgen_augment(one, one, 1, 3, one, one, null, 0, null) == INVALID
    assert L.bevgen_augment(one, C.c_void_p(32), 0, 3, one, one, null, 0, null) == INVALID
    # generation fused with augmentation: host-side parameter screening, workspace sizing, argument checks
    UNSUPPORTED = -3
    from bevgen_b200 import augment as A
    ok = A.pack_params([A.AugParams(angle_deg=15.0, scale=0.95), A.AugParams(angle_deg=-15.0, scale=1.05), A.AugParams(flip=True)])
    bad = A.pack_params([A.AugParams(angle_deg=33.0)])
    pk, pb = C.c_void_p(ok.ctypes.data), C.c_void_p(bad.ctypes.data)
    assert L.bevgen_aug_fused_supported(pk, 3) == 1 and L.bevgen_aug_fused_supported(pb, 1) == 0
    assert L.bevgen_aug_fused_supported(null, 1) == 0
    base = L.bevgen_lidar_workspace_bytes(1000, 3, 10)
    assert L.bevgen_lidar_aug_workspace_bytes(1000, 3, 10, 0) == base
    assert L.bevgen_lidar_aug_workspace_bytes(1000, 3, 10, 2) >= base + 2 * 10 * 29 * 100 * 256 * 8     # per-tile record bins
    assert L.bevgen_lidar_aug_workspace_bytes(1000, 3, 10, 4) == 0                                       # more than B frames
    g = L.bevgen_lidar_bev_augmented
    assert g(one, 0, one, one, 3, 10, 0, 0, one, null, one, one, 1 << 40, null) == INVALID               # no host copy of the params
    assert g(null, 0, one, one, 3, 10, 5, 5, one, pk, one, one, 1 << 40, null) == INVALID                # points missing
    assert g(C.c_void_p(8), 0, one, one, 3, 10, 0, 0, one, pk, one, one, 1 << 40, null) == INVALID       # float4 points must be 16-byte aligned
    assert g(one, 0, one, one, 1, 10, 0, 0, one, pb, one, one, 1 << 40, null) == UNSUPPORTED             # refused, never approximated
    assert g(one, 0, one, one, 3, 10, 0, 0, one, pk, one, one, 1024, null) == WORKSPACE
    for code in (INVALID, WORKSPACE, -3):
        assert L.bevgen_error_string(code).decode().startswith("bevgen:")
