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
    assert L.bevgen_abi_version() == 1
    assert L.bevgen_error_string(-2).decode().startswith("bevgen")
    assert L.bevgen_lidar_workspace_bytes(1000, 1, 10) > 16_000
