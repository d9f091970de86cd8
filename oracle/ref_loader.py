"""Loads the UNMODIFIED reference modules read-only from /root/reference (build container only).

Test infrastructure.  The reference's own ``import utils`` fails on Python <= 3.13 because two
annotations name symbols that were never imported (utils.py:295 with shapely absent,
heuristic_labeling.py:14); compiling the untouched source with lazy annotations
(``from __future__ import annotations`` semantics via the compiler flag) makes it importable
without editing or copying it.  Nothing is written to the reference directory.
"""
from __future__ import annotations

import __future__ as _future
import os
import sys
import types
from pathlib import Path

REF_DIR = Path(os.environ.get("BEVGEN_REFERENCE_DIR", "/root/reference"))


def available() -> bool:
    return (REF_DIR / "utils.py").is_file() and (REF_DIR / "constants.py").is_file()


def load(name: str) -> types.ModuleType:
    """Import reference module ``name`` (e.g. 'utils', 'constants', 'dataset')."""
    if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(str(REF_DIR)):
        return sys.modules[name]
    if not available():
        raise FileNotFoundError(f"reference not present at {REF_DIR}")
    sys.dont_write_bytecode = True
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    path = REF_DIR / f"{name}.py"
    mod = types.ModuleType(name)
    mod.__file__ = str(path)
    sys.modules[name] = mod
    code = compile(path.read_text(), str(path), "exec",
                   flags=_future.annotations.compiler_flag, dont_inherit=True)
    exec(code, mod.__dict__)
    return mod


def utils() -> types.ModuleType:
    load("constants")
    return load("utils")
