"""Helpers to read the committed fixtures under tests/golden/ (made by oracle/make_golden.py from the reference)."""
import hashlib
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    return np.load(GOLD / name, allow_pickle=False)


def dense(z, prefix):
    shape = tuple(int(v) for v in z[f"{prefix}_shape"])
    out = np.zeros(int(np.prod(shape)), np.float32)
    out[z[f"{prefix}_idx"]] = z[f"{prefix}_val"]
    out = out.reshape(shape)
    assert hashlib.sha256(out.tobytes()).digest() == bytes(z[f"{prefix}_sha256"]), "fixture self-check failed"
    return out


def aug_inputs(z):
    lid = np.zeros(2 * 400 * 720, np.float32)
    lid[z["lidar_in_idx"]] = z["lidar_in_val"]
    mp = np.zeros(400 * 720, np.float32)
    mp[z["map_in_idx"]] = 1.0
    return lid.reshape(2, 400, 720), mp.reshape(1, 400, 720)
