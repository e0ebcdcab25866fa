"""Shared helpers of the test-suite (inputs, comparisons)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
MESH_DIRS = [os.environ.get("RTB_MESH_DIR", ""), os.path.join(ROOT, "oracle", "_ref", "data"), "/root/reference/TEST_Dungeonrun"]

# 3_walls.ply cannot be seen from the app's default camera: SURVEY.md section 8(d) C1
WALLS_CAMERA = dict(pos=(-200.0, 150.0, 60.0), look_at=(-307.8, 8.68, 2.22), up=(0.0, 0.0, 1.0))


def mesh_path(name):
    for d in MESH_DIRS:
        if d and os.path.exists(os.path.join(d, name)):
            return os.path.join(d, name)
    return None


def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


def cam12(W, H, pos=(0.0, 0.1, -1.0), look_at=(0.0, 0.1, 0.0), up=(0.0, 1.0, 0.0)):
    """12 camera scalars in the order of the reference constructor (WinMain.cpp:69-74)."""
    ar = np.float32(W) / np.float32(H)
    return [float(ar * np.float32(0.024)), 0.024, 0.055] + list(pos) + list(look_at) + list(up)


def cam_kwargs(W, H, pos=(0.0, 0.1, -1.0), look_at=(0.0, 0.1, 0.0), up=(0.0, 1.0, 0.0)):
    c = cam12(W, H, pos, look_at, up)
    return dict(f_w=c[0], f_h=c[1], fclen=c[2], pos=pos, look_at=look_at, up=up)


def channel_diff(a, b):
    sh = np.array([16, 8, 0], np.uint32)
    ca = ((np.asarray(a, np.uint32)[:, None] >> sh) & 0xff).astype(np.int32)
    cb = ((np.asarray(b, np.uint32)[:, None] >> sh) & 0xff).astype(np.int32)
    return np.abs(ca - cb).max(axis=1)


def hex32(arr):
    return ["%08x" % v for v in np.ascontiguousarray(arr, np.float32).view(np.uint32).ravel()]
