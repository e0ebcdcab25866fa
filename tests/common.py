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


# ---- arithmetic edge-case scenes (shared by the CPU oracle-vs-reference test, the golden generator and the GPU test) ----
AXIS_CAMERA = dict(pos=(0.0, 0.1, -1.0), look_at=(0.0, 0.1, 0.0), up=(0.0, 1.0, 0.0))
GRID_CAMERA = dict(pos=(0.0, 0.5, -0.5), look_at=(0.0, 0.5, 0.5), up=(0.0, 1.0, 0.0))
EDGE_FRAMES = [(65, 49), (64, 48), (129, 1), (1, 97)]


def axis_aligned_soup():
    """Axis-aligned quads (zero-thickness boxes on all three axes), stacked copies (ties), slivers and degenerate triangles."""
    tris = []

    def quad(a, b, c, d):
        tris.append(a + b + c); tris.append(a + c + d)
    for z in (0.25, 0.25, 0.5):                       # two coincident walls and one behind them, facing the camera
        quad([-0.1, 0.0, z], [0.1, 0.0, z], [0.1, 0.2, z], [-0.1, 0.2, z])
    quad([-0.1, 0.0, 0.0], [-0.1, 0.2, 0.0], [-0.1, 0.2, 0.5], [-0.1, 0.0, 0.5])      # x = const wall (edge-on for the centre column)
    quad([0.0, 0.0, 0.0], [0.0, 0.2, 0.0], [0.0, 0.2, 0.2], [0.0, 0.0, 0.2])          # x = 0 wall: contains the centre rays' plane
    quad([-0.1, 0.1, 0.0], [0.1, 0.1, 0.0], [0.1, 0.1, 0.5], [-0.1, 0.1, 0.5])        # y = 0.1 floor: contains the centre row's plane
    tris.append([0.05, 0.05, 0.1, 0.05, 0.05, 0.1, 0.05, 0.05, 0.1])                  # a point
    tris.append([0.0, 0.0, 0.1, 0.05, 0.05, 0.1, 0.1, 0.1, 0.1])                      # collinear
    tris.append([0.02, 0.12, 0.2, 0.02 + 1e-7, 0.12, 0.2, 0.02, 0.12 + 1e-7, 0.2])    # a sliver far below a pixel
    return np.asarray(tris, np.float32)




def grid_mesh(n=8):
    """A regular grid of quads on exactly representable coordinates (rays along box planes and triangle edges)."""
    xs = np.arange(n + 1, dtype=np.float32) / 8 - 0.5
    tris = []
    for i in range(n):
        for j in range(n):
            a, b = [xs[i], xs[j] + 0.5, 0.5], [xs[i + 1], xs[j] + 0.5, 0.5]
            c, d = [xs[i + 1], xs[j + 1] + 0.5, 0.5], [xs[i], xs[j + 1] + 0.5, 0.5]
            tris.append(a + b + c); tris.append(a + c + d)
    return np.asarray(tris, np.float32)


def edge_script():
    """(select, x, y, z, w) ops applied one by one to the axis-aligned scene, a frame rendered after each."""
    h = float(np.float32(np.sqrt(0.5)))
    return [None] + [(10, 0.0, h, 0.0, h)] * 4 + [(32, 0.0, 0.0, 1.0, 0.25)] * 4


def golden_edge():
    with open(os.path.join(GOLDEN_DIR, "golden_edge.json")) as f:
        return json.load(f)


# ---- scene extension cases (shared by the golden generator, the CPU oracle test and the GPU test) ----------------------------
R_KEY = (10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)


def scene_cases():
    """name -> dict(W, H, objects=[(nu, [ops...])...], lights, shadows, sample_rate).  An object is a stand-in icosphere of
    20*nu^2 triangles moved by a list of (select, x, y, z, w) ops."""
    side = [(31, 1.0, 0.0, 0.0, 0.012)] * 10   # TRANSLATE_X: slides the second object sideways (overlapping silhouettes)
    near = [(32, 0.0, 0.0, 1.0, 0.01)] * 6     # TRANSLATE_Z: and towards the camera
    lights3 = [(2.0, 2.0, 2.0), (-2.0, 1.0, -2.0), (0.0, 3.0, -1.0)]
    front = [(-1.5, 1.0, -2.0)]                # a light on the camera's side: the visible surface is mostly lit
    return {
        "shadows_backlit": dict(W=200, H=120, objects=[(12, [R_KEY])], lights=[(2.0, 2.0, 2.0)], shadows=True, sample_rate=0),
        "shadows_frontlit": dict(W=200, H=120, objects=[(12, [R_KEY])], lights=front, shadows=True, sample_rate=0),
        "three_lights": dict(W=200, H=120, objects=[(12, [R_KEY])], lights=lights3, shadows=False, sample_rate=0),
        "three_lights_shadows": dict(W=200, H=120, objects=[(12, [R_KEY] * 2)], lights=lights3, shadows=True, sample_rate=0),
        "samples2": dict(W=160, H=90, objects=[(10, [])], lights=[(2.0, 2.0, 2.0)], shadows=False, sample_rate=2),
        "samples3_shadows": dict(W=97, H=61, objects=[(10, [R_KEY])], lights=front, shadows=True, sample_rate=3),
        "two_objects": dict(W=240, H=136, objects=[(12, [R_KEY]), (9, side)], lights=[(2.0, 2.0, 2.0)], shadows=False, sample_rate=0),
        "two_objects_ties": dict(W=160, H=90, objects=[(8, []), (8, [])], lights=[(2.0, 2.0, 2.0)], shadows=False, sample_rate=0),
        "three_objects_everything": dict(W=192, H=108, objects=[(10, [R_KEY]), (7, side + near), (6, [(31, 1.0, 0.0, 0.0, -0.012)] * 9)],
                                         lights=lights3, shadows=True, sample_rate=2),
    }


def build_scene_case(orc, mesh_fn, case):
    """Oracle scenes of a case (one orc.Scene per object, transforms applied) and the keyword arguments of orc.render_scene."""
    W, H = case["W"], case["H"]
    scenes = []
    for nu, ops in case["objects"]:
        s = orc.Scene(mesh_fn(nu), W, H, cam12(W, H))
        for op in ops:
            s.transform(*op)
        scenes.append(s)
    return scenes, dict(lights=case["lights"], shadows=case["shadows"], sample_rate=case["sample_rate"])


def golden_scene():
    with open(os.path.join(GOLDEN_DIR, "golden_scene.json")) as f:
        return json.load(f)
