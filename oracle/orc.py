"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/librtb_oracle.so, the C restatement of the
reference's ray-cast path (oracle/rtb_oracle.c).  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs only.  The product never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librtb_oracle.so")

NODE_DTYPE = np.dtype([("left", "<i8"), ("right", "<i8"), ("tri", "<i8"), ("parent", "<i8"), ("cut_flag", "<i4"),
                       ("is_leaf", "<i4"), ("x0", "<f4"), ("x1", "<f4"), ("y0", "<f4"), ("y1", "<f4"), ("z0", "<f4"),
                       ("z1", "<f4"), ("s1", "<f4"), ("s2", "<f4")])

# the reference app's literals (WinMain.cpp:69-74, 118-120)
DEFAULT_RGB = (0.1, 0.55, 0.2)


def default_camera(W, H):
    """f_w, f_h, fclen, pos, look-at, up of WinMain.cpp:69-74 for a W x H client area."""
    ar = np.float32(W) / np.float32(H)
    return [float(ar * np.float32(0.024)), 0.024, 0.055, 0.0, 0.1, -1.0, 0.0, 0.1, 0.0, 0.0, 1.0, 0.0]


def build(force=False):
    src = os.path.join(HERE, "rtb_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "librtb_oracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, ci, cl, cf = C.c_void_p, C.c_int, C.c_long, C.c_float
        L.orc_camera_basis.argtypes = [ci, ci, cf, cf, cf, vp, vp, vp, vp]
        L.orc_rays.argtypes = [vp, ci, ci, vp]
        L.orc_xform_init.argtypes = [vp, vp]
        L.orc_xform_apply.argtypes = [vp, ci, cf, cf, cf, cf]
        L.orc_xform_matrix.argtypes = [vp, vp]
        L.orc_xform_sizeof.restype = C.c_size_t
        L.orc_read_ply.argtypes = [C.c_char_p, ci, C.POINTER(vp), C.POINTER(cl)]
        L.orc_free.argtypes = [vp]
        L.orc_build_tree.argtypes = [vp, cl, vp]
        L.orc_scene_create.restype = vp
        L.orc_scene_create.argtypes = [vp, cl, vp, ci, vp, ci, ci, vp, vp]
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_render.argtypes = [vp, vp, ci, ci, vp, vp, vp]
        L.orc_render_bruteforce.argtypes = [vp, vp, ci, ci, vp, vp]
        L.orc_render_scene.argtypes = [ci, vp, vp, vp, ci, vp, ci, ci, ci, ci, vp, vp]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [vp, C.c_size_t]
        L.orc_threads.restype = ci
        L.orc_set_threads.argtypes = [ci]
        _lib = L
    return _lib


def fnv1a64(arr):
    a = np.ascontiguousarray(arr)
    return "%016x" % lib().orc_fnv1a64(a.ctypes.data, a.nbytes)


def read_ply(path, mode):
    p, n = C.c_void_p(), C.c_long()
    rc = lib().orc_read_ply(os.fsencode(path), mode, C.byref(p), C.byref(n))
    if rc != 0:
        raise IOError("orc_read_ply(%s, mode=%d) failed: %d" % (path, mode, rc))
    pts = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n.value, 9)).copy()
    lib().orc_free(p)
    return pts


def build_tree(points9):
    pts = np.ascontiguousarray(points9, np.float32).reshape(-1, 9)
    nodes = np.zeros(2 * pts.shape[0] - 1, NODE_DTYPE)
    rc = lib().orc_build_tree(pts.ctypes.data, pts.shape[0], nodes.ctypes.data)
    if rc != 0:
        raise ValueError("orc_build_tree failed: %d" % rc)
    return nodes


def camera_basis(W, H, cam12):
    c = np.asarray(cam12, np.float32)
    out = np.empty(18, np.float32)
    pos, la, up = c[3:6].copy(), c[6:9].copy(), c[9:12].copy()
    lib().orc_camera_basis(W, H, c[0], c[1], c[2], pos.ctypes.data, la.ctypes.data, up.ctypes.data, out.ctypes.data)
    return out


def rays(basis18, W, H):
    out = np.empty((W * H, 3), np.float32)
    lib().orc_rays(basis18.ctypes.data, W, H, out.ctypes.data)
    return out


class Xform:
    """Host transform recurrence of Object::transform (Camera.cu:254-335)."""

    def __init__(self, cam_pos):
        self.buf = np.zeros(lib().orc_xform_sizeof(), np.uint8)
        pos = np.asarray(cam_pos, np.float32)
        lib().orc_xform_init(self.buf.ctypes.data, pos.ctypes.data)

    def apply(self, select, x, y, z, w):
        lib().orc_xform_apply(self.buf.ctypes.data, select, x, y, z, w)

    def matrix(self):
        m = np.empty(12, np.float32)
        lib().orc_xform_matrix(self.buf.ctypes.data, m.ctypes.data)
        return m


class Scene:
    """points + tree + camera -> precomputed per-(camera, mesh) arrays, then render()."""

    def __init__(self, points9, W, H, cam12, rgb=DEFAULT_RGB, nodes=None):
        self.points = np.ascontiguousarray(points9, np.float32).reshape(-1, 9)
        self.n = self.points.shape[0]
        self.W, self.H = W, H
        self.cam = np.asarray(cam12, np.float32)
        self.nodes = build_tree(self.points) if nodes is None else nodes
        self.basis = camera_basis(W, H, self.cam)
        rad = np.ascontiguousarray(rgb, np.float32)
        per_tri = 1 if rad.size == 3 * self.n and self.n > 1 else 0
        pos = self.cam[3:6].copy()
        self.h = lib().orc_scene_create(self.points.ctypes.data, self.n, rad.ctypes.data, per_tri, self.nodes.ctypes.data,
                                        W, H, pos.ctypes.data, self.basis.ctypes.data)
        self.xform = Xform(pos)
        self.counters = np.zeros(3, np.uint64)

    def transform(self, select, x, y, z, w):
        self.xform.apply(select, x, y, z, w)

    def matrix(self):
        return self.xform.matrix()

    def render(self, m12=None, rows=None, want_ids=True, want_bgra=True):
        m = self.matrix() if m12 is None else np.ascontiguousarray(m12, np.float32)
        y0, y1 = (0, self.H) if rows is None else rows
        ids = np.full(self.W * self.H, -1, np.int64) if want_ids else None
        bgra = np.zeros(self.W * self.H, np.uint32) if want_bgra else None
        lib().orc_render(self.h, m.ctypes.data, y0, y1, ids.ctypes.data if want_ids else None,
                         bgra.ctypes.data if want_bgra else None, self.counters.ctypes.data)
        return ids, bgra

    def render_bruteforce(self, m12=None, rows=None):
        m = self.matrix() if m12 is None else np.ascontiguousarray(m12, np.float32)
        y0, y1 = (0, self.H) if rows is None else rows
        ids = np.full(self.W * self.H, -1, np.int64)
        dist = np.zeros(self.W * self.H, np.float32)
        lib().orc_render_bruteforce(self.h, m.ctypes.data, y0, y1, ids.ctypes.data, dist.ctypes.data)
        return ids, dist

    def close(self):
        if self.h:
            lib().orc_scene_destroy(self.h)
            self.h = None


DEFAULT_LIGHT = ((2.0, 2.0, 2.0),)  # Camera.cu:32


def render_scene(scenes, mats=None, lights=DEFAULT_LIGHT, shadows=False, sample_rate=0, rows=None):
    """The extension of SURVEY.md section 8(f) items 3-4 as defined in rtb_oracle.c (orc_render_scene): all `scenes`
    (Scene objects made for the SAME camera, in registration order) in one frame, closest hit wins, first registered wins
    ties; a light list, optional shadow rays, sample_rate^2 rays per pixel.  Returns (ids int64 with per-object id bases
    added, colours uint32).  With one scene, the default light, no shadows and sample_rate <= 1 it is Scene.render()."""
    s0 = scenes[0]
    m = np.ascontiguousarray([sc.matrix() for sc in scenes] if mats is None else mats, np.float32).reshape(len(scenes), 12)
    base = np.cumsum([0] + [sc.n for sc in scenes[:-1]]).astype(np.int64)
    hs = (C.c_void_p * len(scenes))(*[sc.h for sc in scenes])
    L3 = np.ascontiguousarray(lights, np.float32).reshape(-1, 3)
    y0, y1 = (0, s0.H) if rows is None else rows
    ids = np.full(s0.W * s0.H, -1, np.int64)
    bgra = np.zeros(s0.W * s0.H, np.uint32)
    lib().orc_render_scene(len(scenes), hs, m.ctypes.data, base.ctypes.data, L3.shape[0], L3.ctypes.data, int(bool(shadows)), int(sample_rate),
                           y0, y1, ids.ctypes.data, bgra.ctypes.data)
    return ids, bgra

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
