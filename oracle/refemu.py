"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libref_emu.so (the reference's own
kernels compiled for the host by oracle/build_ref.py).  Used by tests/ to pin the C restatement
and to generate tests/golden/*, and by `bench.py --impl reference`.  Never imported by the product.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_emu.so")
# The reference's .cu files compiled by nvcc for sm_100a (oracle/build_ref.py, build_cuda): its real kernels on the GPU.
# "cuda_fmad" = the reference project's own code generation (the kernel to beat), "cuda_nofmad" = contraction off.
LIB_PATHS = {"emu": LIB_PATH, "cuda_fmad": os.path.join(HERE, "_ref", "libref_cuda_fmad.so"),
             "cuda_nofmad": os.path.join(HERE, "_ref", "libref_cuda_nofmad.so"),
             # the reference's HOST classes with its .cu files replaced by integration/rtb_seam.cpp over librtb.so
             "seam": os.path.join(HERE, "_ref", "libref_seam.so")}

# reference selectors (platform_common.h:16-21)
TRANSLATE_XYZ, TRANSLATE_X, TRANSLATE_Z, ROTATE_TRI_PY, ROTATE_TRI_NY = 30, 31, 32, 10, 11


class RefNode(C.Structure):
    _fields_ = [("left", C.c_longlong), ("right", C.c_longlong), ("tri", C.c_longlong), ("parent", C.c_longlong),
                ("cut_flag", C.c_int), ("is_leaf", C.c_int),
                ("x0", C.c_float), ("x1", C.c_float), ("y0", C.c_float), ("y1", C.c_float), ("z0", C.c_float),
                ("z1", C.c_float), ("s1", C.c_float), ("s2", C.c_float)]


NODE_DTYPE = np.dtype([("left", "<i8"), ("right", "<i8"), ("tri", "<i8"), ("parent", "<i8"), ("cut_flag", "<i4"),
                       ("is_leaf", "<i4"), ("x0", "<f4"), ("x1", "<f4"), ("y0", "<f4"), ("y1", "<f4"), ("z0", "<f4"),
                       ("z1", "<f4"), ("s1", "<f4"), ("s2", "<f4")])
assert NODE_DTYPE.itemsize == C.sizeof(RefNode)

_libs = {}


def available(impl="emu"):
    return os.path.exists(LIB_PATHS[impl])


def lib(impl="emu"):
    if impl not in _libs:
        L = C.CDLL(LIB_PATHS[impl])
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        for name in ("ref_num_tris", "ref_num_nodes"):
            getattr(L, name).restype = C.c_long
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("ref_get_points", "ref_get_nodes", "ref_get_camera", "ref_get_rays", "ref_get_matrix"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        L.ref_transform.restype = None
        L.ref_transform.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_render.restype = None
        L.ref_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_render_nocopy.restype = None
        L.ref_render_nocopy.argtypes = [C.c_void_p]
        L.ref_threads.restype = C.c_int
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_set_objects.argtypes = [C.c_int]
        _libs[impl] = L
    return _libs[impl]


class RefScene:
    """The reference app's scene: one camera, one mesh, one object (WinMain.cpp:69-156)."""

    def __init__(self, W, H, cam14, rgb=(0.1, 0.55, 0.2), ply_path=None, mode=0, points9=None, impl="emu", objects=1):
        """objects=2 replays WinMain.cpp:152-156 literally: two objects over the mesh, both registered with the camera,
        the first one transformed and rendered."""
        L = self.L = lib(impl)
        L.ref_set_objects(objects)
        self.W, self.H = W, H
        cam = np.zeros(14, np.float32)
        cam[:len(cam14)] = np.asarray(cam14, np.float32)
        col = np.asarray(rgb, np.float32)
        if ply_path is not None:
            self.h = L.ref_open(os.fsencode(ply_path), mode, None, 0, W, H, cam.ctypes.data, col.ctypes.data)
        else:
            pts = np.ascontiguousarray(points9, np.float32).reshape(-1, 9)
            self.h = L.ref_open(None, 0, pts.ctypes.data, pts.shape[0], W, H, cam.ctypes.data, col.ctypes.data)
        self.ntri = L.ref_num_tris(self.h)
        self.nnodes = L.ref_num_nodes(self.h)

    def points(self):
        out = np.empty((self.ntri, 9), np.float32)
        self.L.ref_get_points(self.h, out.ctypes.data)
        return out

    def nodes(self):
        out = np.zeros(self.nnodes, NODE_DTYPE)
        self.L.ref_get_nodes(self.h, out.ctypes.data)
        return out

    def camera(self):
        out = np.empty(18, np.float32)
        self.L.ref_get_camera(self.h, out.ctypes.data)
        return out.reshape(6, 3)  # n, v, u, n_mod, v_mod, u_mod

    def rays(self):
        out = np.empty((self.W * self.H, 3), np.float32)
        self.L.ref_get_rays(self.h, out.ctypes.data)
        return out

    def matrix(self):
        out = np.empty(12, np.float32)
        self.L.ref_get_matrix(self.h, out.ctypes.data)
        return out

    def transform(self, select, x, y, z, w):
        self.L.ref_transform(self.h, select, x, y, z, w)

    def render(self):
        ids = np.empty(self.W * self.H, np.int64)
        bgra = np.empty(self.W * self.H, np.uint32)
        self.L.ref_render(self.h, ids.ctypes.data, bgra.ctypes.data)
        return ids, bgra

    def render_nocopy(self):
        self.L.ref_render_nocopy(self.h)


def fnv1a64(buf):
    """64-bit FNV-1a over the raw little-endian bytes (the hash SURVEY.md Appendix C quotes)."""
    data = np.ascontiguousarray(buf).view(np.uint8).ravel()
    h = np.uint64(0xcbf29ce484222325)
    prime = np.uint64(0x100000001b3)
    # vectorising FNV is not possible (sequential); chunk through python ints for speed
    hv = int(h)
    p = int(prime)
    mask = (1 << 64) - 1
    for b in data.tobytes():
        hv = ((hv ^ b) * p) & mask
    return "%016x" % hv
