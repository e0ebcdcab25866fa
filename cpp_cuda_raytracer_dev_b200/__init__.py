"""B200-native ray-cast path of ams3878/cpp_cuda_raytracer_dev -- Python view of the C ABI.

The product is `librtb.so` (include/rtb.h): hand-written CUDA for sm_100a behind the reference's
operator surface.  This module is a thin ctypes binding used by tests/, bench.py and
__graft_entry__.py; the classes keep the reference's names and call order
(WinMain.cpp:69-237): `read_ply` -> `Trixel(points)` -> `set_sorted_voxels()/create_kd()` ->
`Camera(...)` -> `Object(trixel)` -> `camera.add_object(obj)` -> `obj.transform(...)` ->
`obj.render(camera)` -> `camera.color_pixels(PHONG_COLOR_TAG)`.

There is no CPU fallback: if the shared library is missing the import fails, and every call that
needs the GPU raises RtbError when no device is usable.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTB_LIB") or os.path.join(HERE, "librtb.so")  # RTB_LIB: development override

SET_COLOR_TAG, PHONG_COLOR_TAG = 1, 2  # Camera.h:13-14
TRANSLATE_XYZ, TRANSLATE_X, TRANSLATE_Z, ROTATE_TRI_PY, ROTATE_TRI_NY = 30, 31, 32, 10, 11  # platform_common.h:16-21
RENDER_DEFAULT, RENDER_NO_CULL, RENDER_COUNTERS, RENDER_TILE_MAJOR, RENDER_PUSH_PREFILLED = 0, 1, 2, 4, 8
R_KEY_QUAT = (0.0, 0.09950371902099893, 0.0, 0.9950371902099893)  # WinMain.cpp:187
T_KEY_QUAT = (0.0, -0.09950371902099893, 0.0, 0.9950371902099893)  # WinMain.cpp:207
DEFAULT_RGB = (0.1, 0.55, 0.2)  # WinMain.cpp:118-120
TILE = 32

EXPORTS = [
    "rtb_last_error", "rtb_version", "rtb_device_count", "rtb_set_device", "rtb_set_knob", "rtb_read_ply", "rtb_free", "rtb_write_ply",
    "rtb_mesh_geodesic", "rtb_mesh_create", "rtb_mesh_build_tree", "rtb_mesh_build_tree_on", "rtb_mesh_num_triangles", "rtb_mesh_num_nodes",
    "rtb_mesh_get_tree", "rtb_mesh_save_tree", "rtb_mesh_load_tree", "rtb_write_frame", "rtb_mesh_build_seconds", "rtb_mesh_destroy", "rtb_camera_create", "rtb_camera_get_basis",
    "rtb_camera_add_object", "rtb_camera_color_pixels", "rtb_camera_host_color", "rtb_camera_host_ids",
    "rtb_camera_counters", "rtb_camera_counters_ex", "rtb_camera_destroy", "rtb_object_create", "rtb_object_transform", "rtb_object_get_matrix",
    "rtb_object_set_matrix", "rtb_object_destroy", "rtb_object_render", "rtb_render_frame", "rtb_camera_set_lights", "rtb_camera_set_shadows", "rtb_camera_set_sample_rate", "rtb_camera_object_id_base",
    "rtb_camera_render_scene", "rtb_camera_render_scene_device_async", "rtb_render_sweep",
    "rtb_render_frames_device_async", "rtb_render_frames_push_async", "rtb_render_frames_push_striped_async", "rtb_fill_frames_device_async", "rtb_peer_alloc", "rtb_peer_free", "rtb_peer_export", "rtb_peer_open", "rtb_peer_close", "rtb_peer_read", "rtb_object_transform_host", "rtb_transform_sequence_host", "rtb_device_props", "rtb_launch_count", "rtb_tile_major_elements", "rtb_compose_tiles_device_async", "rtb_selftest_exact", "rtb_measure_l2_read_bandwidth", "rtb_measure_host_fill_bandwidth", "rtb_host_alloc", "rtb_host_free",
]


class RtbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("librtb.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the ray-cast path)")
    L = C.CDLL(LIB_PATH)
    vp, ci, cf, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    L.rtb_last_error.restype = C.c_char_p
    L.rtb_version.restype = C.c_char_p
    L.rtb_set_device.argtypes = [ci]
    L.rtb_set_knob.argtypes = [C.c_char_p, ci]
    L.rtb_read_ply.argtypes = [C.c_char_p, ci, C.POINTER(vp), C.POINTER(C.c_uint32)]
    L.rtb_free.argtypes = [vp]
    L.rtb_free.restype = None
    L.rtb_write_ply.argtypes = [C.c_char_p, vp, C.c_uint32]
    L.rtb_mesh_geodesic.argtypes = [ci, cf, vp, cf, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32)]
    L.rtb_mesh_create.argtypes = [vp, i64, vp, vp, C.POINTER(vp)]
    L.rtb_mesh_build_tree.argtypes = [vp]
    L.rtb_mesh_build_tree_on.argtypes = [vp, ci]
    L.rtb_mesh_save_tree.argtypes = [vp, C.c_char_p]
    L.rtb_mesh_load_tree.argtypes = [vp, C.c_char_p]
    L.rtb_write_frame.argtypes = [C.c_char_p, vp, C.c_int32, C.c_int32]
    L.rtb_mesh_num_triangles.argtypes = [vp]
    L.rtb_mesh_num_triangles.restype = i64
    L.rtb_mesh_num_nodes.argtypes = [vp]
    L.rtb_mesh_num_nodes.restype = i64
    L.rtb_mesh_get_tree.argtypes = [vp] * 8
    L.rtb_mesh_build_seconds.argtypes = [vp, vp]
    L.rtb_mesh_destroy.argtypes = [vp]
    L.rtb_mesh_destroy.restype = None
    L.rtb_camera_create.argtypes = [C.c_int32, C.c_int32, cf, cf, cf, vp, vp, vp, C.POINTER(vp)]
    L.rtb_camera_get_basis.argtypes = [vp, vp]
    L.rtb_camera_add_object.argtypes = [vp, vp]
    L.rtb_camera_color_pixels.argtypes = [vp, C.c_uint8]
    L.rtb_camera_host_color.argtypes = [vp]
    L.rtb_camera_host_color.restype = vp
    L.rtb_camera_host_ids.argtypes = [vp]
    L.rtb_camera_host_ids.restype = vp
    L.rtb_camera_counters.argtypes = [vp, vp, ci]
    L.rtb_camera_counters_ex.argtypes = [vp, vp, ci]
    L.rtb_camera_destroy.argtypes = [vp]
    L.rtb_camera_destroy.restype = None
    L.rtb_object_create.argtypes = [vp, C.POINTER(vp)]
    L.rtb_object_transform.argtypes = [vp, vp, C.c_uint8]
    L.rtb_object_transform_host.argtypes = [vp, vp, C.c_uint8, vp]
    L.rtb_object_get_matrix.argtypes = [vp, vp]
    L.rtb_transform_sequence_host.argtypes = [vp, C.c_int32, vp, vp]
    L.rtb_object_set_matrix.argtypes = [vp, vp]
    L.rtb_object_destroy.argtypes = [vp]
    L.rtb_object_destroy.restype = None
    L.rtb_object_render.argtypes = [vp, vp, C.c_uint32]
    L.rtb_render_frame.argtypes = [vp, vp, C.c_uint32]
    L.rtb_camera_set_lights.argtypes = [vp, C.c_int32, vp]
    L.rtb_camera_set_shadows.argtypes = [vp, C.c_int32]
    L.rtb_camera_set_sample_rate.argtypes = [vp, C.c_int32]
    L.rtb_camera_object_id_base.argtypes = [vp, vp]
    L.rtb_camera_object_id_base.restype = C.c_int64
    L.rtb_camera_render_scene.argtypes = [vp, C.c_uint32]
    L.rtb_camera_render_scene_device_async.argtypes = [vp, C.c_uint32, vp, vp, vp]
    L.rtb_render_sweep.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, C.c_uint32, vp, vp]
    L.rtb_render_frames_device_async.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_uint32, vp, vp, vp]
    L.rtb_device_props.argtypes = [vp]
    L.rtb_launch_count.restype = C.c_uint64
    L.rtb_selftest_exact.argtypes = [C.c_uint64, C.c_int64, vp]
    L.rtb_render_frames_push_async.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_uint32, vp, vp, vp]
    L.rtb_render_frames_push_striped_async.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_uint32, C.c_int32, vp, vp, vp]
    L.rtb_fill_frames_device_async.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.rtb_measure_l2_read_bandwidth.argtypes = [C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    L.rtb_measure_host_fill_bandwidth.argtypes = [C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    L.rtb_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.rtb_host_free.argtypes = [vp]
    L.rtb_peer_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.rtb_peer_free.argtypes = [vp]
    L.rtb_peer_export.argtypes = [vp, vp]
    L.rtb_peer_open.argtypes = [vp, C.POINTER(vp)]
    L.rtb_peer_close.argtypes = [vp]
    L.rtb_peer_read.argtypes = [vp, vp, C.c_size_t]
    L.rtb_tile_major_elements.restype = C.c_int64
    L.rtb_tile_major_elements.argtypes = [vp, C.c_int32]
    L.rtb_compose_tiles_device_async.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp]
    return L


lib = _load()


def _check(rc, what=""):
    if rc != 0:
        raise RtbError("%s failed (status %d): %s" % (what, rc, lib.rtb_last_error().decode(errors="replace")))


def device_count():
    return lib.rtb_device_count()


def set_device(i):
    _check(lib.rtb_set_device(i), "rtb_set_device")


def set_knob(name, value):
    """Scheduling knobs of the render kernel (include/rtb.h: rtb_set_knob); process-wide."""
    _check(lib.rtb_set_knob(name.encode(), int(value)), "rtb_set_knob")


def selftest_exact(seed, count):
    out = np.zeros(4, np.uint64)
    _check(lib.rtb_selftest_exact(seed, count, out.ctypes.data), "rtb_selftest_exact")
    return dict(zip(["rsqrt_mismatch", "rcp_mismatch", "decision_mismatch", "decidable"], out.tolist()))


def launch_count():
    return int(lib.rtb_launch_count())


def device_props():
    out = np.zeros(7, np.int64)
    _check(lib.rtb_device_props(out.ctypes.data), "rtb_device_props")
    keys = ["sm_count", "l2_bytes", "persisting_l2_max_bytes", "sm_clock_khz", "mem_clock_khz", "mem_bus_bits", "cc"]
    return dict(zip(keys, out.tolist()))


def default_camera_args(W, H):
    """f_w, f_h, fclen, pos, look-at, up of WinMain.cpp:69-74 for a W x H client area."""
    ar = np.float32(W) / np.float32(H)
    return dict(f_w=float(ar * np.float32(0.024)), f_h=0.024, fclen=0.055, pos=(0.0, 0.1, -1.0), look_at=(0.0, 0.1, 0.0),
                up=(0.0, 1.0, 0.0))


def read_ply(file_name, mode):
    """read_ply.cpp:13 -- returns the (n, 9) float32 triangle soup the reference loader produces."""
    p, n = C.c_void_p(), C.c_uint32()
    _check(lib.rtb_read_ply(os.fsencode(file_name), mode, C.byref(p), C.byref(n)), "rtb_read_ply")
    pts = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n.value, 9)).copy()
    lib.rtb_free(p)
    return pts


def write_ply(file_name, points9):
    pts = np.ascontiguousarray(points9, np.float32).reshape(-1, 9)
    _check(lib.rtb_write_ply(os.fsencode(file_name), pts.ctypes.data, pts.shape[0]), "rtb_write_ply")


def geodesic_mesh(nu, radius=0.08, center=(0.0, 0.1, 0.0), displacement=0.05, seed=1234):
    """Displaced geodesic icosphere with 20*nu^2 triangles (stand-in for the absent Stanford meshes)."""
    p, n = C.c_void_p(), C.c_uint32()
    c = np.asarray(center, np.float32)
    _check(lib.rtb_mesh_geodesic(nu, radius, c.ctypes.data, displacement, seed, C.byref(p), C.byref(n)), "rtb_mesh_geodesic")
    pts = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n.value, 9)).copy()
    lib.rtb_free(p)
    return pts


class Trixel:
    """Mesh + tree (class Trixel, Trixel.h:39)."""

    def __init__(self, points9, colors=None, uniform_rgb=DEFAULT_RGB, require_device=True):
        pts = np.ascontiguousarray(points9, np.float32).reshape(-1, 9)
        self.num_trixels = pts.shape[0]
        self.num_voxels = 2 * self.num_trixels - 1
        rad = None if colors is None else np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
        uni = np.asarray(uniform_rgb, np.float32)
        h = C.c_void_p()
        rc = lib.rtb_mesh_create(pts.ctypes.data, pts.shape[0], None if rad is None else rad.ctypes.data, uni.ctypes.data, C.byref(h))
        self.h = h
        self.device_ok = rc == 0
        if rc != 0 and (require_device or not h.value):
            _check(rc, "rtb_mesh_create")

    def set_sorted_voxels(self):
        """Trixel.h:386.  The six sorted lists are produced together with the partition in create_kd()."""
        return 0

    def create_kd(self, where=0):
        """Trixel.h:135 (+ set_sorted_voxels): build the 2n-1 node tree.  where: 0 auto, 1 host, 2 GPU."""
        _check(lib.rtb_mesh_build_tree_on(self.h, where), "rtb_mesh_build_tree_on")
        return 0

    def save_tree(self, path):
        _check(lib.rtb_mesh_save_tree(self.h, os.fsencode(path)), "rtb_mesh_save_tree")

    def load_tree(self, path):
        """Instead of create_kd(): take the tree from a cache file written by save_tree() for this very mesh."""
        _check(lib.rtb_mesh_load_tree(self.h, os.fsencode(path)), "rtb_mesh_load_tree")

    def build_seconds(self):
        out = np.zeros(3, np.float64)
        _check(lib.rtb_mesh_build_seconds(self.h, out.ctypes.data), "rtb_mesh_build_seconds")
        return dict(sort=out[0], partition=out[1], total=out[2])

    def tree(self):
        N = self.num_voxels
        t = dict(left=np.empty(N, np.int32), right=np.empty(N, np.int32), tri=np.empty(N, np.int32),
                 cut_flag=np.empty(N, np.int32), bounds=np.empty((N, 6), np.float32), s1=np.empty(N, np.float32),
                 s2=np.empty(N, np.float32))
        _check(lib.rtb_mesh_get_tree(self.h, t["left"].ctypes.data, t["right"].ctypes.data, t["tri"].ctypes.data,
                                     t["cut_flag"].ctypes.data, t["bounds"].ctypes.data, t["s1"].ctypes.data,
                                     t["s2"].ctypes.data), "rtb_mesh_get_tree")
        return t

    def close(self):
        if self.h:
            lib.rtb_mesh_destroy(self.h)
            self.h = None


class Camera:
    """class Camera (Camera.h:15): film, primary rays, frame buffers."""

    def __init__(self, r_w, r_h, f_w, f_h, fclen, pos, look_at, up, require_device=True):
        self.W, self.H = int(r_w), int(r_h)
        p, la, u = (np.asarray(v, np.float32) for v in (pos, look_at, up))
        h = C.c_void_p()
        rc = lib.rtb_camera_create(self.W, self.H, f_w, f_h, fclen, p.ctypes.data, la.ctypes.data, u.ctypes.data, C.byref(h))
        self.h = h
        self.pos = p.copy()
        if rc != 0 and (require_device or not h.value):
            _check(rc, "rtb_camera_create")

    def basis(self):
        out = np.empty(18, np.float32)
        _check(lib.rtb_camera_get_basis(self.h, out.ctypes.data), "rtb_camera_get_basis")
        return out

    def add_object(self, obj):
        _check(lib.rtb_camera_add_object(self.h, obj.h), "rtb_camera_add_object")
        obj.camera = self

    def color_pixels(self, tag):
        _check(lib.rtb_camera_color_pixels(self.h, tag), "rtb_camera_color_pixels")

    def h_color(self):
        """Camera::h_mem.h_color.c -- view of the library-owned pinned frame (0x00RRGGBB, row 0 = bottom)."""
        p = lib.rtb_camera_host_color(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(self.W * self.H,))

    def h_ids(self):
        p = lib.rtb_camera_host_ids(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(self.W * self.H,))

    def tile_major_elements(self, tile_stride):
        """Per-frame element count of a RENDER_TILE_MAJOR buffer when tiles are dealt to `tile_stride` ranks."""
        return int(lib.rtb_tile_major_elements(self.h, tile_stride))

    def fill_frames_device_async(self, num_frames, bgra_ptr, ids_ptr, stream_ptr=None):
        """Background colour / -1 into `num_frames` device frames (the SET_COLOR_TAG fill, Camera.cu:12-18)."""
        if stream_ptr == 0:
            stream_ptr = 1  # cudaStreamLegacy
        _check(lib.rtb_fill_frames_device_async(self.h, num_frames, bgra_ptr or None, ids_ptr or None, stream_ptr or None),
               "rtb_fill_frames_device_async")

    def compose_tiles_device_async(self, num_frames, part_ptrs, out_ptr, stream_ptr=None):
        """Scatter gathered tile-major buffers (one device pointer per rank) into row-major frames."""
        if stream_ptr == 0:
            stream_ptr = 1  # cudaStreamLegacy
        arr = (C.c_void_p * len(part_ptrs))(*part_ptrs)
        _check(lib.rtb_compose_tiles_device_async(self.h, num_frames, len(part_ptrs), arr, out_ptr, stream_ptr or None),
               "rtb_compose_tiles_device_async")

    # ---- scene extension (include/rtb.h: lights, shadows, sample_rate, several objects) ----
    def set_lights(self, xyz):
        a = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        _check(lib.rtb_camera_set_lights(self.h, a.shape[0], a.ctypes.data), "rtb_camera_set_lights")

    def set_shadows(self, enable):
        _check(lib.rtb_camera_set_shadows(self.h, int(bool(enable))), "rtb_camera_set_shadows")

    def set_sample_rate(self, n):
        _check(lib.rtb_camera_set_sample_rate(self.h, int(n)), "rtb_camera_set_sample_rate")

    def object_id_base(self, obj):
        return int(lib.rtb_camera_object_id_base(self.h, obj.h))

    def render_scene(self, flags=RENDER_DEFAULT):
        """Camera::render(): every object added to this camera -> the camera's device frame (then color_pixels)."""
        _check(lib.rtb_camera_render_scene(self.h, flags), "rtb_camera_render_scene")

    def render_scene_frame(self, flags=RENDER_DEFAULT):
        """render_scene + color_pixels(PHONG); returns copies of (ids int32, colour uint32)."""
        self.render_scene(flags)
        self.color_pixels(PHONG_COLOR_TAG)
        return self.h_ids().copy(), self.h_color().copy()

    def render_scene_device_async(self, d_bgra_ptr, d_ids_ptr, stream_ptr=None, flags=RENDER_DEFAULT):
        if stream_ptr == 0:
            stream_ptr = 1  # cudaStreamLegacy
        _check(lib.rtb_camera_render_scene_device_async(self.h, flags, d_bgra_ptr or None, d_ids_ptr or None, stream_ptr or None),
               "rtb_camera_render_scene_device_async")

    def counters(self, reset=True):
        out = np.zeros(8, np.uint64)
        _check(lib.rtb_camera_counters_ex(self.h, out.ctypes.data, int(reset)), "rtb_camera_counters_ex")
        return dict(zip(["rays", "nodes", "boxes", "tris", "hits", "stack_depth_sum", "stack_depth_max", "ray_steps_max"], out.tolist()))

    def close(self):
        if self.h:
            lib.rtb_camera_destroy(self.h)
            self.h = None


class Object:
    """class Object (Object.h:6): a mesh instance with its quaternion transform."""

    def __init__(self, trixel):
        self.trixel_list = trixel
        self.camera = None
        h = C.c_void_p()
        _check(lib.rtb_object_create(trixel.h, C.byref(h)), "rtb_object_create")
        self.h = h

    def transform(self, xyzw, transform_select):
        """Input::set_quat(x,y,z,w) + Object::transform(input, select) (WinMain.cpp:186-209)."""
        q = np.asarray(xyzw, np.float32)
        _check(lib.rtb_object_transform(self.h, q.ctypes.data, transform_select), "rtb_object_transform")

    def transform_host(self, xyzw, transform_select):
        q = np.asarray(xyzw, np.float32)
        m = np.empty(12, np.float32)
        _check(lib.rtb_object_transform_host(self.h, q.ctypes.data, transform_select, m.ctypes.data), "rtb_object_transform_host")
        return m

    def matrix(self):
        m = np.empty(12, np.float32)
        _check(lib.rtb_object_get_matrix(self.h, m.ctypes.data), "rtb_object_get_matrix")
        return m

    def set_matrix(self, m12):
        m = np.ascontiguousarray(m12, np.float32)
        _check(lib.rtb_object_set_matrix(self.h, m.ctypes.data), "rtb_object_set_matrix")

    def render(self, camera, flags=RENDER_DEFAULT):
        """Object::render(Camera*) (Object.cpp:10)."""
        _check(lib.rtb_object_render(self.h, camera.h, flags), "rtb_object_render")

    def render_frame(self, camera, flags=RENDER_DEFAULT):
        """render + color_pixels(PHONG); returns copies of (ids int32, colour uint32)."""
        _check(lib.rtb_render_frame(self.h, camera.h, flags), "rtb_render_frame")
        return camera.h_ids().copy(), camera.h_color().copy()

    def render_sweep(self, camera, ops, flags=RENDER_DEFAULT, want_color=True, want_ids=True, out_color=None, out_ids=None):
        """ops: (frames, steps, 5) array of (select, x, y, z, w)."""
        ops = np.ascontiguousarray(ops, np.float32)
        if ops.ndim == 2:
            ops = ops[:, None, :]
        F, S = ops.shape[0], ops.shape[1]
        P = camera.W * camera.H
        if want_color and out_color is None:
            out_color = np.empty((F, P), np.uint32)
        if want_ids and out_ids is None:
            out_ids = np.empty((F, P), np.int32)
        _check(lib.rtb_render_sweep(self.h, camera.h, F, S, ops.ctypes.data, flags,
                                    out_color.ctypes.data if want_color else None, out_ids.ctypes.data if want_ids else None),
               "rtb_render_sweep")
        return out_ids, out_color

    def render_frames_device_async(self, camera, m12, d_bgra_ptr, d_ids_ptr, stream_ptr=None, tile_first=0, tile_stride=1,
                                   flags=RENDER_DEFAULT):
        """stream_ptr: None = the library's stream; 0 = the legacy default stream; else a cudaStream_t."""
        if stream_ptr == 0:
            stream_ptr = 1  # cudaStreamLegacy
        m = np.ascontiguousarray(m12, np.float32).reshape(-1, 12)
        _check(lib.rtb_render_frames_device_async(self.h, camera.h, m.shape[0], m.ctypes.data, tile_first, tile_stride, flags,
                                                  d_bgra_ptr or None, d_ids_ptr or None, stream_ptr or None),
               "rtb_render_frames_device_async")

    def render_frames_push_async(self, camera, m12, frame_bgra_ptr, frame_ids_ptr, stream_ptr=None, tile_first=0, tile_stride=1,
                                 flags=RENDER_DEFAULT):
        """Render this rank's tiles and push every finished work unit straight into the final frames, which may be
        another GPU's memory (peer_open).  See rtb_render_frames_push_async in include/rtb.h."""
        if stream_ptr == 0:
            stream_ptr = 1  # cudaStreamLegacy
        m = np.ascontiguousarray(m12, np.float32).reshape(-1, 12)
        _check(lib.rtb_render_frames_push_async(self.h, camera.h, m.shape[0], m.ctypes.data, tile_first, tile_stride, flags,
                                                frame_bgra_ptr or None, frame_ids_ptr or None, stream_ptr or None),
               "rtb_render_frames_push_async")

    def render_frames_push_striped_async(self, camera, m12, owner_bgra_ptrs, owner_ids_ptrs, stream_ptr=None, tile_first=0, tile_stride=1,
                                         flags=RENDER_DEFAULT):
        """Like render_frames_push_async with striped frame ownership: frame f goes to owner f % len(owners) as its frame
        f // len(owners); owner_*_ptrs are lists of device pointers (one per owner; either list may be None)."""
        if stream_ptr == 0:
            stream_ptr = 1  # cudaStreamLegacy
        m = np.ascontiguousarray(m12, np.float32).reshape(-1, 12)
        owners = len(owner_bgra_ptrs if owner_bgra_ptrs is not None else owner_ids_ptrs)
        ac = (C.c_void_p * owners)(*owner_bgra_ptrs) if owner_bgra_ptrs is not None else None
        ai = (C.c_void_p * owners)(*owner_ids_ptrs) if owner_ids_ptrs is not None else None
        _check(lib.rtb_render_frames_push_striped_async(self.h, camera.h, m.shape[0], m.ctypes.data, tile_first, tile_stride, flags, owners,
                                                        ac, ai, stream_ptr or None), "rtb_render_frames_push_striped_async")

    def close(self):
        if self.h:
            lib.rtb_object_destroy(self.h)
            self.h = None


def measure_l2_read_bandwidth(nbytes=32 << 20, iters=200):
    """GB/s of L1-bypassing 16-byte loads over an L2-resident buffer (the L2 leg of the roofline)."""
    out = C.c_double()
    _check(lib.rtb_measure_l2_read_bandwidth(nbytes, iters, C.byref(out)), "rtb_measure_l2_read_bandwidth")
    return out.value


def measure_host_fill_bandwidth(nbytes=512 << 20, threads=0):
    """GB/s at which `threads` host threads (0 = all) fill pinned memory with streaming stores."""
    out = C.c_double(0.0)
    _check(lib.rtb_measure_host_fill_bandwidth(nbytes, threads, C.byref(out)), "rtb_measure_host_fill_bandwidth")
    return out.value


def write_frame(path, bgra, W, H):
    """.ppm or .png by extension; bgra = W*H uint32 0x00RRGGBB, row 0 at the bottom (the reference's frame buffer)."""
    a = np.ascontiguousarray(bgra, np.uint32)
    _check(lib.rtb_write_frame(os.fsencode(path), a.ctypes.data, W, H), "rtb_write_frame")


class PeerBuffer:
    """A device allocation other ranks' GPUs can write into (rtb_peer_alloc + rtb_peer_export)."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        _check(lib.rtb_peer_alloc(nbytes, C.byref(p)), "rtb_peer_alloc")
        self.ptr, self.nbytes = p.value, nbytes

    def handle(self):
        h = (C.c_uint8 * 64)()
        _check(lib.rtb_peer_export(self.ptr, h), "rtb_peer_export")
        return bytes(h)

    def close(self):
        if self.ptr:
            _check(lib.rtb_peer_free(self.ptr), "rtb_peer_free")
            self.ptr = None


def peer_open(handle):
    """Map a peer rank's PeerBuffer (its 64-byte handle) into this process; returns the device pointer."""
    h = (C.c_uint8 * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    _check(lib.rtb_peer_open(h, C.byref(p)), "rtb_peer_open")
    return p.value


def memcpy_d2h(host_array, device_ptr):
    """Synchronous device-to-host copy of host_array.nbytes bytes (rtb_peer_read)."""
    _check(lib.rtb_peer_read(host_array.ctypes.data, device_ptr, host_array.nbytes), "rtb_peer_read")


def peer_close(ptr):
    _check(lib.rtb_peer_close(ptr), "rtb_peer_close")


def transform_sequence(cam_pos, ops):
    """Matrices of a transform sequence starting from a freshly added object (host only, no GPU): ops (count, 5) ->
    (count, 12).  rtb_transform_sequence_host."""
    ops = np.ascontiguousarray(ops, np.float32).reshape(-1, 5)
    pos = np.asarray(cam_pos, np.float32)
    out = np.empty((ops.shape[0], 12), np.float32)
    _check(lib.rtb_transform_sequence_host(pos.ctypes.data, ops.shape[0], ops.ctypes.data, out.ctypes.data), "rtb_transform_sequence_host")
    return out


def orbit_ops(num_frames, quat=R_KEY_QUAT, select=ROTATE_TRI_PY, first_frame_identity=True):
    """The animation sweep of BASELINE.json configs[2]: frame k = k applications of the R-key step."""
    ops = np.zeros((num_frames, 1, 5), np.float32)
    ops[:, 0, 0] = select
    ops[:, 0, 1:] = np.asarray(quat, np.float32)
    if first_frame_identity:
        ops[0, 0, 0] = 0
    return ops


def write_ppm(path, bgra, W, H):
    """Headless replacement of the GDI blit (WinMain.cpp:217): top-down binary PPM from the bottom-up frame."""
    img = np.asarray(bgra, np.uint32).reshape(H, W)[::-1]
    rgb = np.stack([(img >> 16) & 0xff, (img >> 8) & 0xff, img & 0xff], axis=-1).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (W, H))
        f.write(rgb.tobytes())
