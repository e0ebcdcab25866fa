"""Per-call cost of the reference-style frame loop (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
nu, W, H = (int(a) for a in (sys.argv[1:4] or ["209", "960", "540"]))
pts = rtb.geodesic_mesh(nu); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
N = 300
def loop(fn):
    for _ in range(20): fn()
    t = time.perf_counter()
    for _ in range(N): fn()
    return (time.perf_counter() - t) / N * 1e6
print("transform                  %8.1f us" % loop(lambda: obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)))
print("render (sync)              %8.1f us" % loop(lambda: obj.render(cam)))
print("color_pixels(PHONG) (sync) %8.1f us" % loop(lambda: cam.color_pixels(rtb.PHONG_COLOR_TAG)))
print("render_frame (fused)       %8.1f us" % loop(lambda: obj.render_frame(cam)))
def full():
    obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY); obj.render(cam); cam.color_pixels(rtb.PHONG_COLOR_TAG)
us = loop(full)
print("full loop iteration        %8.1f us  = %.0f FPS" % (us, 1e6 / us))
