"""Single-frame launches for ncu (development aid): `python tools/one_frame.py [count]` renders `count` orbit frames one
launch each (the per-frame path: record in the kernel parameters), synchronising after each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 40
W, H = 960, 540
pts = rtb.geodesic_mesh(209); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
for k in range(count):
    obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)
    obj.render(cam)
    cam.color_pixels(rtb.PHONG_COLOR_TAG)
print("rendered", count, "frames; hits in the last one:", int((cam.h_ids() >= 0).sum()))
