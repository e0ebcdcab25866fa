"""Intrinsic cost of the PUSH kernel variant on ONE GPU: this rank's share of the tiles rendered (a) into the tile-major
exchange buffer, (b) pushed into a local full-frame buffer.  (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
nu, W, H, F, zoom, world = (int(a) for a in (sys.argv[1:7] or ["233", "3840", "2160", "6", "0", "2"]))
P = W * H
pts = rtb.geodesic_mesh(nu); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
n = cam.basis()[0:3]
for _ in range(zoom): obj.transform((float(n[0]), float(n[1]), float(n[2]), 0.005), rtb.TRANSLATE_Z)
mats = np.stack([obj.matrix()] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(F - 1)])
st = torch.cuda.Stream()
PE = cam.tile_major_elements(world)
col = torch.empty(F * PE, dtype=torch.int32, device="cuda"); ids = torch.empty(F * PE, dtype=torch.int32, device="cuda")
full_c = rtb.PeerBuffer(4 * F * P); full_i = rtb.PeerBuffer(4 * F * P)
def timed(name, fn, reps=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    print("%-50s %8.3f ms" % (name, e0.elapsed_time(e1) / reps), flush=True)
timed("tile-major render of rank 0's tiles (1/%d)" % world, lambda: obj.render_frames_device_async(cam, mats, col.data_ptr(), ids.data_ptr(), st.cuda_stream, tile_first=0, tile_stride=world, flags=rtb.RENDER_TILE_MAJOR))
timed("push render of the same tiles into a LOCAL frame", lambda: obj.render_frames_push_async(cam, mats, full_c.ptr, full_i.ptr, st.cuda_stream, tile_first=0, tile_stride=world))
