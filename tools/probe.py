"""Quick GPU probe (development aid): build the dragon stand-in, time kernel-only frames, print counters."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cpp_cuda_raytracer_dev_b200 as rtb

nu = int(sys.argv[1]) if len(sys.argv) > 1 else 209
W = int(sys.argv[2]) if len(sys.argv) > 2 else 960
H = int(sys.argv[3]) if len(sys.argv) > 3 else 540
F = int(sys.argv[4]) if len(sys.argv) > 4 else 60
zoom = int(sys.argv[5]) if len(sys.argv) > 5 else 0
rtb.set_device(0)
print(rtb.device_props())
t = time.time(); pts = rtb.geodesic_mesh(nu); print("mesh", pts.shape, "%.2fs" % (time.time() - t))
mesh = rtb.Trixel(pts); t = time.time(); mesh.create_kd(); print("build", mesh.build_seconds())
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); t = time.time(); cam.add_object(obj); print("add_object %.3fs" % (time.time() - t))
n = cam.basis()[0:3]
for _ in range(zoom):
    obj.transform((float(n[0]), float(n[1]), float(n[2]), 0.005), rtb.TRANSLATE_Z)
mats = [obj.matrix()]
for k in range(1, F):
    mats.append(obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY))
mats = np.stack(mats)
col = torch.empty(F * W * H, dtype=torch.int32, device="cuda"); ids = torch.empty(F * W * H, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for flags, name in ((rtb.RENDER_COUNTERS, "cull+count"), (rtb.RENDER_NO_CULL | rtb.RENDER_COUNTERS, "nocull+count"), (0, "cull"), (rtb.RENDER_NO_CULL, "nocull")):
    cam.counters(reset=True)
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); obj.render_frames_device_async(cam, mats, col.data_ptr(), ids.data_ptr(), s, flags=flags); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    c = cam.counters()
    rays = F * W * H
    print("%-13s %8.3f ms/%d frames  %8.1f FPS %8.1f Mrays/s  hits %.1f%%" % (name, ms, F, F / ms * 1e3, rays / ms / 1e3, 100.0 * (ids >= 0).float().mean().item()),
          {k: round(v / max(c["rays"], 1), 2) for k, v in c.items()} if c["rays"] else "")
# single frame latency through the host API
t = time.time()
for _ in range(20): obj.render_frame(cam)
print("render_frame (sync, D2H) %.3f ms" % ((time.time() - t) / 20 * 1e3))
ops = rtb.orbit_ops(F)
out_c = np.empty((F, W * H), np.uint32); out_i = np.empty((F, W * H), np.int32)
for rep in range(2):
    t = time.time(); obj.render_sweep(cam, ops, out_color=out_c, out_ids=out_i); dt = time.time() - t
print("render_sweep pageable out: %.1f FPS" % (F / dt))
pc = torch.empty((F, W * H), dtype=torch.int32).pin_memory(); pi = torch.empty((F, W * H), dtype=torch.int32).pin_memory()
for rep in range(2):
    t = time.time(); obj.render_sweep(cam, ops, out_color=pc.numpy().view(np.uint32), out_ids=pi.numpy()); dt = time.time() - t
print("render_sweep pinned out: %.1f FPS" % (F / dt))
