"""Where does a large single-frame sweep spend its time?  (development aid)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
nu, W, H = (int(a) for a in (sys.argv[1:4] or ["233", "7680", "4320"]))
pts = rtb.geodesic_mesh(nu)
mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
P = W * H
for F in (1, 2, 4):
    h_col = torch.empty((F, P), dtype=torch.int32).pin_memory(); h_ids = torch.empty((F, P), dtype=torch.int32).pin_memory()
    for rep in range(4):
        ops = rtb.orbit_ops(F, first_frame_identity=False)
        t = time.perf_counter(); obj.render_sweep(cam, ops, out_color=h_col.numpy().view(np.uint32), out_ids=h_ids.numpy()); dt = time.perf_counter() - t
        print("F=%d rep %d: %.2f ms per frame, %.1f GB/s to host" % (F, rep, dt / F * 1e3, F * P * 8 / dt / 1e9), flush=True)
d = torch.empty(P * 2, dtype=torch.int32, device="cuda"); h = torch.empty(P * 2, dtype=torch.int32).pin_memory()
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("plain pinned D2H of %d MB: %.2f ms, %.1f GB/s" % (P * 8 >> 20, dt * 1e3, P * 8 / dt / 1e9))
