"""Kernel-only timing of the dragon stand-in at several zoom levels (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
nu = int(os.environ.get("NU", 209)); W = int(os.environ.get("W", 960)); H = int(os.environ.get("H", 540)); F = int(os.environ.get("F", 60))
zooms = [int(z) for z in (sys.argv[1:] or ["0", "120", "140"])]
rtb.set_device(0)
pts = rtb.geodesic_mesh(nu); mesh = rtb.Trixel(pts); mesh.create_kd()
col = torch.empty(F * W * H, dtype=torch.int32, device="cuda"); ids = torch.empty(F * W * H, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
for z in zooms:
    cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
    n = cam.basis()[0:3]
    for _ in range(z): obj.transform((float(n[0]), float(n[1]), float(n[2]), 0.005), rtb.TRANSLATE_Z)
    mats = np.stack([obj.matrix()] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(F - 1)])
    best = 1e9
    with torch.cuda.stream(st):
        for rep in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); obj.render_frames_device_async(cam, mats, col.data_ptr(), ids.data_ptr(), st.cuda_stream); e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    cov = (ids >= 0).float().mean().item()
    print("zoom %3d cov %5.1f%%  %8.3f ms/%d frames %9.1f FPS %8.1f Mrays/s  chunk=%s" % (z, 100 * cov, best, F, F / best * 1e3, F * W * H / best / 1e3, os.environ.get("RTB_CHUNK", "auto")), flush=True)
    obj.close(); cam.close()
