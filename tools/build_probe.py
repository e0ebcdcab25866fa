"""Time the host and the GPU tree builders (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
for nu in [int(a) for a in (sys.argv[1:] or ["59", "209", "707"])]:
    pts = rtb.geodesic_mesh(nu)
    for where, name in ((2, "gpu"), (2, "gpu"), (2, "gpu")) + (((1, "host"),) if nu <= 300 else ()):
        m = rtb.Trixel(pts)
        t = time.time(); m.create_kd(where=where); dt = time.time() - t
        cam = rtb.Camera(960, 540, **rtb.default_camera_args(960, 540)); obj = rtb.Object(m)
        t = time.time(); cam.add_object(obj); da = time.time() - t
        print("n=%9d %-5s create_kd %.4f s  %s  add_object %.4f s" % (len(pts), name, dt, {k: round(float(v), 4) for k, v in m.build_seconds().items()}, da), flush=True)
        obj.close(); cam.close(); m.close()
