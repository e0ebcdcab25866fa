# round 2: frame order on/off, longest ray of a single frame, frames to host without the copy engine, build timings
set -x
export RTB_TUNE_FRAMES=600
timeout 900 python tools/tune.py FRAME_ORDER=0,1 2>&1 | tail -6
timeout 300 python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
W, H = 960, 540
pts = rtb.geodesic_mesh(209); mesh = rtb.Trixel(pts)
for k in range(3):
    mesh.create_kd(); print("build seconds (sort, partition, total):", mesh.build_seconds())
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
col = torch.empty(W * H, dtype=torch.int32, device="cuda"); ids = torch.empty(W * H, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for f in range(6):
    cam.counters(reset=True)
    obj.render_frames_device_async(cam, np.stack([obj.matrix()] * 2), col.data_ptr(), ids.data_ptr(), s, flags=rtb.RENDER_COUNTERS)
    torch.cuda.synchronize()
    c = cam.counters(reset=True)
    print("frame %d: traced rays %d (of %d pixels x2), steps per ray mean %.1f, longest ray %d steps, hits %d" % (
        f, c["rays"], W * H, (c["nodes"] + c["tris"]) / max(c["rays"], 1), c["ray_steps_max"], c["hits"]))
    for _ in range(5): obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)
PY
timeout 600 python tools/host_push_probe.py 2>&1 | tail -20
RTB_PROBE_ZOOM=140 RTB_PROBE_FRAMES=60 timeout 600 python tools/host_push_probe.py 2>&1 | tail -12
