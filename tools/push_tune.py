"""Knobs of the PUSH kernel variant in tile mode on ONE GPU (development aid): rank 0's share (1/world) of the tiles of F frames at
4K pushed into local, pre-filled frames, against 1/world of the time the same F frames take as whole frames."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
nu, W, H, F, world = 233, 3840, 2160, int(os.environ.get("RTB_TUNE_FRAMES", "96")), int(os.environ.get("RTB_TUNE_WORLD", "4"))
P = W * H
pts = rtb.geodesic_mesh(nu); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
mats = np.stack([obj.matrix()] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(F - 1)])
st = torch.cuda.Stream()
full_c = torch.empty(F * P, dtype=torch.int32, device="cuda"); full_i = torch.empty(F * P, dtype=torch.int32, device="cuda")
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
def measure(fn, reps=4):
    ts = []
    for r in range(reps + 1):
        with torch.cuda.stream(st):
            flush.fill_(r)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st)
        torch.cuda.synchronize()
        if r: ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts))
whole = lambda: obj.render_frames_device_async(cam, mats, full_c.data_ptr(), full_i.data_ptr(), st.cuda_stream)
push = lambda: obj.render_frames_push_async(cam, mats, full_c.data_ptr(), full_i.data_ptr(), st.cuda_stream, tile_first=0, tile_stride=world, flags=rtb.RENDER_PUSH_PREFILLED)
t_whole = measure(whole)
print("%d whole frames: %.3f ms -> 1/%d = %.3f ms" % (F, t_whole, world, t_whole / world))
grid = {}
for arg in sys.argv[1:]:
    k, v = arg.split("="); grid[k.lower()] = [int(x) for x in v.split(",")]
keys = list(grid)
for combo in itertools.product(*[grid[k] for k in keys]):
    for k, v in zip(keys, combo): rtb.set_knob(k, v)
    t = measure(push)
    print("%-44s %8.3f ms  efficiency ceiling %.3f" % (" ".join("%s=%d" % kv for kv in zip(keys, combo)), t, t_whole / world / t), flush=True)
