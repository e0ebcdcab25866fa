"""Sweep the render kernel's scheduling knobs (RTB_T_ACTIVE, RTB_T_LEAF, RTB_UNIT_SHIFT) in one process (development aid).
The library reads them at every launch, so the scene is built once."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
nu, W, H, F = 209, 960, 540, int(os.environ.get("RTB_TUNE_FRAMES", "60"))
pts = rtb.geodesic_mesh(nu); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
st = torch.cuda.Stream()
col = torch.empty(F * W * H, dtype=torch.int32, device="cuda"); ids = torch.empty(F * W * H, dtype=torch.int32, device="cuda")
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
def measure(mats, reps=4):
    ts = []
    for r in range(reps + 1):
        with torch.cuda.stream(st):
            flush.fill_(r)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            obj.render_frames_device_async(cam, mats, col.data_ptr(), ids.data_ptr(), st.cuda_stream)
            e1.record(st)
        torch.cuda.synchronize()
        if r: ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts)), float(np.min(ts))
n = cam.basis()[0:3]
views = {}
views["default"] = np.stack([obj.matrix()] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(F - 1)])
for _ in range(140):
    obj.transform((float(n[0]), float(n[1]), float(n[2]), 0.005), rtb.TRANSLATE_Z)
views["closeup"] = np.stack([obj.matrix()] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(F - 1)])
# arguments: KEY=v1,v2,... (RTB_ prefix implied), e.g.  T_ACTIVE=12,20 PREFETCH=0,1,2,3
grid = {}
for arg in sys.argv[1:]:
    k, v = arg.split("=")
    grid["RTB_" + k] = v.split(",")
if not grid:
    grid = {"RTB_T_ACTIVE": ["12", "16", "20", "24"], "RTB_T_LEAF": ["4", "8", "12"]}
keys = list(grid)
# what the traversal stack looks like on these views (untimed counter pass)
for name in ("default", "closeup"):
    cam.counters(reset=True)
    obj.render_frames_device_async(cam, views[name][:8], col.data_ptr(), ids.data_ptr(), st.cuda_stream, flags=rtb.RENDER_COUNTERS)
    torch.cuda.synchronize()
    c = cam.counters(reset=True)
    print("%s: per ray nodes %.2f tris %.2f hits %.3f  deepest stack: mean %.2f max %d" % (
        name, c["nodes"] / c["rays"], c["tris"] / c["rays"], c["hits"] / c["rays"], c["stack_depth_sum"] / c["rays"], c["stack_depth_max"]))
print("lib:", rtb.LIB_PATH)
print("%-36s %18s %18s" % ("setting", "default mean/min ms", "closeup mean/min ms"))
for combo in itertools.product(*[grid[k] for k in keys]):
    for k, v in zip(keys, combo): rtb.set_knob(k[4:].lower(), int(v))
    if any(k.startswith("RTB_L2_") for k in keys):
        cam.add_object(obj)  # the L2 window is chosen when an object is added
    a = measure(views["default"]); b = measure(views["closeup"])
    print("%-36s %8.3f /%8.3f %8.3f /%8.3f" % (" ".join("%s=%s" % (k[4:], v) for k, v in zip(keys, combo)), a[0], a[1], b[0], b[1]), flush=True)
