"""Frames to HOST memory without the copy engine (development aid): the push variant of the render kernel with a pinned,
mapped host buffer as the frames' owner -- finished work units leave the SM as 128-byte row stores over PCIe -- against
render + cudaMemcpyAsync; with and without the host pre-filling the background (multi-threaded fill of pinned memory)."""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
nu, W, H, F = 209, 960, 540, int(os.environ.get("RTB_PROBE_FRAMES", "120"))
zoom = int(os.environ.get("RTB_PROBE_ZOOM", "0"))
pts = rtb.geodesic_mesh(nu); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
n = cam.basis()[0:3]
for _ in range(zoom):
    obj.transform((float(n[0]), float(n[1]), float(n[2]), 0.005), rtb.TRANSLATE_Z)
mats = np.stack([obj.matrix()] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(F - 1)])
P = W * H
st = torch.cuda.Stream()
h_col = torch.empty(F * P, dtype=torch.int32).pin_memory(); h_ids = torch.empty(F * P, dtype=torch.int32).pin_memory()
d_col = torch.empty(F * P, dtype=torch.int32, device="cuda"); d_ids = torch.empty(F * P, dtype=torch.int32, device="cuda")
threads = len(os.sched_getaffinity(0))
pool = ThreadPoolExecutor(threads)
bg = 0x00f08200  # Camera.cpp:72 background, 0x00RRGGBB

def host_fill(nthreads):
    cols = h_col.numpy(); ids = h_ids.numpy()
    parts = np.linspace(0, F * P, nthreads + 1).astype(np.int64)
    def work(k):
        cols[parts[k]:parts[k + 1]].fill(bg); ids[parts[k]:parts[k + 1]].fill(-1)
    t = time.perf_counter()
    list(pool.map(work, range(nthreads)))
    return time.perf_counter() - t

print("host cores usable: %d; frames %d (%.0f MB of colour + ids)" % (threads, F, F * P * 8 / 1e6))
for nt in (1, 4, 8, threads):
    ts = [host_fill(nt) for _ in range(3)]
    print("host fill, %2d threads: %.1f ms = %.1f GB/s" % (nt, min(ts) * 1e3, F * P * 8 / min(ts) / 1e9))

def timed(fn, reps=4):
    ts = []
    for r in range(reps + 1):
        torch.cuda.synchronize()
        t = time.perf_counter(); fn(); torch.cuda.synchronize()
        if r: ts.append(time.perf_counter() - t)
    return min(ts) * 1e3

def render_copy():
    with torch.cuda.stream(st):
        obj.render_frames_device_async(cam, mats, d_col.data_ptr(), d_ids.data_ptr(), st.cuda_stream)
        h_col.copy_(d_col, non_blocking=True); h_ids.copy_(d_ids, non_blocking=True)
def render_only():
    obj.render_frames_device_async(cam, mats, d_col.data_ptr(), d_ids.data_ptr(), st.cuda_stream)
def push_all():
    obj.render_frames_push_striped_async(cam, mats, [h_col.data_ptr()], [h_ids.data_ptr()], st.cuda_stream)
def push_prefilled():
    obj.render_frames_push_striped_async(cam, mats, [h_col.data_ptr()], [h_ids.data_ptr()], st.cuda_stream, flags=rtb.RENDER_PUSH_PREFILLED)
def fill_then_push():
    host_fill(threads); push_prefilled()

ref_c = ref_i = None
for name, fn in (("render only (device frames)", render_only), ("render + D2H copy", render_copy), ("push everything to host", push_all),
                 ("push, background pre-filled (push only)", push_prefilled), ("host fill + push of the rest", fill_then_push)):
    if name.startswith("push, background"):
        host_fill(threads)
    ms = timed(fn)
    print("%-42s %8.2f ms per %d frames = %7.0f FPS, %6.2f GB/s of frames" % (name, ms, F, F / ms * 1e3, F * P * 8 / ms / 1e6))
    if name == "render + D2H copy":
        ref_c = h_col.clone(); ref_i = h_ids.clone()
    elif ref_c is not None and name != "render only (device frames)":
        print("    frames equal to render + copy:", bool(torch.equal(ref_c, h_col)) and bool(torch.equal(ref_i, h_ids)))
