"""Where a SINGLE frame spends its time (development aid): kernel-only event timings of one-frame launches under
different knobs, an all-background frame (the launch's fixed cost), and the host-side loop through the drop-in calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
nu, W, H = (int(a) for a in (sys.argv[1:4] or ["209", "960", "540"]))
pts = rtb.geodesic_mesh(nu); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
st = torch.cuda.Stream()
col = torch.empty(W * H, dtype=torch.int32, device="cuda"); ids = torch.empty(W * H, dtype=torch.int32, device="cuda")
m0 = obj.matrix()
mats = np.stack([m0] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(31)])
far = m0.copy(); far[3] += 50.0  # the object slides out of view: every pixel is background


def kernel_us(m, reps=30):
    ts = []
    for r in range(reps + 5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            obj.render_frames_device_async(cam, m[r % len(m)] if m.ndim == 2 else m, col.data_ptr(), ids.data_ptr(), st.cuda_stream)
            e1.record(st)
        torch.cuda.synchronize()
        if r >= 5: ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


print("lib:", rtb.LIB_PATH)
print("single-frame kernel, orbit frames  : median %.1f us  min %.1f us" % kernel_us(mats))
print("single-frame kernel, all background: median %.1f us  min %.1f us" % kernel_us(far))
for shift in (5, 6, 7):
    rtb.set_knob("unit_shift", shift)
    print("  unit_shift=%d                     : median %.1f us  min %.1f us" % ((shift,) + kernel_us(mats)))
rtb.set_knob("unit_shift", 0)
for ta in (4, 8, 12, 20):
    rtb.set_knob("t_active", ta)
    print("  t_active=%-2d                      : median %.1f us  min %.1f us" % ((ta,) + kernel_us(mats)))
rtb.set_knob("t_active", 12)
rtb.set_knob("no_rect", 1)
print("  no root-box rectangle            : median %.1f us  min %.1f us" % kernel_us(mats))
rtb.set_knob("no_rect", 0)
# 8 frames in one launch, per frame
e = []
for r in range(12):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    c8 = torch.empty(8 * W * H, dtype=torch.int32, device="cuda"); i8 = torch.empty(8 * W * H, dtype=torch.int32, device="cuda")
    with torch.cuda.stream(st):
        e0.record(st); obj.render_frames_device_async(cam, mats[:8], c8.data_ptr(), i8.data_ptr(), st.cuda_stream); e1.record(st)
    torch.cuda.synchronize()
    if r >= 2: e.append(e0.elapsed_time(e1) * 1e3)
print("8-frame launch                     : median %.1f us per launch" % float(np.median(e)))

N = 400
def loop(fn):
    for _ in range(20): fn()
    t = time.perf_counter()
    for _ in range(N): fn()
    return (time.perf_counter() - t) / N * 1e6
print("transform                          %8.1f us" % loop(lambda: obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)))
print("render (queued, no sync)           %8.1f us" % loop(lambda: obj.render(cam)))
torch.cuda.synchronize()
print("render + color_pixels(PHONG)       %8.1f us" % loop(lambda: (obj.render(cam), cam.color_pixels(rtb.PHONG_COLOR_TAG))))
def full():
    obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY); obj.render(cam); cam.color_pixels(rtb.PHONG_COLOR_TAG)
us = loop(full)
print("full loop iteration                %8.1f us  = %.0f FPS" % (us, 1e6 / us))
# SM clock while the loop runs (the GPU idles at 120 MHz; does a loop of 0.3 ms kernels + copies keep it at boost?)
try:
    import pynvml, threading
    pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    samples, stop = [], [False]
    def sample():
        while not stop[0]:
            samples.append(pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)); time.sleep(0.002)
    th = threading.Thread(target=sample); th.start()
    N = 2000; us = loop(full); stop[0] = True; th.join()
    print("  same loop, %d iterations: %.1f us; SM clock during it: median %d MHz, min %d, max %d (%d samples)" % (
        N, us, int(np.median(samples)), min(samples), max(samples), len(samples)))
    N = 400
except Exception as exc:
    print("  (no NVML: %s)" % exc)
# copy-only: the frame's 8 bytes per pixel to pinned host memory
hc = torch.empty(W * H, dtype=torch.int32).pin_memory(); hi = torch.empty(W * H, dtype=torch.int32).pin_memory()
def copy_only():
    with torch.cuda.stream(st):
        hc.copy_(col, non_blocking=True); hi.copy_(ids, non_blocking=True)
    st.synchronize()
print("D2H of one frame (colour + ids)    %8.1f us" % loop(copy_only))
