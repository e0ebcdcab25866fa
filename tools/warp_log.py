"""Per-warp anatomy of ONE single-frame launch (development aid; needs a library built with -DRTB_WARP_LOG:
python tools/build_variants.py wlog=-DRTB_WARP_LOG=1; RTB_LIB=build/variants/librtb_wlog.so python tools/warp_log.py)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpp_cuda_raytracer_dev_b200 as rtb
rtb.set_device(0)
W, H = 960, 540
pts = rtb.geodesic_mesh(209); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    rtb.set_knob("unit_shift", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    for _ in range(12):
        obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY); obj.render(cam); cam.color_pixels(rtb.PHONG_COLOR_TAG)
    warps = 148 * 8 * 4
    log = np.zeros((warps, 8), np.uint64)
    rtb.lib.rtb_camera_warp_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    assert rtb.lib.rtb_camera_warp_log(cam.h, log.ctypes.data, warps) == 0
    t0, t1, tw, tx = (log[:, k].astype(np.int64) for k in range(4))
    live = t1 > 0
    base = t0[live].min()
    units, bg = (log[:, 4] >> np.uint64(32)).astype(np.int64), (log[:, 4] & np.uint64(0xffffffff)).astype(np.int64)
    iters, rays = (log[:, 5] >> np.uint64(32)).astype(np.int64), (log[:, 5] & np.uint64(0xffffffff)).astype(np.int64)
    sm = log[:, 6].astype(np.int64)
    steals, rounds = (log[:, 7] & np.uint64(0xffffffff)).astype(np.int64), (log[:, 7] >> np.uint64(32)).astype(np.int64)
    dur = (t1 - t0) / 1e3
    print("warps that ran: %d; kernel span %.1f us (first start -> last end); starts spread over %.1f us" % (live.sum(), (t1[live].max() - base) / 1e3, (t0[live].max() - base) / 1e3))
    print("queue exhausted (first warp to see it) at %.1f us, last at %.1f us" % ((tx[tx > 0].min() - base) / 1e3, (tx[tx > 0].max() - base) / 1e3))
    busy = rays > 0
    print("warps with rays: %d; rays per such warp: mean %.1f max %d; iterations: mean %.1f max %d; non-bg units per warp: mean %.2f max %d" % (
        busy.sum(), rays[busy].mean(), rays.max(), iters[busy].mean(), iters.max(), (units - bg)[busy].mean(), (units - bg).max()))
    print("duration of warps with rays: p50 %.1f  p90 %.1f  p99 %.1f  max %.1f us;  ns per iteration of the slowest 1%%: %.0f" % (
        np.percentile(dur[busy], 50), np.percentile(dur[busy], 90), np.percentile(dur[busy], 99), dur[busy].max(),
        1e3 * (dur[busy] / np.maximum(iters[busy], 1))[dur[busy] >= np.percentile(dur[busy], 99)].mean()))
    print("subtrees handed over between lanes: %d in %d warps (max %d in one warp); drain rounds: max %d" % (steals.sum(), (steals > 0).sum(), steals.max(), rounds.max()))
    ends = np.sort((t1[busy] - base) / 1e3)
    print("end times of warps with rays: p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f us" % tuple(np.percentile(ends, q) for q in (10, 50, 90, 99, 100)))
    order = np.argsort(-dur)[:8]
    for w in order:
        print("  warp %5d sm %3d: start %.1f first-work %.1f exhausted %.1f end %.1f us  units %d (bg %d) rays %d iters %d steals %d rounds %d" % (
            w, sm[w], (t0[w] - base) / 1e3, (tw[w] - base) / 1e3 if tw[w] else -1, (tx[w] - base) / 1e3 if tx[w] else -1, (t1[w] - base) / 1e3, units[w], bg[w], rays[w], iters[w], steals[w], rounds[w]))
    per_sm = np.bincount(sm[busy], weights=rays[busy].astype(np.float64), minlength=148)
    print("rays per SM: mean %.0f min %.0f max %.0f; slowest warps' SMs hold %s rays" % (per_sm.mean(), per_sm.min(), per_sm.max(), [int(per_sm[sm[w]]) for w in order[:4]]))
