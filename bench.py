#!/usr/bin/env python3
"""bench.py -- headline benchmark of the ray-cast path (BASELINE.json: Mrays/s + FPS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workload at N=1 (BASELINE.json configs[2], the configuration the README's ~100 FPS is quoted on):
Stanford Dragon (871k triangles) at 960x540, quaternion orbit with the R-key step per frame, default
camera.  The dragon PLY is NOT in the reference checkout (.MISSING_LARGE_BLOBS) and there is no
network, so unless $RTB_MESH_DIR/dragon_vrip_mod.ply exists the mesh is the labelled stand-in: a
displaced geodesic icosphere with 873 620 triangles of the dragon's size.  `config.mesh` says which.

A "step" is one pass of the hot path over one batch: FRAMES_PER_STEP consecutive frames of the
orbit (default: the whole 600-frame orbit of configs[2] in one persistent launch).  `value` = Mrays/s with everything resident
in HBM (kernel-only, CUDA events on the launching stream); `e2e` = the same metric through the C-ABI
call rtb_render_sweep with HOST buffers: transform ops in, every frame's colour + hit-id buffer out
to pinned host memory inside the timed region.  N > 1 (torchrun), scene replicated on every rank:
  --shard frames (default): the sweep is dealt out in blocks of FRAMES_PER_STEP frames per rank and
      step (weak scaling); frames are independent units, so there is NO data-path collective --
      every rank delivers its own frames (BASELINE.json: "for the animation sweep, by frames");
  --shard tiles: every frame is split into interleaved 32x32 tiles (tile t -> rank t % N), a step is
      N * FRAMES_PER_STEP frames (weak scaling), and the finished tiles are gathered to rank 0 with
      NCCL on a side stream and reassembled there -- the path's one real exchange step.

`--impl reference` times the reference's own kernels compiled for the host (oracle/_ref, falling
back to the C port) on a bounded sample of the same workload with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (mesh file, ply mode, stand-in nu, W, H, frames per step, zoom steps)
    # frames per step = the sweep BASELINE.json quotes the configuration on, where one launch can hold it: the whole
    # 600-frame orbit of configs[2] is ONE step (one persistent launch; 2.5 GB of frames), configs[3]'s 360 frames at 4K
    # are ten steps of 36.  A launch has a fixed cost (ramp-up plus the latency of its last rays, ~0.4 ms measured), so the
    # batch a step renders is part of the workload definition and is stated in config.frames_per_step.
    "dragon_orbit_960x540": ("dragon_vrip_mod.ply", 0, 209, 960, 540, 600, 0),
    "dragon_closeup_960x540": ("dragon_vrip_mod.ply", 0, 209, 960, 540, 600, 140),
    "happy_orbit_3840x2160": ("happy_vrip_mod.ply", 0, 233, 3840, 2160, 36, 0),
    "bunny_960x540": ("rabbit_70k.ply", 1, 59, 960, 540, 600, 0),
    "synthetic10m_7680x4320": (None, 0, 707, 7680, 4320, 4, 0),
}
# tiles mode (N > 1) renders N x this many frames per step (weak scaling): kept small, rank 0 holds all of them twice
TILES_FRAMES = {"dragon_orbit_960x540": 60, "dragon_closeup_960x540": 60, "happy_orbit_3840x2160": 6, "bunny_960x540": 60, "synthetic10m_7680x4320": 1}
README_FPS = 100.0  # /root/reference/README.md:19 (Stanford Dragon, 960x540, unnamed GPU)


def find_mesh(name):
    if not name:
        return None
    for d in (os.environ.get("RTB_MESH_DIR", ""), os.path.join(ROOT, "oracle", "_ref", "data")):
        if d and os.path.exists(os.path.join(d, name)):
            return os.path.join(d, name)
    return None


def load_points(rtb, workload):
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[workload]
    path = find_mesh(fname)
    if path:
        return rtb.read_ply(path, mode), os.path.basename(path)
    return rtb.geodesic_mesh(nu), "stand-in: displaced geodesic icosphere nu=%d (%d triangles); %s absent from the reference checkout" % (
        nu, 20 * nu * nu, fname or "no file")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # first calls are slow and take driver locks: make them before any timed region
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def orbit_ops_block(rtb, first_frame, frames):
    """ops for frames [first_frame, first_frame+frames) of the orbit when the object is at frame first_frame-1."""
    ops = rtb.orbit_ops(frames, first_frame_identity=(first_frame == 0))
    return ops


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(workload, frames, repeats, prefer_ref=True):
    """Time the reference's own CPU-executed kernels (oracle/_ref) or the C port on `frames` frames of the
    workload.  Returns (Mrays/s, seconds per repeat list, kind, cores, sample text)."""
    import cpp_cuda_raytracer_dev_b200 as rtb  # mesh input only (loader / generator)
    from oracle import orc, refemu
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[workload]
    pts, mesh_label = load_points(rtb, workload)
    cam = orc.default_camera(W, H)
    # (beyond ~2 M triangles the reference's own host build takes minutes: the C port, which builds with all cores, stands in)
    kind = "reference" if (prefer_ref and refemu.available() and len(pts) <= 2_000_000) else "port"
    t0 = time.time()
    if kind == "reference":
        scene = refemu.RefScene(W, H, cam, points9=pts)
        cores = refemu.lib().ref_threads()
    else:
        scene = orc.Scene(pts, W, H, cam)
        cores = orc.lib().orc_threads()
    build_s = time.time() - t0
    n = np.array([0.0, 0.0, 1.0], np.float32)
    for _ in range(zoom):
        scene.transform(32, float(n[0]), float(n[1]), float(n[2]), 0.005)
    times = []
    for rep in range(repeats):
        t = time.perf_counter()
        for f in range(frames):
            if kind == "reference":
                scene.render_nocopy()
            else:
                scene.render()
            scene.transform(10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
        times.append(time.perf_counter() - t)
    sample = "%d consecutive orbit frames of %s at %dx%d per step (traversal + shading, tree build %.1f s excluded)" % (frames, workload, W, H, build_s)
    return frames * W * H, times, kind, cores, sample, mesh_label


def gpu_reference_run(workload, frames, impl="cuda_fmad"):
    """The reference's own CUDA kernels (Trixel.cu / Camera.cu compiled by nvcc for sm_100a with the project's default
    code generation, oracle/_ref/libref_cuda_fmad.so) on this GPU, on `frames` consecutive frames of the workload, timed
    the way the reference times itself (wall clock per loop iteration, WinMain.cpp:219-228) but WITHOUT its window blit,
    console output and second color_pixels call -- a lower bound of its per-frame cost: "the kernel to beat" on this box."""
    import cpp_cuda_raytracer_dev_b200 as rtb  # mesh input only
    from oracle import orc, refemu
    if not refemu.available(impl):
        return None
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[workload]
    pts, _ = load_points(rtb, workload)
    if len(pts) > 2_000_000:  # the reference's host-side tree build alone would take minutes (80-byte records x 6 lists)
        return {"skipped": "the reference's own single-threaded tree build is not practical for %d triangles" % len(pts)}
    t0 = time.time()
    scene = refemu.RefScene(W, H, orc.default_camera(W, H), points9=pts, impl=impl)
    build_s = time.time() - t0
    n = np.array([0.0, 0.0, 1.0], np.float32)
    for _ in range(zoom):
        scene.transform(32, float(n[0]), float(n[1]), float(n[2]), 0.005)
    per_frame = []
    for f in range(frames + 3):
        t = time.perf_counter()
        scene.render_nocopy()  # Object::render + Camera::color_pixels (kernels, device syncs, colour buffer D2H)
        scene.transform(10, 0.0, 0.09950371902099893, 0.0, 0.9950371902099893)
        per_frame.append(time.perf_counter() - t)
    total = sum(per_frame[3:])
    kind = ("reference CUDA kernels, nvcc sm_100a, default fmad" if impl == "cuda_fmad" else
            "the reference's own host classes with its .cu files replaced by integration/rtb_seam.cpp over librtb.so")
    return {"value": frames * W * H / total / 1e6, "unit": "Mrays/s", "fps": frames / total, "kind": kind,
            "sample": "%d consecutive orbit frames after 3 warm-up frames, wall clock around Object::render + Camera::color_pixels "
                      "(its host tree build took %.1f s)" % (frames, build_s)}


def run_reference_impl(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[args.workload]
    frames = max(1, args.ref_frames)
    rays, times, kind, cores, sample, mesh_label = cpu_reference_run(args.workload, frames, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = rays * len(timed) / total / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s (primary rays, traversal + Phong)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(timed) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "mesh": mesh_label, "resolution": [W, H], "frames_per_step": frames, "fps": frames * len(timed) / total},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import cpp_cuda_raytracer_dev_b200 as rtb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available() or rtb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the ray-cast path has no CPU fallback")
    torch.cuda.set_device(local)
    rtb.set_device(local)
    numa = None
    if world > 1:
        # one process per GPU: run on (and first-touch the pinned frame buffers from) the CPUs next to this GPU, so that
        # eight ranks' 50 GB/s DMA streams do not all cross the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            mask_words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local), (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(mask_words) for b in range(64) if (word >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa = "%d CPUs local to GPU %d" % (len(cpus), local)
        except Exception as exc:  # affinity is an optimisation, never a requirement
            numa = "not set (%s)" % type(exc).__name__
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    fname, mode, nu, W, H, F, zoom = WORKLOADS[args.workload]
    if world > 1 and args.shard == "tiles":
        F = TILES_FRAMES[args.workload]
    if args.frames_per_step:
        F = args.frames_per_step
    P = W * H
    pts, mesh_label = load_points(rtb, args.workload)
    mesh = rtb.Trixel(pts)
    mesh.create_kd()
    build_s = mesh.build_seconds()
    cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H))
    obj = rtb.Object(mesh)
    cam.add_object(obj)
    nvec = cam.basis()[0:3]
    for _ in range(zoom):
        obj.transform((float(nvec[0]), float(nvec[1]), float(nvec[2]), 0.005), rtb.TRANSLATE_Z)

    K, Wm = args.steps, args.warmup
    total_steps = K + Wm
    # Orbit: global frame g of step s on rank r is frame (s*world + r)*F + j.  Matrices come from the
    # host recurrence (the reference's Object::transform), so all ranks derive them identically.
    mats = np.empty((total_steps * world * F, 12), np.float32)
    mats[0] = obj.matrix()
    for g in range(1, len(mats)):
        mats[g] = obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)

    tiles_mode = world > 1 and args.shard == "tiles"
    push_mode = tiles_mode and args.exchange == "push"
    if tiles_mode and not push_mode:  # the NCCL gather runs beside the persistent render kernel: leave it a few SMs
        rtb.set_knob("reserve_sms", int(os.environ.get("RTB_RESERVE_SMS", "8")))
    FS = F * world if tiles_mode else F  # frames a rank touches per step

    def my_mats(step):
        if tiles_mode:  # every rank renders its tiles of all N*F frames of the step
            return mats[step * world * F:(step + 1) * world * F]
        b = (step * world + rank) * F
        return mats[b:b + F]

    stream = torch.cuda.Stream()
    side = torch.cuda.Stream()
    # per-frame elements of this rank's output: the whole frame, or (tiles mode) only its own tiles in
    # the compact tile-major exchange format
    PE = cam.tile_major_elements(world) if tiles_mode else P
    if not (tiles_mode and args.exchange == "push"):
        d_col = [torch.empty(FS * PE, dtype=torch.int32, device="cuda") for _ in range(2)]
        d_ids = [torch.empty(FS * PE, dtype=torch.int32, device="cuda") for _ in range(2)]
    gather_col = gather_ids = final_col = final_ids = None
    push_ptr = None
    if push_mode:
        # rank 0 owns the final frames (two slots); every rank maps them and its render kernel stores finished work
        # units straight into them over NVLink -- no gather, no receive buffers, no reassembly pass
        d_col = d_ids = None
        bufs = [rtb.PeerBuffer(4 * FS * P) for _ in range(4)] if rank == 0 else None
        handles = [b.handle() for b in bufs] if rank == 0 else [None] * 4
        dist.broadcast_object_list(handles, src=0)
        ptrs = [b.ptr for b in bufs] if rank == 0 else [rtb.peer_open(h) for h in handles]
        push_ptr = [(ptrs[0], ptrs[1]), (ptrs[2], ptrs[3])]  # per slot: (colours, ids)
        push_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        if rank == 0:
            cam.fill_frames_device_async(FS, push_ptr[0][0], push_ptr[0][1], stream.cuda_stream)
        torch.cuda.synchronize()
        dist.barrier()
    elif tiles_mode and rank == 0:
        gather_col = [[torch.empty(FS * PE, dtype=torch.int32, device="cuda") for _ in range(world)] for _ in range(2)]
        gather_ids = [[torch.empty(FS * PE, dtype=torch.int32, device="cuda") for _ in range(world)] for _ in range(2)]
        final_col = torch.empty(FS * P, dtype=torch.int32, device="cuda")
        final_ids = torch.empty(FS * P, dtype=torch.int32, device="cuda")

    def compose(slot):
        if tiles_mode and rank == 0:
            cam.compose_tiles_device_async(FS, [t.data_ptr() for t in gather_col[slot]], final_col.data_ptr(), stream.cuda_stream)
            cam.compose_tiles_device_async(FS, [t.data_ptr() for t in gather_ids[slot]], final_ids.data_ptr(), stream.cuda_stream)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    region_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-only: `value` ----------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(total_steps)]
    gather_done = [None, None]
    sampler = ClockSampler(local)
    launches0 = 0
    for step in range(total_steps):
        if step == Wm:
            barrier()
            sampler.start()
            launches0 = rtb.launch_count()
            region_ev[0].record(stream)
        slot = step & 1
        with torch.cuda.stream(stream):
            flush.fill_(step & 0xff)  # L2 flush between timed iterations (inside the timed region, outside the kernel's event pair)
            if gather_done[slot] is not None:
                # the slot's previous gather must have drained; rank 0 then reassembles those frames here,
                # on the render stream, where the scatter kernel has the whole GPU (beside the persistent
                # render kernel it would only get the SMs that kernel leaves free)
                stream.wait_event(gather_done[slot])
                compose(slot)
            ev[step][0].record(stream)
            if push_mode:
                obj.render_frames_push_async(cam, my_mats(step), push_ptr[slot][0], push_ptr[slot][1], stream.cuda_stream,
                                             tile_first=rank, tile_stride=world, flags=rtb.RENDER_PUSH_PREFILLED)
                ev[step][1].record(stream)
                if rank == 0:
                    # the NEXT step's frames are pre-filled with background before this step's all-reduce lets any rank
                    # start pushing into them: work units that hold nothing but background then never cross NVLink
                    cam.fill_frames_device_async(FS, push_ptr[slot ^ 1][0], push_ptr[slot ^ 1][1], stream.cuda_stream)
                dist.all_reduce(push_flag)  # completes when every rank's kernel has: the step's frames are whole on rank 0
            else:
                obj.render_frames_device_async(cam, my_mats(step), d_col[slot].data_ptr(), d_ids[slot].data_ptr(), stream.cuda_stream,
                                               tile_first=rank if tiles_mode else 0, tile_stride=world if tiles_mode else 1,
                                               flags=rtb.RENDER_TILE_MAJOR if tiles_mode else 0)
                ev[step][1].record(stream)
        if tiles_mode and not push_mode:
            side.wait_event(ev[step][1])
            with torch.cuda.stream(side):
                dist.gather(d_col[slot], gather_col[slot] if rank == 0 else None, dst=0)
                dist.gather(d_ids[slot], gather_ids[slot] if rank == 0 else None, dst=0)
                e = torch.cuda.Event()
                e.record(side)
                gather_done[slot] = e
    for slot, e in enumerate(gather_done):
        if e is not None:  # the region ends when the last frames have reached rank 0 and are reassembled
            stream.wait_event(e)
            with torch.cuda.stream(stream):
                compose(slot)
    region_ev[1].record(stream)
    barrier()
    clocks = sampler.result()
    launches = rtb.launch_count() - launches0
    kernel_ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(Wm, total_steps)]
    region = region_ev[0].elapsed_time(region_ev[1]) / 1e3  # device time of exactly K steps on this rank
    if world > 1:
        t = torch.tensor([region], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region = float(t[0])
    rays_total = K * world * F * P
    value = rays_total / region / 1e6
    ms_per_step = region / K * 1e3

    # ---------------- work counters (separate, untimed pass over the timed steps' first block) -------
    cam.counters(reset=True)
    targs = dict(tile_first=rank if tiles_mode else 0, tile_stride=world if tiles_mode else 1)
    if push_mode:
        # the last pushed step on rank 0 must equal this rank's own full render of the same frames (sanity, untimed)
        push_ok = True
        if rank == 0:
            chk_c = torch.empty(FS * P, dtype=torch.int32, device="cuda"); chk_i = torch.empty(FS * P, dtype=torch.int32, device="cuda")
            obj.render_frames_device_async(cam, my_mats(total_steps - 1), chk_c.data_ptr(), chk_i.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            got = np.empty(FS * P, np.int32)
            slot = (total_steps - 1) & 1
            rtb.memcpy_d2h(got, push_ptr[slot][1]); push_ok &= bool(np.array_equal(got, chk_i.cpu().numpy()))
            rtb.memcpy_d2h(got, push_ptr[slot][0]); push_ok &= bool(np.array_equal(got, chk_c.cpu().numpy()))
            del chk_c, chk_i
    if tiles_mode:  # the untimed passes below write whole row-major frames of this rank's tiles
        d_col = [torch.empty(FS * P, dtype=torch.int32, device="cuda")]
        d_ids = [torch.empty(FS * P, dtype=torch.int32, device="cuda")]
    obj.render_frames_device_async(cam, my_mats(Wm), d_col[0].data_ptr(), d_ids[0].data_ptr(), stream.cuda_stream, flags=rtb.RENDER_COUNTERS, **targs)
    torch.cuda.synchronize()
    c_act = cam.counters(reset=True)
    obj.render_frames_device_async(cam, my_mats(Wm), d_col[0].data_ptr(), d_ids[0].data_ptr(), stream.cuda_stream,
                                   flags=rtb.RENDER_COUNTERS | rtb.RENDER_NO_CULL, **targs)
    torch.cuda.synchronize()
    c_ref = cam.counters(reset=True)
    coverage = c_act["hits"] / max(c_act["rays"], 1)

    # ---------------- end to end through the C ABI with host buffers: `e2e` --------------------------
    h_col = torch.empty((F, P), dtype=torch.int32).pin_memory()
    h_ids = torch.empty((F, P), dtype=torch.int32).pin_memory()
    e2e_times = []
    for step in range(total_steps):
        if step == Wm:
            barrier()
        ops = rtb.orbit_ops(F, first_frame_identity=False)  # the orbit simply continues from the object's current state
        t = time.perf_counter()
        obj.render_sweep(cam, ops, out_color=h_col.numpy().view(np.uint32), out_ids=h_ids.numpy())
        e2e_times.append(time.perf_counter() - t)
    barrier()
    e2e_time = sum(e2e_times[Wm:])
    if world > 1:
        t = torch.tensor([e2e_time], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_time = float(t[0])
    e2e_value = rays_total / e2e_time / 1e6
    # the last e2e frame must equal a device-resident render of the same matrix (sanity, not timed)
    obj.render_frames_device_async(cam, obj.matrix(), d_col[0].data_ptr(), d_ids[0].data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    same = bool(torch.equal(d_ids[0][:P].cpu(), h_ids[F - 1])) and bool(torch.equal(d_col[0][:P].cpu(), h_col[F - 1]))

    # ---------------- the reference's own frame loop through the drop-in calls (one frame at a time) ----
    # WinMain.cpp:187-237: Input::set_quat + Object::transform, Object::render, Camera::color_pixels(PHONG), frame in
    # the camera's host buffer after every iteration (one device synchronisation and one colour+id D2H per frame).
    frame_loop = None
    if rank == 0 and world == 1:
        nloop = 200 if P <= (1 << 20) else 20
        for it in range(nloop + 10):
            if it == 10:
                t_loop = time.perf_counter()
            obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)
            obj.render(cam)
            cam.color_pixels(rtb.PHONG_COLOR_TAG)
        t_loop = time.perf_counter() - t_loop
        frame_loop = {"fps": nloop / t_loop, "value": nloop * P / t_loop / 1e6, "unit": "Mrays/s", "frames": nloop,
                      "api": "per frame: rtb_object_transform + rtb_object_render + rtb_camera_color_pixels(PHONG) "
                             "(the reference's WinMain loop, synchronous, frame + ids in host memory after every iteration)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (render_kernel) --------------------------------
    hbm_peak, peak_src, sm_max = measured_peaks()
    props = rtb.device_props()
    per_launch_rays = F * P
    launch_s = float(np.mean(kernel_ms)) / 1e3

    def bytes_actual(c):  # this layout: 64 B per interior record, 48 B per triangle test, hit: 48 B normal refetch; out 8 B/ray
        return 64.0 * c["nodes"] + 48.0 * c["tris"] + 48.0 * c["hits"] + 8.0 * c["rays"]

    def bytes_reference(c):  # 36*N_int + 36*N_leaf + 24*[hit] + 8, N from the reference's visit sequence
        return 36.0 * (c["boxes"] - c["tris"]) + 36.0 * c["tris"] + 24.0 * c["hits"] + 8.0 * c["rays"]

    def flops_reference(c):  # 21*N_int + 45*N_leaf + 110 + 60*[hit]
        return 21.0 * (c["boxes"] - c["tris"]) + 45.0 * c["tris"] + 110.0 * c["rays"] + 60.0 * c["hits"]

    traffic, traffic_src = None, None
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True) if os.path.isdir(os.path.join(ROOT, "profiles")) else []:
        if name.endswith("_traffic.json"):
            with open(os.path.join(ROOT, "profiles", name)) as f:
                rec = json.load(f).get(args.workload)
            if rec and rec.get("bytes"):
                traffic, traffic_src = rec["bytes"], "%s (%s)" % (name, rec.get("report"))
                break
    achieved = bytes_actual(c_act) / launch_s / 1e9
    l2_gbs = rtb.measure_l2_read_bandwidth(32 << 20, 200)  # measured here, now: the L2 leg of the roofline
    fp32_peak = props["sm_count"] * 128 * 2 * sm_max * 1e6 / 1e12
    roofline = {
        "bound": "hbm", "kernel": "rtb::render_stream_kernel<true,false>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_actual(c_act),
        "peak_source": peak_src,
        "note": "algorithmic bytes = 64 B x interior records entered + 48 B x triangle tests + 48 B x hits + 8 B x rays, counted by the kernel itself "
                "in an untimed pass; the scene is L2-resident by design, so DRAM traffic is far below this and the binding limits are L2 latency and "
                "FP32/ALU issue (see DESIGN.md); HBM copy peak used as the denominator per the bench contract",
        "bytes_per_ray": bytes_actual(c_act) / c_act["rays"],
        "l2": {"peak": l2_gbs, "unit": "GB/s", "frac": achieved / l2_gbs,
               "peak_source": "measured in this run: 16-byte L1-bypassing loads over a 32 MB L2-resident buffer (rtb_measure_l2_read_bandwidth)"},
        "reference_work": {"bytes_per_ray": bytes_reference(c_ref) / c_ref["rays"], "gb_per_s": bytes_reference(c_ref) / launch_s / 1e9,
                           "flops_per_ray": flops_reference(c_ref) / c_ref["rays"], "tflops": flops_reference(c_ref) / launch_s / 1e12,
                           "fp32_peak_tflops": fp32_peak, "fp32_frac": flops_reference(c_ref) / launch_s / 1e12 / fp32_peak},
        "per_ray": {"interior_nodes": c_act["nodes"] / c_act["rays"], "triangle_tests": c_act["tris"] / c_act["rays"],
                    "reference_node_pops": c_ref["boxes"] / c_ref["rays"], "reference_triangle_tests": c_ref["tris"] / c_ref["rays"]},
    }

    # ---------------- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rays_c, times_c, kind, cores, sample, _ = cpu_reference_run(args.workload, max(1, args.ref_frames), 2)
        cpu = {"value": rays_c / times_c[-1] / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample}

    ref_gpu = ref_seam = None
    if world == 1 and not args.no_cpu_baseline and not args.no_reference_gpu:
        ref_gpu = gpu_reference_run(args.workload, 30 if P <= (1 << 20) else 4)
        ref_seam = gpu_reference_run(args.workload, 100 if P <= (1 << 20) else 8, impl="seam")

    fps = K * world * F / region
    line = {
        "metric": "Mrays/s (primary rays, traversal + Phong)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "mesh": mesh_label, "triangles": int(len(pts)), "resolution": [W, H], "frames_per_step": F,
                   "frames_total": K * world * F, "camera": "WinMain.cpp:69-74 default, R-key quaternion step per frame", "coverage": coverage,
                   "parallelism": (("tiles x%d (scene replicated, 32x32 tiles round-robin, finished work units pushed by the render kernel into rank 0's "
                                    "frames over NVLink peer memory; pushed frames == single-GPU frames: %s)" % (world, push_ok)) if push_mode else
                                   "tiles x%d (scene replicated, 32x32 tiles round-robin, NCCL gather to rank 0 + reassembly on a side stream)" % world
                                   if tiles_mode else "frames x%d (scene replicated, blocks of %d frames per rank, no collective)" % (world, F)) if world > 1 else "single GPU",
                   "l2": "explicit flush (160 MB write) before every step; per-step working set = scene %.0f MB + %.0f MB output" % (
                       (64.0 * (len(pts) - 1) + 48.0 * len(pts)) / 1e6, F * P * 8 / 1e6),
                   "fps": fps, "fps_vs_readme_100fps": fps / README_FPS, "tree_build_s": build_s["total"], "cpu_affinity": numa},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(F * 5 * 4), "d2h_bytes_per_step": int(F * P * 8),
                "fps": K * world * F / e2e_time, "api": "rtb_render_sweep (host ops in, pinned host colour+id frames out)", "matches_device_run": same},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "reference_classes_over_librtb": ref_seam, "frame_loop": frame_loop,
        "kernel_ms": {"mean": float(np.mean(kernel_ms)), "min": float(np.min(kernel_ms)), "max": float(np.max(kernel_ms))},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dragon_orbit_960x540", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--shard", default="frames", choices=["frames", "tiles"], help="multi-GPU partition (N > 1)")
    ap.add_argument("--exchange", default="push", choices=["push", "nccl"],
                    help="tiles mode: fused peer-memory push from the render kernel (default) or NCCL gather + reassembly")
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the CPU reference arm / cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip timing the reference's own CUDA kernels on this GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_impl(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: at least 3 warm-up steps
        run_ours(args)


if __name__ == "__main__":
    main()
