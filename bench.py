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
      N * FRAMES_PER_STEP frames (weak scaling).  --exchange push (default): the render kernel stores
      every finished work unit straight into the frame's owner over NVLink peer memory, frames owned
      round-robin (frame f -> rank f % N); --exchange nccl: NCCL gather to rank 0 + reassembly.
The default line carries, beside the headline: `workloads` (N=1: the other BASELINE.json configurations,
measured the same way in the same run with fewer steps) and `tiles` (N>1: the 4K configuration
sharded by tiles with the fused peer push, and the same frames on one GPU of the same run).

`--impl reference` times the reference's own kernels compiled for the host (oracle/_ref, falling
back to the C port) on a bounded sample of the same workload with all host threads; it never loads
the product library.
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
# tiles mode (N > 1) renders N x this many frames per step (weak scaling)
# (a launch ends with the latency of its longest rays, ~0.3 ms that no number of GPUs shortens, and a step with an exchange ends
# with an all-reduce that waits for the slowest rank: a step must be long enough for that not to dominate what is measured.  At
# 4K a step is N x 36 frames -- every rank renders as many pixels per step as in frames mode, where a step is 36 frames per rank.
# Measured at N = 8 with N x 6 / 18 / 24 frames: 64 % (round 1) / 84 % / 86 % before the single-GPU kernel's re-tune, 81 % after.)
TILES_FRAMES = {"dragon_orbit_960x540": 60, "dragon_closeup_960x540": 60, "happy_orbit_3840x2160": 36, "bunny_960x540": 60, "synthetic10m_7680x4320": 1}
# the other configurations of BASELINE.json, measured beside the headline in the default N=1 run: (steps, warm-up steps)
EXTRA_WORKLOADS = {"bunny_960x540": (3, 3), "dragon_closeup_960x540": (3, 3), "happy_orbit_3840x2160": (5, 3), "synthetic10m_7680x4320": (5, 3)}
README_FPS = 100.0  # /root/reference/README.md:19 (Stanford Dragon, 960x540, unnamed GPU)
R_KEY = (0.0, 0.09950371902099893, 0.0, 0.9950371902099893)  # WinMain.cpp:187


def find_mesh(name):
    if not name:
        return None
    for d in (os.environ.get("RTB_MESH_DIR", ""), os.path.join(ROOT, "oracle", "_ref", "data")):
        if d and os.path.exists(os.path.join(d, name)):
            return os.path.join(d, name)
    return None


def mesh_label(workload):
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[workload]
    path = find_mesh(fname)
    if path:
        return os.path.basename(path)
    return "stand-in: displaced geodesic icosphere nu=%d (%d triangles); %s absent from the reference checkout" % (nu, 20 * nu * nu, fname or "no file")


def load_points(rtb, workload):
    """Mesh of a workload through the PRODUCT's loader / generator (our arm)."""
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[workload]
    path = find_mesh(fname)
    return rtb.read_ply(path, mode) if path else rtb.geodesic_mesh(nu)


def load_points_reference(workload):
    """The same mesh WITHOUT the product (reference arm, cpu_baseline): the oracle's restatement of the reference
    loader, or oracle/standin.py, which tests/test_host_cpu.py pins bit for bit to rtb_mesh_geodesic."""
    from oracle import orc, standin
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[workload]
    path = find_mesh(fname)
    return orc.read_ply(path, mode) if path else standin.geodesic_mesh(nu)


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


def profile_record(workload):
    """DRAM bytes and executed warp instructions of ONE render launch of `workload` (a whole bench step) from the newest
    committed ncu capture: profiles/*_traffic.json, written by tools/profile_summary.py from `ncu --set full`."""
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir), reverse=True) if os.path.isdir(pdir) else []:
        if name.endswith("_traffic.json"):
            with open(os.path.join(pdir, name)) as f:
                rec = json.load(f).get(workload)
            if rec and rec.get("bytes"):
                return rec, "%s (%s)" % (name, rec.get("report"))
    return None, None


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (never touches the product library)
# --------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(workload, frames, repeats, prefer_ref=True):
    """Time the reference's own CPU-executed kernels (oracle/_ref) or the C port on `frames` frames of the
    workload, on every host core this process may use.  Returns (rays, seconds per repeat, kind, cores, sample text)."""
    from oracle import orc, refemu
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[workload]
    pts = load_points_reference(workload)
    cam = orc.default_camera(W, H)
    # (beyond ~2 M triangles the reference's own host build takes minutes: the C port, which builds with all cores, stands in)
    kind = "reference" if (prefer_ref and refemu.available() and len(pts) <= 2_000_000) else "port"
    # torchrun exports OMP_NUM_THREADS=1: ask for all cores explicitly (round 1's N>1 arm ran on one core)
    threads = host_threads()
    if kind == "reference":
        refemu.lib().ref_set_threads(threads)
    else:
        orc.lib().orc_set_threads(threads)
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
            scene.transform(10, *R_KEY)
        times.append(time.perf_counter() - t)
    sample = "%d consecutive orbit frames of %s at %dx%d per step (traversal + shading, tree build %.1f s excluded)" % (frames, workload, W, H, build_s)
    return frames * W * H, times, kind, cores, sample


def gpu_reference_run(workload, frames, impl="cuda_fmad"):
    """The reference's own CUDA kernels (Trixel.cu / Camera.cu compiled by nvcc for sm_100a with the project's default
    code generation, oracle/_ref/libref_cuda_fmad.so) on this GPU, on `frames` consecutive frames of the workload, timed
    the way the reference times itself (wall clock per loop iteration, WinMain.cpp:219-228) but WITHOUT its window blit,
    console output and second color_pixels call -- a lower bound of its per-frame cost: "the kernel to beat" on this box."""
    from oracle import orc, refemu
    if not refemu.available(impl):
        return None
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[workload]
    pts = load_points_reference(workload)
    if len(pts) > 2_000_000:  # the reference's host-side tree build alone would take minutes (80-byte records x 6 lists)
        return {"skipped": "the reference's own single-threaded tree build is not practical for %d triangles" % len(pts)}
    t0 = time.time()
    scene = refemu.RefScene(W, H, orc.default_camera(W, H), points9=pts, impl=impl, objects=2 if impl == "seam" else 1)
    build_s = time.time() - t0
    n = np.array([0.0, 0.0, 1.0], np.float32)
    for _ in range(zoom):
        scene.transform(32, float(n[0]), float(n[1]), float(n[2]), 0.005)
    per_frame = []
    for f in range(frames + 3):
        t = time.perf_counter()
        scene.render_nocopy()  # Object::render + Camera::color_pixels (kernels, device syncs, colour buffer D2H)
        scene.transform(10, *R_KEY)
        per_frame.append(time.perf_counter() - t)
    total = sum(per_frame[3:])
    kind = ("reference CUDA kernels, nvcc sm_100a, default fmad" if impl == "cuda_fmad" else
            "the reference's own host classes (two objects registered, WinMain.cpp:152-156) with its .cu files replaced by "
            "integration/rtb_seam.cpp over librtb.so")
    return {"value": frames * W * H / total / 1e6, "unit": "Mrays/s", "fps": frames / total, "kind": kind,
            "sample": "%d consecutive orbit frames after 3 warm-up frames, wall clock around Object::render + Camera::color_pixels "
                      "(its host tree build took %.1f s)" % (frames, build_s)}


def run_reference_impl(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fname, mode, nu, W, H, fps, zoom = WORKLOADS[args.workload]
    frames = max(1, args.ref_frames)
    rays, times, kind, cores, sample = cpu_reference_run(args.workload, frames, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = rays * len(timed) / total / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s (primary rays, traversal + Phong)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(timed) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "mesh": mesh_label(args.workload), "resolution": [W, H], "frames_per_step": frames,
                   "fps": frames * len(timed) / total},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "product_library_loaded": "cpp_cuda_raytracer_dev_b200" in sys.modules,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class Env:
    """Process-wide state of our arm: torch, torch.distributed, the binding, rank layout."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import cpp_cuda_raytracer_dev_b200 as rtb
        self.torch, self.dist, self.rtb = torch, dist, rtb
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available() or rtb.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device -- the ray-cast path has no CPU fallback")
        torch.cuda.set_device(self.local)
        rtb.set_device(self.local)
        self.numa = None
        if self.world > 1:
            # one process per GPU: run on (and first-touch the pinned frame buffers from) the CPUs next to this GPU, so that
            # eight ranks' 50 GB/s DMA streams do not all cross the socket interconnect
            try:
                import pynvml
                pynvml.nvmlInit()
                mask_words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.local), (os.cpu_count() + 63) // 64)
                cpus = {64 * w + b for w, word in enumerate(mask_words) for b in range(64) if (word >> b) & 1}
                cpus &= os.sched_getaffinity(0)
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    self.numa = "%d CPUs local to GPU %d" % (len(cpus), self.local)
            except Exception as exc:  # affinity is an optimisation, never a requirement
                self.numa = "not set (%s)" % type(exc).__name__
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, seconds):
        if self.world == 1:
            return seconds
        t = self.torch.tensor([seconds], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])


class Scene:
    """One workload on this rank's GPU: mesh, tree, camera, object, and the orbit's matrices."""

    def __init__(self, env, workload, frames_per_step=0):
        rtb = env.rtb
        self.env, self.workload = env, workload
        fname, mode, nu, self.W, self.H, self.F, self.zoom = WORKLOADS[workload]
        if frames_per_step:
            self.F = frames_per_step
        self.P = self.W * self.H
        self.pts = load_points(rtb, workload)
        self.mesh = rtb.Trixel(self.pts)
        self.mesh.create_kd()
        self.build_s = self.mesh.build_seconds()
        self.cam = rtb.Camera(self.W, self.H, **rtb.default_camera_args(self.W, self.H))
        self.obj = rtb.Object(self.mesh)
        self.cam.add_object(self.obj)
        nvec = self.cam.basis()[0:3]
        for _ in range(self.zoom):
            self.obj.transform((float(nvec[0]), float(nvec[1]), float(nvec[2]), 0.005), rtb.TRANSLATE_Z)

    def orbit(self, count):
        """Matrices of `count` consecutive orbit frames from the object's current state (host recurrence, the
        reference's Object::transform): identical on every rank."""
        rtb = self.env.rtb
        mats = np.empty((count, 12), np.float32)
        mats[0] = self.obj.matrix()
        for g in range(1, count):
            mats[g] = self.obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)
        return mats

    def close(self):
        self.obj.close(); self.cam.close(); self.mesh.close()


def kernel_only(env, sc, K, Wm, shard="frames", exchange="push", sample_clocks=True):
    """`value`: K timed steps after Wm warm-up steps, everything resident in HBM, CUDA events on the launching stream, L2
    flushed before every step.  Returns a dict (value in Mrays/s over all ranks, ms_per_step, kernel_ms, launches, ...)."""
    torch, dist, rtb = env.torch, env.dist, env.rtb
    rank, world = env.rank, env.world
    W, H, P, F = sc.W, sc.H, sc.P, sc.F
    cam, obj = sc.cam, sc.obj
    total_steps = K + Wm
    tiles_mode = world > 1 and shard == "tiles"
    push_mode = tiles_mode and exchange == "push"
    if tiles_mode and not push_mode:  # the NCCL gather runs beside the persistent render kernel: leave it a few SMs
        rtb.set_knob("reserve_sms", int(os.environ.get("RTB_RESERVE_SMS", "8")))
    FS = F * world if tiles_mode else F  # frames a rank touches per step
    # Orbit: global frame g of step s on rank r is frame (s*world + r)*F + j (frames mode); in tiles mode every rank renders
    # its tiles of all N*F frames of the step.
    mats = sc.orbit(total_steps * world * F)

    def my_mats(step):
        if tiles_mode:
            return mats[step * world * F:(step + 1) * world * F]
        b = (step * world + rank) * F
        return mats[b:b + F]

    stream = torch.cuda.Stream()
    side = torch.cuda.Stream()
    PE = cam.tile_major_elements(world) if tiles_mode else P
    d_col = d_ids = None
    if not push_mode:
        d_col = [torch.empty(FS * PE, dtype=torch.int32, device="cuda") for _ in range(2)]
        d_ids = [torch.empty(FS * PE, dtype=torch.int32, device="cuda") for _ in range(2)]
    gather_col = gather_ids = final_col = final_ids = None
    owned = push_ptrs = push_flag = my_bufs = None
    if push_mode:
        # Striped ownership: frame f of a step belongs to rank f % N (its frame f // N).  Every rank owns two slots of
        # FS/N final frames (colours + ids), maps everybody else's, and its render kernel stores finished work units
        # straight into the owners' frames over NVLink -- no gather, no receive buffers, no reassembly pass.
        owned = FS // world
        my_bufs = [rtb.PeerBuffer(4 * owned * P) for _ in range(4)]  # slot 0: colours, ids; slot 1: colours, ids
        handles = [None] * world
        dist.all_gather_object(handles, [b.handle() for b in my_bufs])
        push_ptrs = [[None] * world for _ in range(4)]
        for r in range(world):
            for k in range(4):
                push_ptrs[k][r] = my_bufs[k].ptr if r == rank else rtb.peer_open(handles[r][k])
        push_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        cam.fill_frames_device_async(owned, my_bufs[0].ptr, my_bufs[1].ptr, stream.cuda_stream)
        env.barrier()
    elif tiles_mode and rank == 0:
        gather_col = [[torch.empty(FS * PE, dtype=torch.int32, device="cuda") for _ in range(world)] for _ in range(2)]
        gather_ids = [[torch.empty(FS * PE, dtype=torch.int32, device="cuda") for _ in range(world)] for _ in range(2)]
        final_col = torch.empty(FS * P, dtype=torch.int32, device="cuda")
        final_ids = torch.empty(FS * P, dtype=torch.int32, device="cuda")

    def compose(slot):
        if tiles_mode and not push_mode and rank == 0:
            cam.compose_tiles_device_async(FS, [t.data_ptr() for t in gather_col[slot]], final_col.data_ptr(), stream.cuda_stream)
            cam.compose_tiles_device_async(FS, [t.data_ptr() for t in gather_ids[slot]], final_ids.data_ptr(), stream.cuda_stream)

    region_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(total_steps)]
    gather_done = [None, None]
    allreduce_done = [None]
    sampler = ClockSampler(env.local) if sample_clocks else None
    launches0 = 0
    for step in range(total_steps):
        if step == Wm:
            env.barrier()
            if sampler:
                sampler.start()
            launches0 = rtb.launch_count()
            region_ev[0].record(stream)
        slot = step & 1
        with torch.cuda.stream(stream):
            env.flush.fill_(step & 0xff)  # L2 flush between timed iterations (inside the timed region, outside the kernel's event pair)
            if gather_done[slot] is not None:
                # the slot's previous gather must have drained; rank 0 then reassembles those frames here,
                # on the render stream, where the scatter kernel has the whole GPU (beside the persistent
                # render kernel it would only get the SMs that kernel leaves free)
                stream.wait_event(gather_done[slot])
                compose(slot)
            ev[step][0].record(stream)
            if push_mode:
                # The NEXT step's frames of this owner are pre-filled with background before this step's all-reduce lets any
                # rank start pushing into them: work units that hold nothing but background then never cross NVLink.  The
                # other slot has been free since the previous step's all-reduce, so the fill -- a few thread blocks per SM on
                # a second stream, queued IN FRONT of the persistent render kernel -- runs beside this step's rendering.
                if allreduce_done[0] is not None:
                    side.wait_event(allreduce_done[0])
                cam.fill_frames_device_async(owned, my_bufs[2 * (slot ^ 1)].ptr, my_bufs[2 * (slot ^ 1) + 1].ptr, side.cuda_stream)
                fill_done = torch.cuda.Event(); fill_done.record(side)
                ev[step][0].record(stream)  # (again: the kernel's event pair starts after the fill has been queued)
                obj.render_frames_push_striped_async(cam, my_mats(step), push_ptrs[2 * slot], push_ptrs[2 * slot + 1], stream.cuda_stream,
                                                     tile_first=rank, tile_stride=world, flags=rtb.RENDER_PUSH_PREFILLED)
                ev[step][1].record(stream)
                stream.wait_event(fill_done)
                dist.all_reduce(push_flag)  # completes when every rank's kernel has: the step's frames are whole on their owners
                allreduce_done[0] = torch.cuda.Event(); allreduce_done[0].record(stream)
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
    env.barrier()
    clocks = sampler.result() if sampler else None
    launches = rtb.launch_count() - launches0
    kernel_ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(Wm, total_steps)]
    region = env.max_over_ranks(region_ev[0].elapsed_time(region_ev[1]) / 1e3)  # device time of exactly K steps, max over ranks
    rays_total = K * world * F * P
    out = {"value": rays_total / region / 1e6, "ms_per_step": region / K * 1e3, "fps": K * world * F / region, "kernel_ms": kernel_ms,
           "launches": int(launches), "clocks": clocks, "region_s": region, "frames_per_rank_step": FS, "rays_total": rays_total,
           "first_timed_mats": my_mats(Wm), "last_mats": my_mats(total_steps - 1)}
    if push_mode:
        # the last pushed step: this rank's owned frames must equal its own full render of the same frames (sanity, untimed)
        slot = (total_steps - 1) & 1
        mine = my_mats(total_steps - 1)[rank::world]
        chk_c = torch.empty(owned * P, dtype=torch.int32, device="cuda"); chk_i = torch.empty(owned * P, dtype=torch.int32, device="cuda")
        obj.render_frames_device_async(cam, mine, chk_c.data_ptr(), chk_i.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        got = np.empty(owned * P, np.int32)
        rtb.memcpy_d2h(got, my_bufs[2 * slot + 1].ptr); ok = bool(np.array_equal(got, chk_i.cpu().numpy()))
        rtb.memcpy_d2h(got, my_bufs[2 * slot].ptr); ok &= bool(np.array_equal(got, chk_c.cpu().numpy()))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["pushed_frames_equal_single_gpu_frames"] = bool(int(flag[0]))
        env.barrier()
        for r in range(world):
            if r != rank:
                for k in range(4):
                    rtb.peer_close(push_ptrs[k][r])
        env.barrier()
        for b in my_bufs:
            b.close()
    return out


def count_work(env, sc, mats):
    """Work counters of one step's frames (separate, untimed passes): the kernel's own counts and, with culling off, the
    reference's visit sequence."""
    torch, rtb = env.torch, env.rtb
    n = len(mats)
    d_col = torch.empty(n * sc.P, dtype=torch.int32, device="cuda"); d_ids = torch.empty(n * sc.P, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    sc.cam.counters(reset=True)
    sc.obj.render_frames_device_async(sc.cam, mats, d_col.data_ptr(), d_ids.data_ptr(), s, flags=rtb.RENDER_COUNTERS)
    torch.cuda.synchronize()
    c_act = sc.cam.counters(reset=True)
    sc.obj.render_frames_device_async(sc.cam, mats, d_col.data_ptr(), d_ids.data_ptr(), s, flags=rtb.RENDER_COUNTERS | rtb.RENDER_NO_CULL)
    torch.cuda.synchronize()
    c_ref = sc.cam.counters(reset=True)
    return c_act, c_ref


def end_to_end(env, sc, K, Wm):
    """`e2e`: the same metric through the C-ABI call rtb_render_sweep with HOST buffers -- transform ops in, every frame's
    colour + hit-id buffer in pinned host memory when the call returns; wall clock, max over ranks.  Also the plain D2H
    copy of the same bytes by all ranks at once: what the host-memory path of this box can take."""
    torch, rtb = env.torch, env.rtb
    F, P = sc.F, sc.P
    h_col = torch.empty((F, P), dtype=torch.int32).pin_memory()
    h_ids = torch.empty((F, P), dtype=torch.int32).pin_memory()
    times = []
    for step in range(K + Wm):
        if step == Wm:
            env.barrier()
        ops = rtb.orbit_ops(F, first_frame_identity=False)  # the orbit simply continues from the object's current state
        t = time.perf_counter()
        sc.obj.render_sweep(sc.cam, ops, out_color=h_col.numpy().view(np.uint32), out_ids=h_ids.numpy())
        times.append(time.perf_counter() - t)
    env.barrier()
    total = env.max_over_ranks(sum(times[Wm:]))
    rays_total = K * env.world * F * P
    # the last e2e frame must equal a device-resident render of the same matrix (sanity, not timed)
    d_col = torch.empty(P, dtype=torch.int32, device="cuda"); d_ids = torch.empty(P, dtype=torch.int32, device="cuda")
    sc.obj.render_frames_device_async(sc.cam, sc.obj.matrix(), d_col.data_ptr(), d_ids.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    same = bool(torch.equal(d_ids.cpu(), h_ids[F - 1])) and bool(torch.equal(d_col.cpu(), h_col[F - 1]))
    # copy-only: F frames' colours + ids, device -> the same pinned buffers, every rank at the same time
    src_c = torch.empty(F * P, dtype=torch.int32, device="cuda"); src_i = torch.empty(F * P, dtype=torch.int32, device="cuda")
    copy_times = []
    for rep in range(4):
        env.barrier()
        t = time.perf_counter()
        h_col.view(-1).copy_(src_c, non_blocking=True); h_ids.view(-1).copy_(src_i, non_blocking=True)
        torch.cuda.synchronize()
        copy_times.append(time.perf_counter() - t)
    copy_s = env.max_over_ranks(min(copy_times[1:]))
    return {"value": rays_total / total / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(F * 5 * 4), "d2h_bytes_per_step": int(F * P * 8),
            "fps": K * env.world * F / total,
            "api": "rtb_render_sweep (host ops in, pinned host colour+id frames out): host threads pre-fill each chunk of frames with the background while "
                   "the kernel traces the chunk before and stores every work unit that holds anything else straight into the caller's buffers over PCIe; "
                   "RTB_SWEEP_DIRECT=0 selects the copy-engine path (device frames + cudaMemcpyAsync), which is what pageable buffers get",
            "matches_device_run": same, "host_fill_gbs": rtb.measure_host_fill_bandwidth(256 << 20, 0) if env.rank == 0 else None,
            "host_fill_note": "what this host's cores write into pinned memory with streaming stores (all usable cores, nothing else running): the ceiling of the "
                              "pre-fill; d2h_bytes_per_step counts every byte of the frames delivered to host memory, whoever wrote it",
            "d2h_gbs": K * env.world * F * P * 8 / total / 1e9,
            "d2h_copy_only_gbs": env.world * F * P * 8 / copy_s / 1e9,
            "d2h_copy_only_note": "plain cudaMemcpyAsync of one step's frames (colour + ids) from device to the same pinned buffers, all ranks "
                                  "concurrently, max over ranks: the ceiling of any frame-to-host figure on this box"}


def frame_loop(env, sc):
    """The reference's own frame loop through the drop-in calls, one frame at a time (WinMain.cpp:187-237: Input::set_quat +
    Object::transform, Object::render, Camera::color_pixels(PHONG)); frame and ids in the camera's host buffers after every
    iteration.  Measured twice: as shipped (frames rendered ahead while the steps repeat, which they do in this loop exactly as
    with the reference's R key held down) and with the knob `lookahead` at 0 (every frame starts when it is asked for)."""
    rtb = env.rtb
    nloop = 300 if sc.P <= (1 << 20) else 30
    out = {}
    for name, look in (("fps", None), ("fps_no_lookahead", 0)):
        if look is not None:
            rtb.set_knob("lookahead", look)
        for it in range(nloop + 10):
            if it == 10:
                t_loop = time.perf_counter()
            sc.obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)
            sc.obj.render(sc.cam)
            sc.cam.color_pixels(rtb.PHONG_COLOR_TAG)
        out[name] = nloop / (time.perf_counter() - t_loop)
    rtb.set_knob("lookahead", 2)
    out.update({"value": out["fps"] * sc.P / 1e6, "unit": "Mrays/s", "frames": nloop,
                "api": "per frame: rtb_object_transform + rtb_object_render + rtb_camera_color_pixels(PHONG) (the reference's WinMain loop: "
                       "frame + ids in host memory after every iteration, one synchronisation per frame; the kernel stores the frame straight into "
                       "a pinned host buffer, shares long rays between the lanes of a warp, and -- `fps` -- up to two predicted frames are in "
                       "flight behind the current one while the transform steps repeat; `fps_no_lookahead`: knob lookahead = 0)"})
    return out


def scene_extension_leg(env, sc):
    """The scene kernel (csrc/rtb_scene.cuh) on the headline mesh: a second object beside the
    first, two lights, shadow rays, then 2x2 rays per pixel -- frames per second, kernel-only (CUDA events)."""
    torch, rtb = env.torch, env.rtb
    second = rtb.Object(sc.mesh)
    sc.cam.add_object(second)
    u = sc.cam.basis()[6:9]
    for _ in range(14):
        second.transform((float(u[0]), float(u[1]), float(u[2]), 0.012), rtb.TRANSLATE_X)
    d_c = torch.empty(sc.P, dtype=torch.int32, device="cuda"); d_i = torch.empty(sc.P, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    out = {}
    for name, lights, shadows, samples in (("two_objects", [(2, 2, 2)], False, 0), ("two_objects_two_lights_shadows", [(2, 2, 2), (-1.5, 1, -2)], True, 0),
                                           ("two_objects_two_lights_shadows_2x2_samples", [(2, 2, 2), (-1.5, 1, -2)], True, 2)):
        sc.cam.set_lights(lights); sc.cam.set_shadows(shadows); sc.cam.set_sample_rate(samples)
        ms = []
        for rep in range(24):
            sc.obj.transform(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record(st)
                sc.cam.render_scene_device_async(d_c.data_ptr(), d_i.data_ptr(), st.cuda_stream)
                e1.record(st)
            torch.cuda.synchronize()
            if rep >= 4:
                ms.append(e0.elapsed_time(e1))
        hits = int((d_i >= 0).sum())
        out[name] = {"ms_per_frame": float(np.mean(ms)), "fps": 1e3 / float(np.mean(ms)), "rays_per_pixel": max(1, samples) ** 2,
                     "coverage": hits / sc.P, "second_object_pixels": int((d_i >= len(sc.pts)).sum())}
    sc.cam.set_lights([(2, 2, 2)]); sc.cam.set_shadows(False); sc.cam.set_sample_rate(0)
    second.close()
    out["note"] = ("one launch per frame of the general scene kernel (one thread per pixel); primary rays of the default scene go through the "
                   "persistent kernel instead (value / e2e above)")
    return out


def roofline_of(env, sc, workload, kernel_ms, c_act, c_ref, clocks):
    """Roofline of the dominant kernel (render_stream_kernel) for one launch = one bench step.

    Three legs, each a fraction of a measured peak, and `bound` / `frac` / `achieved` / `peak` are those of the BINDING one:
      hbm   : DRAM bytes of the launch (ncu dram__bytes_read + write of the committed capture of this very command) / live
              launch time, against the measured HBM copy bandwidth;
      l2    : the kernel's own algorithmic bytes (64 B per interior record entered + 48 B per triangle test + 48 B per hit
              + 8 B per ray, counted live by the kernel) / live launch time, against the L2 read bandwidth measured in this run;
      issue : warp instructions of the launch (ncu smsp__inst_executed of the same capture; the instruction stream of a
              launch is deterministic) / (live launch time x SMs x 4 schedulers x SM clock sampled live)."""
    rtb = env.rtb
    hbm_peak, peak_src, sm_max = measured_peaks()
    props = rtb.device_props()
    launch_s = float(np.mean(kernel_ms)) / 1e3

    def bytes_layout(c):  # this layout: 64 B per interior record, 48 B per triangle test, hit: 48 B normal refetch; out 8 B/ray
        return 64.0 * c["nodes"] + 48.0 * c["tris"] + 48.0 * c["hits"] + 8.0 * c["rays"]

    def bytes_survey(c):  # 36*N_int + 36*N_leaf + 24*[hit] + 8 per ray
        return 36.0 * (c["boxes"] - c["tris"]) + 36.0 * c["tris"] + 24.0 * c["hits"] + 8.0 * c["rays"]

    def flops_survey(c):  # 21*N_int + 45*N_leaf + 110 + 60*[hit]
        return 21.0 * (c["boxes"] - c["tris"]) + 45.0 * c["tris"] + 110.0 * c["rays"] + 60.0 * c["hits"]

    prof, prof_src = profile_record(workload)
    l2_gbs = rtb.measure_l2_read_bandwidth(32 << 20, 200)  # measured here, now
    fp32_peak = props["sm_count"] * 128 * 2 * sm_max * 1e6 / 1e12
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
    scale_counts = sc.F / float(max(1, c_act["rays"] // sc.P))  # the counter pass may cover fewer frames than a launch
    legs = {}
    algo = bytes_layout(c_act) * scale_counts
    legs["l2"] = {"achieved": algo / launch_s / 1e9, "peak": l2_gbs, "unit": "GB/s", "frac": algo / launch_s / 1e9 / l2_gbs,
                  "peak_source": "measured in this run: 16-byte L1-bypassing loads over a 32 MB L2-resident buffer (rtb_measure_l2_read_bandwidth)",
                  "bytes_per_launch": algo}
    traffic = None
    if prof:
        same_launch = prof.get("frames_per_launch") == sc.F
        scale = 1.0 if same_launch else sc.F / float(prof.get("frames_per_launch") or sc.F)
        traffic = prof["bytes"] * scale
        legs["hbm"] = {"achieved": traffic / launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": traffic / launch_s / 1e9 / hbm_peak,
                       "peak_source": peak_src, "bytes_per_launch": traffic,
                       "compulsory_bytes_per_launch": 8.0 * sc.F * sc.P + 64.0 * (len(sc.pts) - 1) + 48.0 * len(sc.pts),
                       "source": prof_src + ("" if same_launch else " scaled to %d frames per launch" % sc.F)}
        if prof.get("inst_executed"):
            inst = prof["inst_executed"] * scale
            slots = launch_s * props["sm_count"] * 4 * sm_mhz * 1e6
            legs["issue"] = {"achieved": inst / launch_s / 1e9, "peak": props["sm_count"] * 4 * sm_mhz * 1e6 / 1e9, "unit": "G warp-inst/s",
                             "frac": inst / slots, "warp_instructions_per_launch": inst, "sm_mhz": sm_mhz,
                             "ncu_issue_slots_busy_pct": prof.get("issue_slots_busy_pct"),
                             "ncu_threads_per_instruction": prof.get("threads_per_inst"), "source": prof_src}
    bound = max(legs, key=lambda k: legs[k]["frac"])
    b = legs[bound]
    return {
        "bound": bound, "kernel": "rtb::render_stream_kernel<true,false,false>", "achieved": b["achieved"], "peak": b["peak"], "unit": b["unit"],
        "frac": b["frac"], "traffic": traffic, "traffic_source": prof_src, "legs": legs,
        "note": "latency / issue bound gather: the scene is L2-resident, DRAM carries little more than the frames; frac is the largest of "
                "the three legs (see legs), none of which is a tensor-core or TMA leg -- the path is not a dense contraction",
        "bytes_per_ray": {"this_layout_kernel_counts": bytes_layout(c_act) / c_act["rays"], "survey_8d_kernel_counts": bytes_survey(c_act) / c_act["rays"],
                          "survey_8d_reference_counts": bytes_survey(c_ref) / c_ref["rays"]},
        "survey_8d": {"kernel_counts_gb_per_s": bytes_survey(c_act) * scale_counts / launch_s / 1e9,
                      "reference_counts_gb_per_s": bytes_survey(c_ref) * scale_counts / launch_s / 1e9,
                      "reference_counts_frac_of_l2": bytes_survey(c_ref) * scale_counts / launch_s / 1e9 / l2_gbs,
                      "flops_per_ray_reference_counts": flops_survey(c_ref) / c_ref["rays"],
                      "tflops_reference_counts": flops_survey(c_ref) * scale_counts / launch_s / 1e12, "fp32_peak_tflops": fp32_peak,
                      "fp32_frac_reference_counts": flops_survey(c_ref) * scale_counts / launch_s / 1e12 / fp32_peak},
        "per_ray": {"interior_nodes": c_act["nodes"] / c_act["rays"], "triangle_tests": c_act["tris"] / c_act["rays"],
                    "reference_node_pops": c_ref["boxes"] / c_ref["rays"], "reference_triangle_tests": c_ref["tris"] / c_ref["rays"],
                    "deepest_stack_mean": c_act["stack_depth_sum"] / c_act["rays"], "deepest_stack_max": c_act["stack_depth_max"]},
    }


def run_ours(args):
    env = Env(args)
    rtb, torch, dist = env.rtb, env.torch, env.dist
    rank, world = env.rank, env.world
    args.gpus = world
    K, Wm = args.steps, args.warmup
    tiles_mode = world > 1 and args.shard == "tiles"
    F = args.frames_per_step or (TILES_FRAMES[args.workload] if tiles_mode else 0)
    sc = Scene(env, args.workload, F)
    W, H, P, F = sc.W, sc.H, sc.P, sc.F

    head = kernel_only(env, sc, K, Wm, shard=args.shard, exchange=args.exchange)
    c_act, c_ref = count_work(env, sc, head["first_timed_mats"][:F])
    coverage = c_act["hits"] / max(c_act["rays"], 1)
    e2e = end_to_end(env, sc, K, Wm)
    loop = frame_loop(env, sc) if (rank == 0 and world == 1) else None
    scene_ext = scene_extension_leg(env, sc) if (rank == 0 and world == 1 and not args.no_other_workloads) else None

    # ---- N > 1, sweep sharded by frames: the path's one real exchange measured beside it -- the 4K configuration split by
    # tiles, finished work units pushed by the render kernel into the frames' owners (striped) over NVLink ------------------
    tiles = frames_4k = None
    if world > 1 and args.shard == "frames" and not args.no_tiles_leg:
        tw = "happy_orbit_3840x2160"
        ts = Scene(env, tw, TILES_FRAMES[tw])
        t = kernel_only(env, ts, 5, 3, shard="tiles", exchange="push", sample_clocks=False)
        # the same N*F frames of a step on ONE GPU (rank 0's) of this very run: the 100 % mark
        single = torch.tensor([0.0], dtype=torch.float64, device="cuda")
        if rank == 0:
            FS = TILES_FRAMES[tw] * world
            d_c = torch.empty(FS * ts.P, dtype=torch.int32, device="cuda"); d_i = torch.empty(FS * ts.P, dtype=torch.int32, device="cuda")
            st = torch.cuda.Stream()
            ms = []
            for rep in range(6):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(st):
                    env.flush.fill_(rep)
                    e0.record(st)
                    ts.obj.render_frames_device_async(ts.cam, t["last_mats"], d_c.data_ptr(), d_i.data_ptr(), st.cuda_stream)
                    e1.record(st)
                torch.cuda.synchronize()
                if rep >= 2:
                    ms.append(e0.elapsed_time(e1))
            single[0] = float(np.mean(ms))
            del d_c, d_i
        dist.broadcast(single, src=0)
        one_gpu_ms = float(single[0])
        one_gpu_value = TILES_FRAMES[tw] * world * ts.P / one_gpu_ms / 1e3
        tiles = {"workload": tw, "mesh": mesh_label(tw), "resolution": [ts.W, ts.H], "frames_per_step": TILES_FRAMES[tw] * world,
                 "value": t["value"], "unit": "Mrays/s", "fps": t["fps"], "ms_per_step": t["ms_per_step"],
                 "one_gpu_same_frames_ms": one_gpu_ms, "one_gpu_value": one_gpu_value,
                 "speedup_vs_one_gpu": t["value"] / one_gpu_value, "efficiency_vs_one_gpu": t["value"] / one_gpu_value / world,
                 "pushed_frames_equal_single_gpu_frames": t.get("pushed_frames_equal_single_gpu_frames"),
                 "exchange": "fused into the render kernel: finished 32x4-pixel work units stored as 128-byte rows into the frame's owner "
                             "(frame f -> rank f % N) over NVLink peer memory; one 4-byte all-reduce per step is the only collective",
                 "scaling": "strong (one step's frames split by tiles over N GPUs against the same frames on one GPU)"}
        # the same 4K configuration sharded by FRAMES (36 per rank and step, no collective): BASELINE.json's "near-linear 8-GPU scaling
        # at 4K" in the driver's own record; its N=1 figure is workloads.happy_orbit_3840x2160.value of the N=1 line
        f4 = kernel_only(env, ts, 3, 3, shard="frames", sample_clocks=False)
        frames_4k = {"workload": tw, "resolution": [ts.W, ts.H], "frames_per_rank_and_step": ts.F, "value": f4["value"], "unit": "Mrays/s", "fps": f4["fps"],
                     "ms_per_step": f4["ms_per_step"], "per_gpu_value": f4["value"] / world, "scaling": "weak (blocks of frames per rank, no collective)"}
        ts.close()

    if rank != 0:
        # the other ranks take part in the collective legs above; only rank 0 reports
        if world > 1:
            dist.destroy_process_group()
        return

    roofline = roofline_of(env, sc, args.workload, head["kernel_ms"], c_act, c_ref, head["clocks"])

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------------------
    cpu = ref_gpu = ref_seam = None
    if world == 1 and not args.no_cpu_baseline:
        rays_c, times_c, kind, cores, sample = cpu_reference_run(args.workload, max(1, args.ref_frames), 2)
        cpu = {"value": rays_c / times_c[-1] / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample}
        if not args.no_reference_gpu:
            ref_gpu = gpu_reference_run(args.workload, 30 if P <= (1 << 20) else 4)
            ref_seam = gpu_reference_run(args.workload, 100 if P <= (1 << 20) else 8, impl="seam")

    # ---- the other configurations of BASELINE.json, same measurement, fewer steps (N = 1 default run) -----------------------
    others = None
    n_tri = int(len(sc.pts))
    build_total = sc.build_s["total"]
    if world == 1 and not args.no_other_workloads and args.workload == "dragon_orbit_960x540":
        others = {}
        sc.close()
        for name, (k2, w2) in EXTRA_WORKLOADS.items():
            o = Scene(env, name)
            h2 = kernel_only(env, o, k2, w2, sample_clocks=False)
            ca, cr = count_work(env, o, h2["first_timed_mats"][:min(o.F, 8)])
            e2 = end_to_end(env, o, 2, 1)
            l2 = frame_loop(env, o)
            others[name] = {"mesh": mesh_label(name), "triangles": int(len(o.pts)), "resolution": [o.W, o.H], "frames_per_step": o.F, "steps": k2,
                            "value": h2["value"], "unit": "Mrays/s", "fps": h2["fps"], "ms_per_step": h2["ms_per_step"],
                            "coverage": ca["hits"] / max(ca["rays"], 1),
                            "e2e": {"value": e2["value"], "fps": e2["fps"], "d2h_gbs": e2["d2h_gbs"], "matches_device_run": e2["matches_device_run"]},
                            "frame_loop_fps": l2["fps"], "tree_build_s": o.build_s["total"],
                            "per_ray": {"interior_nodes": ca["nodes"] / ca["rays"], "triangle_tests": ca["tris"] / ca["rays"],
                                        "reference_node_pops": cr["boxes"] / cr["rays"], "reference_triangle_tests": cr["tris"] / cr["rays"]}}
            o.close()

    kernel_ms = head["kernel_ms"]
    par = "single GPU"
    if world > 1:
        if tiles_mode and args.exchange == "push":
            par = ("tiles x%d (scene replicated, 32x32 tiles round-robin, finished work units pushed by the render kernel into the frames' owners "
                   "(frame f -> rank f %% N) over NVLink peer memory; pushed frames == single-GPU frames: %s)" % (world, head.get("pushed_frames_equal_single_gpu_frames")))
        elif tiles_mode:
            par = "tiles x%d (scene replicated, 32x32 tiles round-robin, NCCL gather to rank 0 + reassembly on a side stream)" % world
        else:
            par = "frames x%d (scene replicated, blocks of %d frames per rank, no collective)" % (world, F)
    line = {
        "metric": "Mrays/s (primary rays, traversal + Phong)", "value": head["value"], "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "mesh": mesh_label(args.workload), "triangles": n_tri, "resolution": [W, H], "frames_per_step": F,
                   "frames_total": K * world * F, "camera": "WinMain.cpp:69-74 default, R-key quaternion step per frame", "coverage": coverage,
                   "parallelism": par,
                   "l2": "explicit flush (160 MB write) before every step; per-step working set = scene %.0f MB + %.0f MB output" % (
                       (64.0 * (n_tri - 1) + 48.0 * n_tri) / 1e6, F * P * 8 / 1e6),
                   "fps": head["fps"], "fps_vs_readme_100fps": head["fps"] / README_FPS, "tree_build_s": build_total, "cpu_affinity": env.numa},
        "e2e": e2e, "gpu_launches": head["launches"], "clocks": head["clocks"], "roofline": roofline, "cpu_baseline": cpu, "reference_gpu": ref_gpu,
        "reference_classes_over_librtb": ref_seam, "frame_loop": loop, "tiles": tiles, "frames_4k": frames_4k, "workloads": others, "scene_extension": scene_ext,
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
    ap.add_argument("--no-other-workloads", action="store_true", help="N=1: skip the other BASELINE.json configurations (`workloads`)")
    ap.add_argument("--no-tiles-leg", action="store_true", help="N>1, frames mode: skip the 4K tiles / peer-push leg (`tiles`)")
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
