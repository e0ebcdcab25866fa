"""Fused render + tile exchange over NVLink peer memory (run under torchrun): correctness against the single-GPU frames
and timing beside the NCCL gather + reassembly path.  Development aid / evidence for DESIGN.md section 7."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import cpp_cuda_raytracer_dev_b200 as rtb
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); rtb.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nu, W, H, F, zoom = (int(a) for a in (sys.argv[1:6] or ["233", "3840", "2160", "6", "0"]))
P = W * H
pts = rtb.geodesic_mesh(nu); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
n = cam.basis()[0:3]
for _ in range(zoom):
    obj.transform((float(n[0]), float(n[1]), float(n[2]), 0.005), rtb.TRANSLATE_Z)
mats = np.stack([obj.matrix()] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(F - 1)])
st = torch.cuda.Stream()
# rank 0 owns the final frames; everybody maps them
if rank == 0:
    bufs = [rtb.PeerBuffer(4 * F * P), rtb.PeerBuffer(4 * F * P)]
    handles = [b.handle() for b in bufs]
else:
    handles = [None, None]
dist.broadcast_object_list(handles, src=0)
ptr_c, ptr_i = (bufs[0].ptr, bufs[1].ptr) if rank == 0 else (rtb.peer_open(handles[0]), rtb.peer_open(handles[1]))
flag = torch.zeros(1, dtype=torch.int32, device="cuda")

def push():
    obj.render_frames_push_async(cam, mats, ptr_c, ptr_i, st.cuda_stream, tile_first=rank, tile_stride=world)
    with torch.cuda.stream(st):
        dist.all_reduce(flag)  # all ranks' kernels have finished => every unit has landed on rank 0

# ---- the NCCL path for comparison ----------------------------------------------------------------
PE = cam.tile_major_elements(world)
col = torch.empty(F * PE, dtype=torch.int32, device="cuda"); ids = torch.empty(F * PE, dtype=torch.int32, device="cuda")
gc = [torch.empty(F * PE, dtype=torch.int32, device="cuda") for _ in range(world)] if rank == 0 else None
gi = [torch.empty(F * PE, dtype=torch.int32, device="cuda") for _ in range(world)] if rank == 0 else None
fin_c = torch.empty(F * P, dtype=torch.int32, device="cuda"); fin_i = torch.empty(F * P, dtype=torch.int32, device="cuda")
def nccl():
    obj.render_frames_device_async(cam, mats, col.data_ptr(), ids.data_ptr(), st.cuda_stream, tile_first=rank, tile_stride=world, flags=rtb.RENDER_TILE_MAJOR)
    with torch.cuda.stream(st):
        dist.gather(col, gc, dst=0); dist.gather(ids, gi, dst=0)
    if rank == 0:
        cam.compose_tiles_device_async(F, [t.data_ptr() for t in gc], fin_c.data_ptr(), st.cuda_stream)
        cam.compose_tiles_device_async(F, [t.data_ptr() for t in gi], fin_i.data_ptr(), st.cuda_stream)
def local_only():
    obj.render_frames_device_async(cam, mats, col.data_ptr(), ids.data_ptr(), st.cuda_stream, tile_first=rank, tile_stride=world, flags=rtb.RENDER_TILE_MAJOR)

def timed(name, fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print("%-44s %8.3f ms per %d frames  (%.0f FPS, %.0f Mrays/s)" % (name, float(t[0]), F, F / float(t[0]) * 1e3, F * P / float(t[0]) / 1e3), flush=True)

timed("render own tiles only (no exchange)", local_only)
timed("render + NCCL gather + reassembly", nccl)
timed("fused render + peer push (this work)", push)
# ---- correctness: the pushed frames against this rank's own full single-GPU render ----------------
push(); torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    ref_c = torch.empty(F * P, dtype=torch.int32, device="cuda"); ref_i = torch.empty(F * P, dtype=torch.int32, device="cuda")
    obj.render_frames_device_async(cam, mats, ref_c.data_ptr(), ref_i.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    got_c = np.empty(F * P, np.int32); got_i = np.empty(F * P, np.int32)
    rtb.memcpy_d2h(got_c, ptr_c); rtb.memcpy_d2h(got_i, ptr_i)
    ok = np.array_equal(got_c, ref_c.cpu().numpy()) and np.array_equal(got_i, ref_i.cpu().numpy())
    print("pushed frames == single-GPU frames:", ok, " coverage %.3f" % float((got_i >= 0).mean()), flush=True)
    if not ok:
        print("differing ids:", int((got_i != ref_i.cpu().numpy()).sum()), "colours:", int((got_c != ref_c.cpu().numpy()).sum()))
dist.barrier()
if rank != 0:
    rtb.peer_close(ptr_c); rtb.peer_close(ptr_i)
dist.barrier()
dist.destroy_process_group()
