"""Time the pieces of the multi-GPU tile exchange separately (development aid; run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import cpp_cuda_raytracer_dev_b200 as rtb
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); rtb.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, F = 3840, 2160, 12
P = W * H
pts = rtb.geodesic_mesh(233); mesh = rtb.Trixel(pts); mesh.create_kd()
cam = rtb.Camera(W, H, **rtb.default_camera_args(W, H)); obj = rtb.Object(mesh); cam.add_object(obj)
mats = np.stack([obj.matrix()] + [obj.transform_host(rtb.R_KEY_QUAT, rtb.ROTATE_TRI_PY) for _ in range(F - 1)])
PE = cam.tile_major_elements(world)
st = torch.cuda.Stream()
col = torch.empty(F * PE, dtype=torch.int32, device="cuda"); ids = torch.empty(F * PE, dtype=torch.int32, device="cuda")
gc = [torch.empty(F * PE, dtype=torch.int32, device="cuda") for _ in range(world)] if rank == 0 else None
fin = torch.empty(F * P, dtype=torch.int32, device="cuda")
def timed(name, fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps): fn()
        e1.record(st)
    torch.cuda.synchronize()
    if rank == 0: print("%-28s %.3f ms" % (name, e0.elapsed_time(e1) / reps), flush=True)
def render():
    obj.render_frames_device_async(cam, mats, col.data_ptr(), ids.data_ptr(), st.cuda_stream, tile_first=rank, tile_stride=world, flags=rtb.RENDER_TILE_MAJOR)
def gather():
    with torch.cuda.stream(st):
        dist.gather(col, gc, dst=0)
def compose():
    if rank == 0: cam.compose_tiles_device_async(F, [t.data_ptr() for t in gc], fin.data_ptr(), st.cuda_stream)
timed("render tile-major 12 frames", render)
timed("gather 1 buffer", gather)
timed("compose 1 buffer", compose)
def all3():
    render(); gather(); compose()
timed("render+gather+compose serial", all3)
dist.destroy_process_group()
