"""NCCL transport / bandwidth probe (development aid): gather and all_gather of 200 MB between ranks."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 50 * 1024 * 1024
x = torch.full((n,), rank, dtype=torch.int32, device="cuda")
outs = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(world)] if rank == 0 else None
big = torch.empty(n * world, dtype=torch.int32, device="cuda")
for name, fn in (("gather", lambda: dist.gather(x, outs, dst=0)), ("all_gather_into_tensor", lambda: dist.all_gather_into_tensor(big, x)),
                 ("send/recv", lambda: (dist.send(x, 0) if rank == 1 else dist.recv(big[:n], 1)) if rank < 2 else None)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 10
    if rank == 0: print("%-24s %.3f ms  %.1f GB/s per sender" % (name, dt * 1e3, n * 4 / dt / 1e9), flush=True)
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
