"""symm_bw.py — pull bandwidth from a peer's symmetric-memory buffer with a plain tensor copy (torchrun, >= 2 GPUs)."""
import os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 128 << 20
buf = symm.empty(n, dtype=torch.uint8, device=torch.device("cuda", local))
hdl = symm.rendezvous(buf, dist.group.WORLD)
buf.fill_(rank + 1)
hdl.barrier(channel=0)
prev = (rank - 1) % world
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
for mib in (32, 64, 128):
    m = mib << 20
    src = hdl.get_buffer(prev, (m,), torch.uint8, 0)
    for _ in range(3): dst[:m].copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dst[:m].copy_(src)
    e1.record(); torch.cuda.synchronize()
    ok = bool((dst[:m] == prev + 1).all())
    if rank == 0: print(f"symm pull {mib} MiB: {e0.elapsed_time(e1)/10:.3f} ms = {m / (e0.elapsed_time(e1)/10) / 1e6:.0f} GB/s, data ok {ok}, multicast {hdl.has_multicast_support}", flush=True)
# barrier cost
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): hdl.barrier(channel=1)
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"symm barrier: {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
dist.destroy_process_group()
