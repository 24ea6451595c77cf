"""hop_bw.py — NCCL send/recv ring-hop bandwidth (torchrun, >= 2 GPUs) for a few message sizes."""
import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mib in (32, 64, 128):
    a = torch.empty(mib << 20, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
    for _ in range(3):
        for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, a, (rank + 1) % world), dist.P2POp(dist.irecv, b, (rank - 1) % world)]): r.wait()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, a, (rank + 1) % world), dist.P2POp(dist.irecv, b, (rank - 1) % world)]): r.wait()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{os.environ.get('TAG','')} {mib} MiB hop: {e0.elapsed_time(e1)/10:.3f} ms = {(mib << 20) / (e0.elapsed_time(e1)/10) / 1e6:.0f} GB/s per direction", flush=True)
dist.destroy_process_group()
