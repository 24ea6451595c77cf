"""ring_profile.py — per-step timing of the ring-KV driver (torchrun, >= 2 GPUs): where does a step's time go?
Usage: torchrun ... scripts/ring_profile.py [sm_reserve]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, torch.distributed as dist
import fa_b200, sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
reserve = int(sys.argv[1]) if len(sys.argv) > 1 else 0
transport = sys.argv[2] if len(sys.argv) > 2 else "auto"
fa_b200.set_sm_reserve(reserve)
B, H, N, d = 1, 8, 131072, 128
n = N // world
q, k, v = (torch.randn(B, H, n, d, device="cuda").to(torch.bfloat16) for _ in range(3))
for _ in range(3): sharding.ring_attention(q, k, v, causal=True, transport=transport)
torch.cuda.synchronize(); dist.barrier()
ms = []
for _ in range(5):
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); sharding.ring_attention(q, k, v, causal=True, transport=transport); b.record(); t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize(); ms.append((a.elapsed_time(b), t_cpu * 1e3))
# compute-only: the same calls without communication
half = n // 2
acc_o = torch.zeros(B, H, half, d, device="cuda"); acc_l = torch.full((B, H, half), float("-inf"), device="cuda")
qa = q[:, :, :half].contiguous()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for s in range(world):
    fa_b200.attention_forward_carry(qa, k[:, :, :half], v[:, :, :half], acc_o, acc_l, causal=(s == 0))
    fa_b200.attention_forward_carry(qa, k, v, acc_o, acc_l, causal=(s == 0)) if s == 0 else fa_b200.attention_forward_carry(qa, k[:, :, :half], v[:, :, :half], acc_o, acc_l)
b.record(); torch.cuda.synchronize()
comp = a.elapsed_time(b)
# comm-only: one hop
kv = torch.stack([k, v]).contiguous(); nxt = torch.empty_like(kv)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, kv, (rank + 1) % world), dist.P2POp(dist.irecv, nxt, (rank - 1) % world)])
for r in reqs: r.wait()
b.record(); torch.cuda.synchronize()
hop = a.elapsed_time(b)
if rank == 0:
    F = 4.0 * B * H * N * N * d / 2
    best = min(m for m, _ in ms)
    print(f"world {world} reserve {reserve} transport {transport}: ring GPU ms {[round(m, 2) for m, _ in ms]} cpu-side ms {[round(c, 2) for _, c in ms]} -> {F / best / 1e9:.0f} TFLOP/s; "
          f"compute-only (same calls, no comm) {comp:.2f} ms; one hop of {kv.numel() * 2 / 2**20:.0f} MiB {hop:.2f} ms ({kv.numel() * 2 / hop / 1e6:.0f} GB/s)", flush=True)
dist.destroy_process_group()
