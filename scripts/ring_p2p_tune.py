"""ring_p2p_tune.py — under torchrun: ring-KV (BASELINE configs[4], N=128K causal, B=1 H=8) with the NCCL send/recv transport,
swept over the number of SMs the persistent attention kernel leaves free for NCCL's copy kernels.  The number of NCCL p2p
channels is fixed per process by NCCL_MIN_P2P_NCHANNELS / NCCL_MAX_P2P_NCHANNELS (set by the caller).  Prints JSON lines on rank 0."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, torch.distributed as dist
import fa_b200, sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, H, N, d = 1, 8, 131072, 128
g = torch.Generator(device=dev).manual_seed(1234)
q, k, v = (torch.randn(B, H, N, d, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
ql, kl, vl = (sharding.zigzag_split(t, world, rank) for t in (q, k, v))
del q, k, v
F = 4.0 * B * H * N * N * d / 2


def timed(fn, iters=5):
    for _ in range(2): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for transport in ("p2p", "peer"):
    for reserve in ((0, 4, 8, 16, 24, 32) if transport == "p2p" else (4,)):
        fa_b200.set_sm_reserve(reserve)
        ms = timed(lambda: sharding.ring_attention(ql, kl, vl, causal=True, transport=transport))
        if rank == 0:
            print(json.dumps({"world": world, "transport": transport, "sm_reserve": reserve, "nccl_p2p_channels": os.environ.get("NCCL_MIN_P2P_NCHANNELS"),
                              "ms_per_pass": round(ms, 3), "tflops": round(F / ms / 1e9, 1)}), flush=True)
fa_b200.set_sm_reserve(0)
dist.barrier(); dist.destroy_process_group()
