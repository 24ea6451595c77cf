"""ring_check.py — run under torchrun on >= 2 GPUs: ring-KV (NCCL send/recv) against single-GPU full attention.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ring_check.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, torch.distributed as dist
import fa_b200, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import itertools
ok = True
transports = sys.argv[1:] or ["auto", "p2p"]
for transport, (B, Hq, Hkv, N, d, causal) in itertools.product(transports, [(1, 4, 2, 4096, 128, True), (2, 4, 4, 2048, 128, False), (1, 8, 8, 16384, 128, True), (1, 2, 2, 2048, 64, True)]):
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, Hq, N, d, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(torch.bfloat16)
    full, full_lse = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True)
    if causal:
        ql, kl, vl = (sharding.zigzag_split(t, world, rank) for t in (q, k, v))
        ref = sharding.zigzag_split(full, world, rank); ref_lse = sharding.zigzag_split(full_lse.unsqueeze(-1), world, rank).squeeze(-1)
    else:
        ql, kl, vl = (t.chunk(world, dim=2)[rank].contiguous() for t in (q, k, v))
        ref = full.chunk(world, dim=2)[rank]; ref_lse = full_lse.chunk(world, dim=2)[rank]
    for _ in range(2):   # twice: the second call reuses the symmetric-memory slots of the first
        o, lse = sharding.ring_attention(ql, kl, vl, causal=causal, return_lse=True, transport=transport)
    torch.cuda.synchronize()
    err = (o.float() - ref.float()).abs().max().item(); lerr = (lse - ref_lse).abs().max().item()
    good = err <= 2e-2 and lerr <= 2e-3
    ok &= good
    print(f"rank {rank}/{world} [{transport}] B{B} Hq{Hq} Hkv{Hkv} N{N} d{d} causal={causal}: max|dO|={err:.3e} max|dLSE|={lerr:.3e} {'OK' if good else 'FAIL'}", flush=True)
t = torch.tensor([0 if ok else 1], device="cuda"); dist.all_reduce(t)
dist.destroy_process_group()
if rank == 0: print("RING PASSED" if t.item() == 0 else "RING FAILED")
sys.exit(0 if t.item() == 0 else 1)
