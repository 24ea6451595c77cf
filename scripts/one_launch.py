"""one_launch.py <lib.so> B H N d causal [launches] — a warm-up plus a few launches of one build on one shape (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
fa_b200.LIB_PATH = os.path.abspath(sys.argv[1])
B, H, N, d, causal = [int(x) for x in sys.argv[2:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
q, k, v = (torch.randn(B, H, N, d, device="cuda").to(torch.bfloat16) for _ in range(3))
o = torch.empty_like(q)
for _ in range(1 + reps):
    fa_b200.attention_forward(q, k, v, causal=bool(causal), out=o)
torch.cuda.synchronize()
print("ok")
