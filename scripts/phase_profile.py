"""phase_profile.py — run the FA_PHASE_PROFILE build (libfa_b200_prof.so) on one shape and print cycles per phase.
Build: nvcc ... -DFA_PHASE_PROFILE -o flash-attention-cuda-c_b200/libfa_b200_prof.so kernels/FlashAttention.cu"""
import ctypes, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
fa_b200.LIB_PATH = os.path.abspath(os.environ.get("FA_LIB", os.path.join(ROOT, "flash-attention-cuda-c_b200", "libfa_b200_prof.so")))
L = fa_b200.lib()
L.fa_debug_set_profile_buffer.argtypes = [ctypes.c_void_p]
B, H, N, d, causal = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (8, 32, 8192, 128, 1))]
q, k, v = (torch.randn(B, H, N, d, device="cuda").to(torch.bfloat16) for _ in range(3))
fa_b200.attention_forward(q, k, v, causal=bool(causal)); torch.cuda.synchronize()
prof = torch.zeros(32, dtype=torch.int64, device="cuda")
L.fa_debug_set_profile_buffer(prof.data_ptr())
fa_b200.attention_forward(q, k, v, causal=bool(causal)); torch.cuda.synchronize()
p = prof.cpu().tolist()
nq = (N + 255) // 256
tiles = sum(min((N + 127) // 128, ((qb * 256 + 255) // 128 + 1)) if causal else (N + 127) // 128 for qb in range(nq)) * B * H   # kv iterations summed over CTAs
names = ["wait_S", "ld_S", "mask_max_rescale", "exp_pack_st", "st_drain_arrive", "loop_misc"]
sm = {n: p[i] / (8 * tiles) for i, n in enumerate(names)}      # 8 softmax warps per CTA report, per kv iteration
mnames = ["prologue(per CTA, amortised)", "wait_KV", "wait_P", "issue", "wait_S_buffer", "-"]
mma = {f"issuer{t}": {n: round(p[8 + 8 * t + i] / tiles, 1) for i, n in enumerate(mnames)} for t in (0, 1)}
sm = {k: round(x, 1) for k, x in sm.items()}
lag = {f"tile{t}_pickup_after_other_pickup": round(p[24 + t] / max(tiles - B * H * nq, 1), 1) for t in (0, 1)}   # per kv iteration with j > 0
print(json.dumps({"lib": os.path.basename(fa_b200.LIB_PATH), "shape": [B, H, N, d, causal], "kv_iterations": tiles, "softmax_warp_cycles_per_kv_iteration": sm,
                  "softmax_total": sum(sm.values()), "warpgroup_lag_cycles": lag, "mma_thread_cycles_per_kv_iteration": mma, "mma_total": {k: sum(v.values()) for k, v in mma.items()}}))
