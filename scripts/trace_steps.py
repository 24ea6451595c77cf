"""trace_steps.py [B H N d causal] — timeline of the hand-offs inside CTA 0 (FA_TRACE build: scripts/build_variants.sh TRACE
"-DFA_TRACE" -> variants/libfa_v_TRACE.so, or FA_LIB):
per 128-key step j = 8..23 of the CTA's second work item, when each softmax warpgroup got its S tile, freed the buffer,
started its exponentials and delivered the two halves of P, and when each MMA issuer issued Q K^T and the two halves of
P V.  Times in clocks relative to the first traced event."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
fa_b200.LIB_PATH = os.path.abspath(os.environ.get("FA_LIB", os.path.join(ROOT, "variants", "libfa_v_TRACE.so")))
L = fa_b200.lib()
L.fa_debug_set_profile_buffer.argtypes = [ctypes.c_void_p]
B, H, N, d, causal = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else "8 32 8192 128 0".split())]
q, k, v = (torch.randn(B, H, N, d, device="cuda").to(torch.bfloat16) for _ in range(3))
o = torch.empty_like(q)
for _ in range(2): fa_b200.attention_forward(q, k, v, causal=bool(causal), out=o)
prof = torch.zeros(64 + 4 * 16 * 8, dtype=torch.int64, device="cuda")
L.fa_debug_set_profile_buffer(prof.data_ptr())
fa_b200.attention_forward(q, k, v, causal=bool(causal), out=o)
torch.cuda.synchronize()
L.fa_debug_set_profile_buffer(None)
p = prof.cpu().tolist()
ev = lambda tile, role, j, e: p[64 + ((tile * 2 + role) * 16 + (j - 8)) * 8 + e]
t0 = min(x for x in p[64:] if x > 0)
SM = ["got_S", "freed_S", "exp_start", "P_half0", "P_half1", "ready_for_S"]
MM = ["QK_issue", "QK_issued", "PVh0_issue", "PVh0_issued", "PVh1_issue", "PVh1_issued"]
rows = []
for j in range(8, 24):
    for t in (0, 1):
        for e, n in enumerate(SM):
            if ev(t, 0, j, e): rows.append((ev(t, 0, j, e) - t0, f"softmax{t} j={j} {n}"))
        for e, n in enumerate(MM):
            if ev(t, 1, j, e): rows.append((ev(t, 1, j, e) - t0, f"  issuer{t} j={j} {n}"))
rows.sort()
for c, s in rows: print(f"{c:8d}  {s}")
# per-step summaries
def avg(f):
    xs = [f(j, t) for j in range(10, 22) for t in (0, 1)]
    xs = [x for x in xs if x is not None]
    return round(sum(xs) / len(xs), 1) if xs else None
d = lambda a, b: (a - b) if a and b else None
print(json.dumps({
    "period": avg(lambda j, t: d(ev(t, 0, j + 1, 0), ev(t, 0, j, 0))),
    "wait_for_S": avg(lambda j, t: d(ev(t, 0, j, 0), ev(t, 0, j, 5))),
    "got_S_to_freed": avg(lambda j, t: d(ev(t, 0, j, 1), ev(t, 0, j, 0))),
    "freed_to_exp_start": avg(lambda j, t: d(ev(t, 0, j, 2), ev(t, 0, j, 1))),
    "exp_start_to_P0": avg(lambda j, t: d(ev(t, 0, j, 3), ev(t, 0, j, 2))),
    "P0_to_P1": avg(lambda j, t: d(ev(t, 0, j, 4), ev(t, 0, j, 3))),
    "P1_to_ready": avg(lambda j, t: d(ev(t, 0, j + 1, 5), ev(t, 0, j, 4))),
    "other_freed_S_to_QK_issue": avg(lambda j, t: d(ev(t, 1, j + 1, 0), ev(1 - t, 0, j + (1 if t == 0 else 0), 1)) if True else None),
    "QK_issue_to_got_S": avg(lambda j, t: d(ev(t, 0, j, 0), ev(t, 1, j, 0))),
    "QK_issue_duration": avg(lambda j, t: d(ev(t, 1, j, 1), ev(t, 1, j, 0))),
    "P0_to_PVh0_issue": avg(lambda j, t: d(ev(t, 1, j, 2), ev(t, 0, j, 3))),
    "PVh0_issue_duration": avg(lambda j, t: d(ev(t, 1, j, 3), ev(t, 1, j, 2))),
    "P1_to_PVh1_issue": avg(lambda j, t: d(ev(t, 1, j, 4), ev(t, 0, j, 4))),
    "PVh1_issue_duration": avg(lambda j, t: d(ev(t, 1, j, 5), ev(t, 1, j, 4))),
    "lag_tile1_after_tile0": avg(lambda j, t: d(ev(1, 0, j, 0), ev(0, 0, j, 0)) if t == 0 else None),
}))
