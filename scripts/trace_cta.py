"""trace_cta.py B H N d causal dtype [cta] — event log of one CTA of a launch (FA_TRACE2 build: scripts/build_variants.sh T2
"-DFA_TRACE2 -DFA_T2_CTA=<cta>" -> variants/libfa_v_T2.so): when the CTA set itself up, published items, issued loads and
Q K^T, took score tiles, delivered P, ran its epilogues.  Clocks relative to kernel entry."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
fa_b200.LIB_PATH = os.path.abspath(os.environ.get("FA_LIB", os.path.join(ROOT, "variants", "libfa_v_T2.so")))
L = fa_b200.lib(); L.fa_debug_set_profile_buffer.argtypes = [ctypes.c_void_p]
B, H, N, d, causal = [int(x) for x in sys.argv[1:6]]
dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[6]]
q, k, v = (torch.randn(B, H, N, d, device="cuda").to(dt) for _ in range(3))
o = torch.empty_like(q)
for _ in range(3): fa_b200.attention_forward(q, k, v, causal=bool(causal), out=o)
prof = torch.zeros(64 + 128 * 6, dtype=torch.int64, device="cuda")
L.fa_debug_set_profile_buffer(prof.data_ptr())
fa_b200.attention_forward(q, k, v, causal=bool(causal), out=o); torch.cuda.synchronize()
L.fa_debug_set_profile_buffer(None)
p = prof.cpu().tolist()
NAMES = {1: "kernel entry", 2: "set-up done", 3: "item published", 4: "Q loads issued", 5: "K/V tile issued", 6: "QK issued slot0", 7: "QK issued slot1",
         10: "S taken slot0", 11: "S taken slot1", 20: "P delivered slot0", 21: "P delivered slot1", 30: "epilogue begins slot0", 31: "epilogue begins slot1",
         35: "last PV retired (epilogue)", 36: "O chunk in registers", 37: "staging piece free", 38: "piece written", 39: "piece visible to TMA", 60: "producer running", 61: "item decoded", 62: "Q buffers free", 40: "epilogue done slot0", 41: "epilogue done slot1", 50: "CTA done",
         70: "S buffer free for QK slot0", 71: "S buffer free for QK slot1", 74: "P half0 ready slot0", 75: "P half1 ready slot0", 76: "P half0 ready slot1", 77: "P half1 ready slot1",
         78: "PV half0 issued slot0", 79: "PV half1 issued slot0", 80: "PV half0 issued slot1", 81: "PV half1 issued slot1"}
ev = sorted(((x & 0xffffffffffff), (x >> 48) & 0xffff) for x in p[64:] if x)
t0 = ev[0][0]
for c, code in ev: print(f"{c - t0:8d}  {NAMES.get(code, code)}")
print("slowest CTA cycles:", p[30])
