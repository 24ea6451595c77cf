"""cycles.py name1 name2 ... — A/B builds (variants/libfa_v_<name>.so, or "shipped") by SM CYCLES instead of wall clock:
each CTA of fa::fwdSm100Kernel times itself (clock64, idle warp) into the debug profile buffer; reported are the
slowest CTA's cycles per launch (min / median over FA_CYC_REPS launches) and, from them, cycles per 128-key step of a
CTA.  Cycle counts repeat to ~0.1 %, wall-clock medians on a power-capped B200 do not.
Env: FA_AB_SHAPES = comma list of indices into SHAPES (default "0,1"), FA_CYC_REPS (default 5)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
SHAPES = [(8, 32, 32, 8192, 128, True, "bf16"), (8, 32, 32, 8192, 128, False, "bf16"), (32, 32, 32, 2048, 128, True, "bf16"),
          (8, 32, 32, 8192, 64, True, "bf16"), (64, 32, 32, 1024, 128, True, "bf16"), (4, 12, 12, 1024, 64, False, "fp16"),
          (8, 32, 32, 8192, 128, True, "fp16"), (4, 64, 8, 8192, 128, True, "bf16")]
names = sys.argv[1:]
sel = [int(x) for x in os.environ.get("FA_AB_SHAPES", "0,1").split(",")]
reps = int(os.environ.get("FA_CYC_REPS", "5"))
libs = {}
force = {}
for n in names:      # "name" = variants/libfa_v_<name>.so, "shipped" = the in-tree build; "<lib>@sw,emu" forces a compiled kernel variant
    base, _, fv = n.partition("@")
    fa_b200._lib = None
    fa_b200.LIB_PATH = os.path.join(ROOT, "variants", f"libfa_v_{base}.so") if base != "shipped" else os.path.join(ROOT, "flash-attention-cuda-c_b200", "libfa_b200.so")
    libs[n] = fa_b200.lib()
    libs[n].fa_debug_set_profile_buffer.argtypes = [ctypes.c_void_p]
    force[n] = tuple(int(x) for x in (fv + ",0,0").split(",")[:4]) if fv else None
prof = torch.zeros(32, dtype=torch.int64, device="cuda")
for si in sel:
    B, Hq, Hkv, N, d, causal, dt = SHAPES[si]
    t = {"bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, Hq, N, d, device="cuda", generator=g).to(t); k = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(t); v = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(t)
    o = torch.empty_like(q)
    ref = torch.nn.functional.scaled_dot_product_attention(q[:1, :2, :1024].float(), k[:1, :2, :1024].float(), v[:1, :2, :1024].float(), is_causal=causal) if Hkv == Hq else None
    nq = (N + 255) // 256
    steps = B * Hq * (sum(min((256 * (i + 1) + 127) // 128, (N + 127) // 128) for i in range(nq)) if causal else nq * ((N + 127) // 128))
    for r in range(2):                      # two passes over the builds: shows the repeatability
        for n in names:
            fa_b200._lib = libs[n]
            if force[n] is not None or hasattr(libs[n], "fa_debug_force_variant"):
                try:
                    fv4 = force[n] or (0, 0, 0, 0)
                    libs[n].fa_debug_force_variant(*fv4[:3])
                    if hasattr(libs[n], "fa_debug_force_cta_group"): libs[n].fa_debug_force_cta_group(fv4[3])
                except (AttributeError, TypeError): pass
            libs[n].fa_debug_set_profile_buffer(None)
            fa_b200.attention_forward(q, k, v, causal=causal, out=o)
            err = None
            if ref is not None:
                chk = fa_b200.attention_forward(q[:1, :2, :1024].contiguous(), k[:1, :2, :1024].contiguous(), v[:1, :2, :1024].contiguous(), causal=causal)
                err = round((chk.float() - ref).abs().max().item(), 5)
            libs[n].fa_debug_set_profile_buffer(prof.data_ptr())
            mx, sm = [], []
            for _ in range(reps):
                prof.zero_()
                fa_b200.attention_forward(q, k, v, causal=causal, out=o)
                torch.cuda.synchronize()
                p = prof.cpu().tolist(); mx.append(p[30]); sm.append(p[31])
            libs[n].fa_debug_set_profile_buffer(None)
            ctas = min(148, B * Hq * nq)
            print(json.dumps({"shape": f"N{N}_d{d}_{'c' if causal else 'nc'}_{dt}_B{B}", "lib": n, "pass": r, "max_cta_cycles_min": min(mx), "max_cta_cycles_med": sorted(mx)[len(mx) // 2],
                              "mean_cta_cycles": round(sorted(sm)[len(sm) // 2] / ctas), "cycles_per_step": round(sorted(sm)[len(sm) // 2] / steps, 1), "err": err}), flush=True)
    del q, k, v, o
