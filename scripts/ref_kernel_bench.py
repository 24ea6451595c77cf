"""ref_kernel_bench.py — the UNMODIFIED reference CUDA kernel (oracle/_ref/libref_kernel.so, built from
/root/reference by oracle/Makefile) beside ours on the shapes it can execute: fp32, B*H = 1, grid = 1.
Prints one JSON object.  Run in its own process under a timeout: the reference kernel contains undefined-behaviour
shuffles and a data race.  Numerics of the reference are only trusted on its all-ones KAT."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch
import fa_b200

so = os.path.join(ROOT, "oracle", "_ref", "libref_kernel.so")
if not os.path.exists(so):
    print(json.dumps({"unavailable": "oracle/_ref/libref_kernel.so not built"})); sys.exit(0)
L = ctypes.CDLL(so)
vp, ip = ctypes.c_void_p, ctypes.c_int
L.ref_kernel_launch.argtypes = [ip, vp, vp, vp, vp, ip, ip, ip, ctypes.c_float, ip, vp]
L.ref_kernel_launch.restype = ip


def ref(variant, q, k, v, scale, causal=False):
    o = torch.zeros_like(q)
    B, H, N, d = q.shape
    rc = L.ref_kernel_launch(variant, q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, scale, int(causal),
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, f"reference launch failed rc={rc}"
    return o


def timeit(fn, iters):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


out = {}
# (i) the reference's own known-answer test, launched exactly as tests/main.cu does
ones = torch.ones(1, 1, 16, 16, device="cuda")
o_ref = ref(0, ones, ones, ones, 0.25)
o_our = fa_b200.attention_forward(ones, ones, ones)
torch.cuda.synchronize()
out["kat_all_ones"] = {"reference_max_abs_dev_from_1": float((o_ref - 1).abs().max()), "ours_max_abs_dev_from_1": float((o_our - 1).abs().max()),
                       "ours_vs_reference_max_abs": float((o_ref - o_our).abs().max())}
# (ii) timing on BASELINE configs[0]: B=1 H=1 N=256 d=64 fp32 (the reference needs grid=1, so one SM)
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(1, 1, 256, 64, device="cuda", generator=g) for _ in range(3))
F = 4.0 * 256 * 256 * 64
ms_ref = timeit(lambda: ref(1, q, k, v, 0.125), 5)
ms_f32 = timeit(lambda: fa_b200.attention_forward(q, k, v), 50)
qb, kb, vb = (t.to(torch.bfloat16) for t in (q, k, v))
ms_b16 = timeit(lambda: fa_b200.attention_forward(qb, kb, vb), 50)
o_r = ref(1, q, k, v, 0.125); o_o = fa_b200.attention_forward(q, k, v); torch.cuda.synchronize()
out["cfg1_fp32_N256_d64"] = {"reference_kernel_ms": ms_ref, "reference_kernel_gflops": F / ms_ref / 1e6,
                             "ours_fp32_ms": ms_f32, "ours_fp32_gflops": F / ms_f32 / 1e6, "ours_bf16_ms": ms_b16,
                             "speedup_fp32_vs_reference_kernel": ms_ref / ms_f32,
                             "reference_vs_ours_max_abs_on_random_data": float((o_r - o_o).abs().max()),
                             "note": "reference numerics on random data are not trusted (scores read the V buffer, SURVEY.md App. A1)"}
# (iii) a longer single-head shape the reference can still run
q, k, v = (torch.randn(1, 1, 2048, 64, device="cuda", generator=g) for _ in range(3))
F = 4.0 * 2048 * 2048 * 64
ms_ref = timeit(lambda: ref(2, q, k, v, 0.125), 2)
ms_f32 = timeit(lambda: fa_b200.attention_forward(q, k, v), 20)
out["fp32_N2048_d64"] = {"reference_kernel_ms": ms_ref, "reference_kernel_gflops": F / ms_ref / 1e6, "ours_fp32_ms": ms_f32,
                         "ours_fp32_gflops": F / ms_f32 / 1e6, "speedup_fp32_vs_reference_kernel": ms_ref / ms_f32}
print(json.dumps(out))
