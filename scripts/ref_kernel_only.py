"""ref_kernel_only.py — times the UNMODIFIED reference CUDA kernel (oracle/_ref/libref_kernel.so, built from
/root/reference by oracle/Makefile) on the shapes it can execute: fp32, B*H = 1, grid = 1 (it never reads blockIdx).
Nothing of this repository's product is imported or loaded here; bench.py's reference arm runs this file in a
subprocess under a timeout (the reference kernel contains undefined-behaviour shuffles and a data race, SURVEY.md
App. A/E) and embeds the one JSON object it prints."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch

so = os.path.join(ROOT, "oracle", "_ref", "libref_kernel.so")
if not os.path.exists(so) or not torch.cuda.is_available():
    print(json.dumps({"unavailable": "oracle/_ref/libref_kernel.so not built" if not os.path.exists(so) else "no GPU"})); sys.exit(0)
L = ctypes.CDLL(so)
vp, ip = ctypes.c_void_p, ctypes.c_int
L.ref_kernel_launch.argtypes = [ip, vp, vp, vp, vp, ip, ip, ip, ctypes.c_float, ip, vp]
L.ref_kernel_launch.restype = ip


def ref(variant, q, k, v, scale):
    o = torch.zeros_like(q)
    B, H, N, d = q.shape
    rc = L.ref_kernel_launch(variant, q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, scale, 0,
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, f"reference launch failed rc={rc}"
    return o


def timeit(fn, iters):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


out = {"kernel": "twoLoaderMhaFlashAttentionKernel (reference kernels/FlashAttention.cuh:59-84, unmodified, sm_100a build), grid 1"}
ones = torch.ones(1, 1, 16, 16, device="cuda")
out["kat_all_ones_max_abs_dev_from_1"] = float((ref(0, ones, ones, ones, 0.25) - 1).abs().max())   # tests/main.cu:33-35 launch contract
g = torch.Generator(device="cuda").manual_seed(0)
for name, variant, N, iters in (("cfg1_fp32_N256_d64", 1, 256, 5), ("fp32_N2048_d64", 2, 2048, 2)):
    q, k, v = (torch.randn(1, 1, N, 64, device="cuda", generator=g) for _ in range(3))
    ms = timeit(lambda: ref(variant, q, k, v, 0.125), iters)
    out[name] = {"ms": ms, "gflops": 4.0 * N * N * 64 / ms / 1e6}
print(json.dumps(out))
