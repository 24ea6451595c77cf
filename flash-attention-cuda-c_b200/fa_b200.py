"""fa_b200.py — host-side Python mirror of the reference's operator interface, over the C ABI (include/fa_b200.h).

PyTorch is used only for device memory and streams.  Every call goes through libfa_b200.so (hand-written
sm_100a CUDA); if the library is missing or the device is not a B200 the call raises — there is no fallback.

Mirrors of the reference interface:
  multi_head_attention(Q, K, V, num_heads)      same signature as the reference's check.py:4 ((batch, seq_len,
                                                d_model) tensors); runs on the GPU through fa_fwd_strided, so
                                                check.py's layout needs no transpose.  Returns the output only
                                                (the fused kernel never materialises `attn`).
  two_loader_mha_flash_attention(Q,K,V,O,...)   the kernel's argument list (kernels/FlashAttention.cuh:59-63)
                                                on [B,H,N,d] tensors of any supported dtype.
  attention_forward(q, k, v, ...)               general entry: GQA, causal, LSE, strides taken from the tensors.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("FA_B200_LIB") or os.path.join(_HERE, "libfa_b200.so")   # env override: tuning builds only
SOURCES = [os.path.join(_HERE, "kernels", f) for f in
           ("FlashAttention.cu", "FlashAttention.cuh", "loaders.cuh", "computers.cuh", "utils.cuh")]
NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-shared", "-Xcompiler", "-fPIC"]

FA_DTYPE_F32, FA_DTYPE_F16, FA_DTYPE_BF16 = 0, 1, 2
EXPORTS = ["fa_fwd", "fa_fwd_strided", "fa_fwd_carry", "fa_fwd_carry_window", "fa_mha_fwd_f32", "fa_fwd_host", "fa_merge_partial", "fa_cast_out",
           "fa_workspace_bytes", "fa_set_sm_reserve", "fa_device_info", "fa_block_q", "fa_block_kv", "fa_tile_table", "fa_choose_tile", "fa_choose_kernel", "fa_num_cta", "fa_last_error", "fa_launch_count", "fa_version"]

_lib = None


class FaError(RuntimeError):
    pass


class TileChoice(ctypes.Structure):
    """fa_tile_choice_t of include/fa_b200.h: one row of the measured tile table."""
    _fields_ = [(n, ctypes.c_int) for n in ("d", "causal", "n_min", "block_q", "block_kv", "stages", "softmax_warps",
                                            "emu_pairs_per_8", "staged_epilogue", "issuer_by_type", "cta_group")] + [("tflops", ctypes.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class KernelChoice(ctypes.Structure):
    """fa_kernel_choice_t: what fa_fwd launches for one problem (table row as it runs + the launcher's GQA / small-launch rules)."""
    _fields_ = [("tile", TileChoice), ("heads_per_item", ctypes.c_int), ("work_items", ctypes.c_longlong)]


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc-compile kernels/FlashAttention.cu for sm_100a into libfa_b200.so (in-tree, so it travels with gpurun)."""
    newest = max(os.path.getmtime(s) for s in SOURCES + [os.path.join(_ROOT, "include", "fa_b200.h")])
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SOURCES[0]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise FaError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU or PyTorch fallback for this path)")
        L = ctypes.CDLL(LIB_PATH)
        vp, ip, fl, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
        sigs = {
            "fa_fwd": [vp, vp, vp, vp, vp] + [ip] * 7 + [fl, ip, vp],
            "fa_fwd_strided": [vp, vp, vp, vp, vp] + [ip] * 7 + [fl, ip, ctypes.POINTER(ll), vp],
            "fa_fwd_carry": [vp, vp, vp, vp, vp] + [ip] * 7 + [fl, ip, ctypes.POINTER(ll), vp],
            "fa_fwd_carry_window": [vp, vp, vp, vp, vp] + [ip] * 9 + [fl, ip, ctypes.POINTER(ll), vp],
            "fa_mha_fwd_f32": [vp, vp, vp, vp] + [ip] * 4 + [fl, ip, vp],
            "fa_fwd_host": [vp, vp, vp, vp, vp] + [ip] * 7 + [fl, ip],
            "fa_merge_partial": [vp, vp, vp, vp, ll, ip, ip, vp],
            "fa_cast_out": [vp, vp, ll, ip, vp],
            "fa_device_info": [ip, vp],
            "fa_tile_table": [ctypes.POINTER(ctypes.POINTER(TileChoice))],
            "fa_choose_tile": [ip] * 5 + [ctypes.POINTER(TileChoice)],
            "fa_choose_kernel": [ip] * 8 + [ctypes.POINTER(KernelChoice)],
            "fa_workspace_bytes": [ip] * 7,
            "fa_debug_force_variant": [ip, ip, ip],
            "fa_debug_half_items": [ip],
            "fa_set_sm_reserve": [ip],
            "fa_block_q": [ip, ip], "fa_block_kv": [ip, ip], "fa_num_cta": [ip, ip],
        }
        older = os.environ.get("FA_B200_ALLOW_OLDER_LIB") == "1"     # A/B tools load builds of earlier commits
        for name, args in sigs.items():
            if older and not hasattr(L, name):
                continue
            getattr(L, name).argtypes = args
        for name in EXPORTS:
            if older and not hasattr(L, name):
                continue
            getattr(L, name).restype = ip
        L.fa_last_error.restype = ctypes.c_char_p
        L.fa_version.restype = ctypes.c_char_p
        L.fa_launch_count.restype = ll
        _lib = L
        if os.environ.get("FA_FORCE_VARIANT"):     # tuning / parity runs of one compiled variant: "softmax_warps,emu,staged[,cta_group]"
            sw, emu, epi, cg = (int(x) for x in (os.environ["FA_FORCE_VARIANT"] + ",0,0").split(",")[:4])
            L.fa_debug_force_variant(sw, emu, epi)
            if cg and hasattr(L, "fa_debug_force_cta_group"):
                L.fa_debug_force_cta_group(cg)
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise FaError(f"{what} failed (code {rc}): {lib().fa_last_error().decode()}")


def launch_count() -> int:
    return int(lib().fa_launch_count())


def _dtype_code(t):
    import torch
    return {torch.float32: FA_DTYPE_F32, torch.float16: FA_DTYPE_F16, torch.bfloat16: FA_DTYPE_BF16}[t.dtype]


def _stream_ptr(t):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def attention_forward(q, k, v, causal=False, scale=None, return_lse=False, out=None):
    """q [B,Hq,Nq,d], k/v [B,Hkv,Nk,d] CUDA tensors (any strides with a contiguous last dim). Returns O like q."""
    import torch
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise FaError("attention_forward needs CUDA tensors (no CPU path exists)")
    B, Hq, Nq, d = q.shape
    Bk, Hkv, Nk, dk = k.shape
    if (Bk, dk) != (B, d) or v.shape != k.shape or k.dtype != q.dtype or v.dtype != q.dtype:
        raise FaError("shape / dtype mismatch between q, k, v")
    for t in (q, k, v):
        if t.stride(-1) != 1:
            raise FaError("last dimension must be contiguous")
    if out is None:
        out = torch.empty_like(q, memory_format=torch.contiguous_format) if q.is_contiguous() else torch.empty_like(q)
    lse = torch.empty((B, Hq, Nq), device=q.device, dtype=torch.float32) if return_lse else None
    strides = (ctypes.c_longlong * 12)(*(list(q.stride()[:3]) + list(k.stride()[:3]) + list(v.stride()[:3]) + list(out.stride()[:3])))
    with torch.cuda.device(q.device):
        rc = lib().fa_fwd_strided(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                  lse.data_ptr() if lse is not None else None,
                                  B, Hq, Hkv, Nq, Nk, d, _dtype_code(q), float(scale) if scale else 0.0,
                                  int(bool(causal)), strides, _stream_ptr(q))
    _check(rc, "fa_fwd_strided")
    return (out, lse) if return_lse else out


def two_loader_mha_flash_attention(Q, K, V, O, batchSize, numHeads, seqLen, scale, is_causal):
    """The reference kernel's argument list (kernels/FlashAttention.cuh:59-63) on contiguous [B,H,N,d] CUDA tensors."""
    import torch
    d = Q.numel() // (batchSize * numHeads * seqLen)
    with torch.cuda.device(Q.device):
        rc = lib().fa_fwd(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), None, batchSize, numHeads, numHeads,
                          seqLen, seqLen, d, _dtype_code(Q), float(scale), int(bool(is_causal)), _stream_ptr(Q))
    _check(rc, "fa_fwd")
    return O


def multi_head_attention(Q, K, V, num_heads, causal=False):
    """check.py:4 signature on CUDA tensors: Q, K, V (batch, seq_len, d_model) -> output (batch, seq_len, d_model)."""
    bsz, n, d_model = Q.shape
    dk = d_model // num_heads

    def heads(x):   # a strided view, no copy: [B, N, H, dk] -> [B, H, N, dk]
        return x.view(bsz, n, num_heads, dk).permute(0, 2, 1, 3)

    import torch
    out = torch.empty_like(Q)
    attention_forward(heads(Q), heads(K), heads(V), causal=causal, out=heads(out))
    return out


def attention_forward_host(q, k, v, out, causal=False, scale=None):
    """End-to-end on HOST tensors (pinned for full PCIe speed): H2D, kernel, D2H inside the C library."""
    B, Hq, Nq, d = q.shape
    _, Hkv, Nk, _ = k.shape
    rc = lib().fa_fwd_host(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), None, B, Hq, Hkv, Nq, Nk, d,
                           _dtype_code(q), float(scale) if scale else 0.0, int(bool(causal)))
    _check(rc, "fa_fwd_host")
    return out


def merge_partial(acc_o, acc_lse, part_o, part_lse):
    """Ring-KV carry: fold a 16-bit partial (O, LSE) over a disjoint key range into the fp32 accumulator, in place."""
    import torch
    rows = acc_lse.numel()
    d = acc_o.shape[-1]
    with torch.cuda.device(acc_o.device):
        rc = lib().fa_merge_partial(acc_o.data_ptr(), acc_lse.data_ptr(), part_o.data_ptr(), part_lse.data_ptr(),
                                    rows, d, _dtype_code(part_o), _stream_ptr(acc_o))
    _check(rc, "fa_merge_partial")


def attention_forward_carry(q, k, v, acc_o, acc_lse, causal=False, scale=None, row_offset=0):
    """Ring step: fold attention(q, k, v) over this key range into the fp32 running (acc_o, acc_lse) pair, in place.
    q [B,Hq,Nq,d], k/v [B,Hkv,Nk,d] may be strided views (last dim contiguous); acc_o [B,Hq,R,d] / acc_lse [B,Hq,R] must be
    contiguous with R >= Nq: q's rows are the accumulator's rows [row_offset, row_offset + Nq)."""
    import torch
    B, Hq, Nq, d = q.shape
    _, Hkv, Nk, _ = k.shape
    if not (acc_o.is_contiguous() and acc_lse.is_contiguous() and acc_o.dtype == torch.float32):
        raise FaError("acc_o / acc_lse must be contiguous fp32")
    R = acc_o.shape[2]
    if acc_o.shape != (B, Hq, R, d) or acc_lse.shape != (B, Hq, R):
        raise FaError("acc_o / acc_lse shape mismatch")
    strides = (ctypes.c_longlong * 9)(*(list(q.stride()[:3]) + list(k.stride()[:3]) + list(v.stride()[:3])))
    with torch.cuda.device(q.device):
        rc = lib().fa_fwd_carry_window(q.data_ptr(), k.data_ptr(), v.data_ptr(), acc_o.data_ptr(), acc_lse.data_ptr(), R, int(row_offset),
                                       B, Hq, Hkv, Nq, Nk, d, _dtype_code(q), float(scale) if scale else 0.0,
                                       int(bool(causal)), strides, _stream_ptr(q))
    _check(rc, "fa_fwd_carry_window")


def tile_table():
    """The measured tile table the launcher dispatches from (replaces calculateSizeBlockQ / KV, reference helpers.hpp:8-30)."""
    rows = ctypes.POINTER(TileChoice)()
    n = lib().fa_tile_table(ctypes.byref(rows))
    return [rows[i].as_dict() for i in range(n)]


def choose_tile(d, dtype_code, causal, nq, nk):
    out = TileChoice()
    _check(lib().fa_choose_tile(d, dtype_code, int(bool(causal)), nq, nk, ctypes.byref(out)), "fa_choose_tile")
    return out.as_dict()


def choose_kernel(B, Hq, Hkv, Nq, Nk, d, dtype_code, causal):
    """Exactly what fa_fwd would launch for this problem (fa_choose_kernel): the tile row as it runs, heads per pair item, items."""
    out = KernelChoice()
    _check(lib().fa_choose_kernel(B, Hq, Hkv, Nq, Nk, d, dtype_code, int(bool(causal)), ctypes.byref(out)), "fa_choose_kernel")
    r = out.tile.as_dict()
    r["heads_per_item"] = out.heads_per_item
    r["work_items"] = out.work_items
    return r


def force_variant(softmax_warps=0, emu=0, staged=0, cta_group=0):
    """A/B tooling: run every following launch with this kernel variant (0 = back to the tile table); cta_group 2 = the
    CTA-pair kernel wherever it exists (d = 128, 8 softmax warps)."""
    _check(lib().fa_debug_force_variant(int(softmax_warps), int(emu), int(staged)), "fa_debug_force_variant")
    if hasattr(lib(), "fa_debug_force_cta_group"):
        lib().fa_debug_force_cta_group.argtypes = [ctypes.c_int]
        _check(lib().fa_debug_force_cta_group(int(cta_group)), "fa_debug_force_cta_group")


def set_sm_reserve(sms: int):
    """Leave `sms` SMs free of attention CTAs (for a concurrent NCCL send/recv kernel)."""
    _check(lib().fa_set_sm_reserve(int(sms)), "fa_set_sm_reserve")


def cast_out(src_f32, dst16):
    with __import__("torch").cuda.device(src_f32.device):
        rc = lib().fa_cast_out(src_f32.data_ptr(), dst16.data_ptr(), src_f32.numel(), _dtype_code(dst16), _stream_ptr(src_f32))
    _check(rc, "fa_cast_out")
    return dst16
