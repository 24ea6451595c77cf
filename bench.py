#!/usr/bin/env python
"""bench.py — attention forward throughput on B200, the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg4|cfg3nc]

A step is one pass of the hot path (fa_fwd through the C ABI) over one batch of synthetic Q/K/V.
Default workload = BASELINE.json configs[2]: B=8 H=32 N=8192 d=128 causal bf16 (the shape the metric is quoted on).
  value      whole-job TFLOP/s with inputs resident in HBM: the K steps are ONE device-timed region (a CUDA-event pair on
             the launch stream around all of them, launch gaps included), max over ranks
  e2e        same metric through fa_fwd_host: pinned HOST buffers, H2D + kernel + D2H inside the timed region
  roofline   tensor-core bound: algorithmic FLOPs (4*B*Hq*Nq*Nk*d, halved when causal) / mean per-launch kernel time (one
             event pair per step) vs the measured
             cuBLAS bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the torch-CPU restatement of the reference's check.py path timed on this box's host cores on a
             bounded sample of the same workload (rank 0, N=1 only)
N > 1: one process per GPU (torchrun); (batch x head) units are sharded with NO data-path collective — every
rank runs the per-GPU workload on its own units ("weak" scaling); only the timing reduction uses NCCL.
--impl reference times the reference's CPU implementation (check.py path, torch-CPU port in oracle/) instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: whatever NCCL has to say (NCCL_DEBUG=INFO through its logger, NCCL_DEBUG=VERSION
# through a bare printf) goes to stderr — file descriptor 1 points at stderr until the line is printed (emit()).
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)


ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "flash-attention-cuda-c_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (B, Hq, Hkv, N, d, causal, dtype, description)
    "cfg3": (8, 32, 32, 8192, 128, True, "bf16", "Llama-3-8B attention shape B=8 H=32 N=8192 d=128 causal bf16 (BASELINE configs[2])"),
    "cfg3nc": (8, 32, 32, 8192, 128, False, "bf16", "B=8 H=32 N=8192 d=128 non-causal bf16"),
    "cfg2": (4, 12, 12, 1024, 64, False, "fp16", "GPT-2 shape B=4 H=12 N=1024 d=64 non-causal fp16 (BASELINE configs[1])"),
    "cfg4": (16, 64, 8, 32768, 128, True, "bf16", "GQA Hq=64 Hkv=8 N=32K B=16 d=128 causal bf16 (BASELINE configs[3]); per GPU: B=16/N ranks"),
    "cfg5": (1, 8, 8, 131072, 128, True, "bf16", "long context N=128K d=128 causal bf16, B=1 H=8, sequence-sharded ring-KV (BASELINE configs[4]); per GPU: N/ranks rows, zig-zag"),
}
STRONG = {"cfg4", "cfg5"}   # total work fixed as ranks grow; cfg2/cfg3 replicate the per-GPU workload (weak)
METRIC = "attention fwd TFLOP/s (bf16, d=128, N=8K causal), whole job; roofline.frac = fraction of measured dense bf16 TC peak"


def flops(B, Hq, Nq, Nk, d, causal):
    f = 4.0 * B * Hq * Nq * Nk * d
    return f / 2 if causal else f


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms from the last warm-up steps through the timed region
    (the timed region of the default run is < 0.1 s, so the sampler is started during warm-up to be sure it is
    already delivering samples under load when timing starts)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        load = [c for c, w in zip(sm, power) if w >= 0.5 * max(power)] if power else []
        return {"sm_mhz": statistics.median(load) if load else (statistics.median(sm) if sm else None),
                "sm_mhz_min_under_load": min(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "samples_under_load": len(load),
                "reasons": sorted(reasons)}


def cpu_attention_sample(N, d, causal, heads, torch):
    """The check.py path (torch-CPU port, oracle.multi_head_attention) on `heads` heads of the workload, fp32."""
    from oracle import oracle
    g = torch.Generator().manual_seed(0)
    Q = torch.randn(1, N, heads * d, generator=g)
    K = torch.randn(1, N, heads * d, generator=g)
    V = torch.randn(1, N, heads * d, generator=g)
    t0 = time.perf_counter()
    out, _ = oracle.multi_head_attention(Q, K, V, heads, causal=causal)
    dt = time.perf_counter() - t0
    return dt, float(out.abs().mean())


def run_reference(args, wl):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B, Hq, Hkv, N, d, causal, dtype, desc = wl
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    heads = 2 if N >= 8192 else min(Hq, 12)
    t1, _ = cpu_attention_sample(N, d, causal, 1, torch)   # calibrate: keep the whole run within a few minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    heads = max(1, min(heads * 4, int(budget / max(t1, 1e-3)), Hq))
    for _ in range(args.warmup):
        cpu_attention_sample(N, d, causal, heads, torch)
    times = [cpu_attention_sample(N, d, causal, heads, torch)[0] for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    val = flops(1, heads, N, N, d, causal) / (ms * 1e-3) / 1e12
    sample = f"{heads} of {B * Hq} (batch, head) slices of the workload per step, fp32, [N,N] scores materialised (check.py path)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def run_ours(args, wl, wl_name):
    import torch
    import fa_b200
    B, Hq, Hkv, N, d, causal, dtype, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if wl_name == "cfg4":
        B = max(1, B // world)   # 128 kv groups sharded by batch: strong-scaled config, 16/N batches per GPU
    ring = wl_name == "cfg5"
    N_total = N
    if ring:
        import sharding
        N = N // world           # this rank's rows (two zig-zag chunks of N/(2*world))
        if world > 1:
            fa_b200.set_sm_reserve(4)   # room for the one-CTA barrier kernels (peer transport) / NCCL kernels (p2p)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16}[dtype]
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(rank)
    q = torch.randn(B, Hq, N, d, device=dev, generator=g).to(tdt)
    k = torch.randn(B, Hkv, N, d, device=dev, generator=g).to(tdt)
    v = torch.randn(B, Hkv, N, d, device=dev, generator=g).to(tdt)
    o = torch.empty_like(q)
    F = flops(B, Hq, N, N, d, causal)
    if ring:
        F = flops(B, Hq, N_total, N_total, d, causal) / world   # this rank's share of the whole-sequence work
    es = 2
    alg_bytes = (2 * B * Hq * N * d + 2 * B * Hkv * N * d) * es
    flush = None
    if alg_bytes < 256 << 20:   # working set near L2 size: flush L2 between timed iterations
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        if ring:
            sharding.ring_attention(q, k, v, causal=causal)
        else:
            fa_b200.attention_forward(q, k, v, causal=causal, out=o)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = fa_b200.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for a, b in evs:
        if flush is not None:
            flush.zero_()
        a.record()
        step()
        b.record()
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = fa_b200.launch_count() - launches0
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    # the K steps as ONE device-timed region (first launch to last retirement, gaps between launches included); with an L2
    # flush between steps the flush writes are not part of a step, so the per-step event times are summed instead
    total_ms = ev0.elapsed_time(ev1) if flush is None else sum(kernel_ms)
    clocks = sampler.stop() if rank == 0 else None
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = float(tt.item()) / args.steps
    value = F * world / (ms_per_step * 1e-3) / 1e12

    # ---- end to end: pinned host buffers through fa_fwd_host, copies inside the timed region --------------
    e2e = None
    e2e_steps = max(1, min(args.steps, 3))
    try:
        if ring:
            raise RuntimeError("ring-KV keeps Q/K/V sharded and resident on the GPUs; the host-buffer path is measured on cfg3")
        hq, hk, hv = (t.cpu().pin_memory() for t in (q, k, v))
        ho = torch.empty_like(hq).pin_memory()
        fa_b200.attention_forward_host(hq, hk, hv, ho, causal=causal)   # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fa_b200.attention_forward_host(hq, hk, hv, ho, causal=causal)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = (hq.numel() + hk.numel() + hv.numel()) * es
        d2h = ho.numel() * es
        e2e = {"value": F * world / float(te.item()) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": float(te.item()) * 1e3, "steps": e2e_steps,
               "api": "fa_fwd_host (C ABI, pinned host buffers, 3-stream H2D/kernel/D2H pipeline)",
               "result_check": float(ho[0, 0, :4].float().abs().mean())}
    except Exception as ex:   # report, never hide
        e2e = {"value": None, "error": str(ex)}

    if rank == 0:
        burst, sustained, hbm, how = peaks()
        k_ms = sum(kernel_ms) / len(kernel_ms)          # this rank's mean per-step time from the per-step event pairs
        per_gpu = F / (k_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(wl_name)
            except Exception:
                traffic = None
        roofline = {"bound": "tensor", "achieved": per_gpu, "peak": burst, "unit": "TFLOP/s", "frac": per_gpu / burst,
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({how}, burst: kernel timed alone)",
                    "frac_of_sustained": per_gpu / sustained, "frac_of_nominal_2250": per_gpu / 2250.0,
                    "traffic": traffic, "algorithmic_bytes": alg_bytes, "algorithmic_flops": F,
                    "hbm_gbs_achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                    "kernel": "fa::fwdSm100Kernel", "kernel_ms": k_ms}
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            t1, _ = cpu_attention_sample(N, d, causal, 1, torch)
            heads = max(1, min(Hq * B, int(12.0 / max(t1, 1e-3))))
            dt, _ = cpu_attention_sample(N, d, causal, heads, torch)
            cpu = {"value": flops(1, heads, N, N, d, causal) / dt / 1e12, "unit": "TFLOP/s", "cores": torch.get_num_threads(),
                   "kind": "port", "seconds": dt,
                   "sample": f"{heads} of {B * Hq} (batch, head) slices, fp32, torch-CPU port of check.py:4-25 (scores materialised)"}
        line = {"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if wl_name in STRONG else "weak", "vs_baseline": None,
                "dtype": dtype, "data": "synthetic",
                "config": {"workload": desc, "per_gpu": {"B": B, "Hq": Hq, "Hkv": Hkv, "N": N, "d": d, "causal": causal},
                           "sharding": ("sequence-sharded ring-KV, zig-zag causal layout: one K/V hop per step overlapped with the MMAs; transport "
                                        + ("symmetric-memory peer pull over NVLink (copy on a side stream), NCCL send/recv as fallback"
                                           if (world > 1 and sharding._PEER_RINGS) else "NCCL send/recv")
                                        if ring else "(batch x head) units per rank, no data-path collective"),
                           "l2": ("working set %.2f GiB > 126 MB L2" % (alg_bytes / 2**30)) if flush is None else "L2 flushed (256 MiB write) between timed iterations",
                           "timing": ("one CUDA-event pair around the K steps on the launch stream (launch gaps included)" if flush is None
                                      else "CUDA events per step on the launch stream, summed (the L2 flush between steps is not timed)") + "; max over ranks"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "wall_s_timed_region": t_wall, "kernel_ms_min": min(kernel_ms), "kernel_ms_max": max(kernel_ms)}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_ours(args, wl, args.workload)


if __name__ == "__main__":
    sys.exit(main())
