"""ncu_summary.py <report.ncu-rep> <title> — markdown summary of the headline metrics of every captured launch, plus the
wait attribution of scripts/ncu_waits.py.  Also prints the DRAM bytes per launch (for profiles/traffic.json)."""
import csv, io, os, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
command = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, launches = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
print(f"# {title}\n")
if command: print(f"Command: `{command}` (after the same command exited 0 without ncu)\n")
for li, r in enumerate(launches):
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print(f"## launch {li}: {d.get('Kernel Name', '')[:90]}, grid {d.get('launch__grid_size', '?')}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d: print(f"| {k} | {d[k]} | {u.get(k, '')} |")
    try:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tr = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]] + float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
        print(f"\nDRAM traffic (read + write): {tr:.0f} bytes\n")
    except Exception:
        print()
print("## where the warps wait (PC samples, last launch in the report)\n\n```")
print(subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_waits.py"), rep], capture_output=True, text=True).stdout.rstrip())
print("```")
