"""ncu_waits.py <report.ncu-rep> — where do the warps of fa::fwdSm100Kernel wait?  Reads the source page of an
`ncu --set full --import-source on` capture and sums the PC samples (a) per mbarrier wait loop, named by the barrier's
offset in the shared-memory barrier block, and (b) for everything else in the softmax pass."""
import csv, io, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]       # one block per captured launch: take the last
hi = his[-1]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
names = {}
def bar_name(off, stages=5):
    i = (off & 0xfff) // 8
    tbl = [("q_full", 1), ("q_empty", 1), ("kv_full", stages), ("kv_empty", stages), ("s_full", 2), ("p_full", 4), ("o_full", 2), ("o_free", 2),
           ("sched_full", 2), ("sched_empty", 2), ("s_free", 2), ("o_half", 2)]
    for n, c in tbl:
        if i < c: return n
        i -= c
    return f"bar+{off:#x}"
total = sum(int(r[iS]) for r in data if r[iS].isdigit())
waits = Counter(); other = 0; cur = None; cur_left = 0
for k, r in enumerate(data):
    n = int(r[iS]) if r[iS].isdigit() else 0
    src = r[iSrc]
    m = re.search(r"SYNCS\.PHASECHK\.TRANS64\.TRYWAIT\s+\w+, \[(?:(R\w+)\+)?(UR\w+)(?:\+(0x[0-9a-f]+))?\]", src)
    if m:
        off = int(m.group(3), 16) if m.group(3) else -1
        if m.group(1) and m.group(2) != "URZ" and off >= 0 and (off & 0xfff) == 0:
            cur = "ring barrier (register-indexed: kv_full / kv_empty / sched)"     # base of the barrier block + an index register
        else:
            cur = bar_name(off) if off >= 0 else "?"
        cur_left = 14
    if cur and cur_left > 0 and re.search(r"SYNCS\.PHASECHK|BRA|CS2R|IADD3|IMAD\.X|ISETP|BPT|YIELD|NOP|VIADD|WARPSYNC", src):
        waits[cur] += n
    elif "EXIT" in src or "BAR.SYNC" in src:
        waits["CTA barriers / exit (idle warp)"] += n
    else:
        other += n
    cur_left -= 1
print(f"total samples {total}; per warp (11 warps) {total / 11:.0f}")
for n, c in waits.most_common():
    print(f"  wait {n:28s} {c:8d}  {100 * c / total:5.1f}% of all samples  ({100 * c / (total / 11):6.1f}% of one warp's time)")
print(f"  everything else            {other:8d}  {100 * other / total:5.1f}%")
