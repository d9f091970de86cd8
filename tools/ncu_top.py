"""Print the top stall sites of a kernel from an .ncu-rep (source page, SASS view)."""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
kf = ["-k", "regex:" + sys.argv[3], "-c", "1"] if len(sys.argv) > 3 else []
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *kf], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out = []
hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        print("==", r[1][:100]); continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr): out.append(r)
si = hdr.index("# Samples"); src = hdr.index("Source")
names = ["stall_lg", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_barrier", "stall_membar", "stall_drain", "stall_sleep"]
tot = sum(float(r[si] or 0) for r in out) or 1
tots = {k: sum(float(r[hdr.index(k)] or 0) for r in out) for k in names}
print("stall totals %:", {k: round(v / tot * 100, 1) for k, v in tots.items() if v / tot > 0.01})
for r in sorted(out, key=lambda r: -float(r[si] or 0))[:n]:
    why = max(names, key=lambda k: float(r[hdr.index(k)] or 0))
    print(f"{float(r[si]) / tot * 100:5.1f}%  {why:14s} {r[src][:100]}")
