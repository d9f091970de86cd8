"""Summarise an .ncu-rep (raw page) into a small text + json for profiles/: duration, DRAM traffic, key rates."""
import csv, json, subprocess, sys
rep, out_prefix = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines())); hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum"]
def num(s, unit):
    v = float(s.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3}.get(unit, 1)
summary = {}
lines = []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].replace("void ", "").strip()
    name = name.split("<")[0]
    d = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            try: d[w] = num(r[i], units[i])
            except ValueError: d[w] = r[i]
    d["dram_bytes_total"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    d["dram_gbs"] = d["dram_bytes_total"] / max(d.get("gpu__time_duration.sum", 1), 1e-9) / 1e3
    summary.setdefault(name, []).append(d)
    lines.append(f"== {name}")
    for k, v in d.items():
        lines.append(f"   {k:72s} {v}")
open(out_prefix + ".txt", "w").write("\n".join(lines) + "\n")
json.dump(summary, open(out_prefix + ".json", "w"), indent=1)
print("\n".join(lines))
