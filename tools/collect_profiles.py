"""After tools/refresh_profiles.sh has run on a GPU box (outputs in gpurun_out/): copy the evidence into profiles/, summarise
the ncu captures (needs the `ncu` CLI), regenerate profiles/traffic.json and print the launch-list shares.
usage: python tools/collect_profiles.py [round=r2]"""
import collections, csv, json, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
src, dst = ROOT / "gpurun_out", ROOT / "profiles"
names = ["bench_c1", "bench_c2", "bench_c3", "bench_c4", "bench_c5", "bench_c2_f32rows", "bench_c4_tiled", "bench_c5_dense",
         "bench_reference_arm"]
for n in names:
    f = src / f"{R}_{n}.json"
    if f.exists() and f.stat().st_size:
        shutil.copy(f, dst / f.name)
for n in ["fused_aug_classes", "fused_aug_empty_input", "fused_aug_dense_scene", "batch_sizes", "strict_signature", "pytest_gpu", "smoke",
          "call_latency", "av2_directory_from_disk", "feed_vit_consumer"]:
    f = src / f"{R}_{n}.txt"
    if f.exists() and f.stat().st_size:
        shutil.copy(f, dst / f.name)
if (src / f"{R}_launches_bench.csv").exists():
    shutil.copy(src / f"{R}_launches_bench.csv", dst / f"{R}_launches_bench.csv")
for cap in ("tiled", "tile_aug", "tile_aug_both"):
    rep = src / f"{R}_prof_{cap}.ncu-rep"
    if rep.exists():
        subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep), str(dst / f"{R}_ncu_full_{cap}")],
                       stdout=subprocess.DEVNULL, check=True)

def load(p):
    return json.load(open(p))

tpath = dst / "traffic.json"
old = load(tpath)
new = {"_note": old["_note"]}
t, a, b = (load(dst / f"{R}_ncu_full_{c}.json") for c in ("tiled", "tile_aug", "tile_aug_both"))
for k in ("k_tile_store", "k_bin_fixed", "k_overflow_red"):
    new[k] = int(t[k][0]["dram_bytes_total"])
for k in ("k_tile_aug", "k_rebin_aug"):
    new[k] = int(a[k][0]["dram_bytes_total"])
new["k_tile_aug_rotate_scale"] = int(b["k_tile_aug"][0]["dram_bytes_total"])
for k in ("k_zero_fill_tma", "k_scatter_atomic"):
    new[k] = old[k]
json.dump(new, open(tpath, "w"), indent=1)
print("traffic.json:", {k: v for k, v in new.items() if k != "_note"})

rows = list(csv.reader(open(dst / f"{R}_launches_bench.csv")))
hdr, agg, cnt = None, collections.Counter(), collections.Counter()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        try:
            v = float(r[hdr.index("Metric Value")].replace(",", ""))
        except ValueError:
            continue
        n = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
        agg[n] += v
        cnt[n] += 1
tot = sum(agg.values())
print("launch list shares (ncu gpu__time_duration.sum):")
for n, v in agg.most_common(8):
    print(f"  {n:36s} {cnt[n]:4d} launches {v / 1e3:10.1f} us  {100 * v / tot:5.1f} %")
for f in sorted(dst.glob(f"{R}_bench_*.json")):
    try:
        j = json.loads(f.read_text().strip().splitlines()[-1])
    except Exception:
        continue
    sw = (j.get("e2e") or {}).get("sliding_window")
    print(f"  {f.name:34s} N={j.get('n_gpus')} value {j['value']:10.1f}  step {j['ms_per_step']:.4f} ms  e2e {j['e2e']['value']:9.1f}"
          + (f"  sliding {sw['value']:9.1f}" if sw else ""))
