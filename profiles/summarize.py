"""Turn the raw ncu captures in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py r01 config3
"""
import collections
import csv
import re
import subprocess
import sys

tag, wl = sys.argv[1], sys.argv[2].replace(":", "_")
src = f"gpurun_out/launches_{tag}_{wl}.csv"
lines = [ln for ln in open(src) if ln.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
tot = 0.0
n = 0
imb = {}                 # launch id -> {metric: value} for the SM-balance columns
all_rows = [row for row in r if len(row) >= len(hdr)]
for row in all_rows:
    if row[idx["Metric Name"]] != "gpu__time_duration.sum":
        imb.setdefault(row[idx["ID"]], {})[row[idx["Metric Name"]]] = float(row[idx["Metric Value"]].replace(",", ""))
for row in all_rows:
    name = row[idx["Kernel Name"]]
    if row[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    val = float(row[idx["Metric Value"]].replace(",", ""))
    unit = row[idx["Metric Unit"]]
    val *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    short = re.sub(r"\(.*", "", re.sub(r"<.*", "", name)).replace("bt::", "").replace("void ", "")
    a = agg.setdefault(short, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += val
    m = imb.get(row[idx["ID"]], {})
    if m.get("sm__cycles_elapsed.max"):
        # time the average SM was NOT busy while the kernel ran (tail / under-filled grid)
        a[2] += val * (1.0 - m.get("sm__cycles_active.avg", 0.0) / m["sm__cycles_elapsed.max"])
    tot += val
    n += 1
with open(f"profiles/{tag}_launches_{wl}.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (profiles/run_ncu.sh {tag} {wl})\n")
    f.write(f"# {n} launches captured (one warm step), serialised: compare SHARES, not absolutes\n")
    f.write(f"# total {tot / 1e6:.3f} ms\n")
    f.write(f"{'kernel':42s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg SM idle ms':>15s}\n")
    for k, (c, v, idle) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:42s} {c:8d} {v / 1e6:10.3f} {100 * v / tot:6.1f}% {idle / 1e6:15.3f}\n")

import os
raw_csv = f"gpurun_out/prof_{tag}_{wl}_raw.csv"
if os.path.exists(raw_csv):
    raw = open(raw_csv).read()
else:
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}_{wl}.ncu-rep", "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active"]
cols = [hdr.index(w) for w in want if w in hdr]
with open(f"profiles/{tag}_ncu_full_{wl}.csv", "w") as f:
    w = csv.writer(f)
    w.writerow([f"{hdr[c]} [{units[c]}]" for c in cols])
    for row in rows[2:]:
        w.writerow([row[c][:60] for c in cols])
print(open(f"profiles/{tag}_launches_{wl}.txt").read())

# per-scope DRAM traffic of one step (bench.py's roofline.traffic): kernel -> bench scope
import json
SCOPE_OF = [("list13_coop_kernel<double, 3, 0>", "l13_walk_count"), ("list13_coop_kernel<double, 3, 1>", "l13_walk_fill"),
            ("list13_coop_kernel<float, 3, 0>", "l13_walk_count"), ("list13_coop_kernel<float, 3, 1>", "l13_walk_fill"),
            ("coll_topdown_kernel", "trav_colleagues_count"), ("list2_masked_fill_kernel", "trav_list2_fill"),
            ("list13_unstage_kernel", "l13_unstage"), ("rs_onesweep_kernel<0, 1>", "l13_heavy_sort_pass"),
            ("rs_onesweep_kernel", "rs_onesweep_pass"), ("permute_kernel", "bt_permute"),
            ("make_keys_kernel", "bt_make_keys"), ("box_extents_own_kernel", "bt_box_extents"),
            ("heavy_map_extract_kernel", "l13_heavy_extract"), ("heavy_map_step_kernel", "l13_heavy_expand"),
            ("heavy_map_hist_kernel", "l13_heavy_expand"), ("heavy_map_rowscan_kernel", "l13_heavy_expand"),
            ("heavy_map_plan_kernel", "l13_heavy_expand"), ("heavy_map_seed_kernel", "l13_heavy_expand"),
            ("coll_compact", "trav_colleagues_fill"), ("list_kernel<double, 3, 4, 0>", "trav_list4_count"),
            ("list_kernel<double, 3, 4, 1>", "trav_list4_fill")]
# scopes whose one call launches several kernels (per level / per step): bytes per CALL
PER_CALL = {"trav_colleagues_count", "l13_heavy_expand"}
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for row in rows[2:]:
    for pat, scope in SCOPE_OF:
        if pat in row[ki]:
            b = float(row[ri].replace(",", "")) * mult.get(units[ri], 1.0) + \
                float(row[wi].replace(",", "")) * mult.get(units[wi], 1.0)
            a = acc.setdefault(scope, [0, 0.0])
            a[0] += 1
            a[1] += b
            break
traffic = {sc: (v[1] if sc in PER_CALL else v[1] / v[0]) for sc, v in acc.items()}
path = f"profiles/{tag}_traffic.json"
try:
    allt = json.load(open(path))
except (OSError, ValueError):
    allt = {}
allt[wl] = {"source": f"ncu --set full, profiles/{tag}_ncu_full_{wl}.csv (dram__bytes_read.sum + dram__bytes_write.sum)",
            "bytes_per_launch": traffic}
json.dump(allt, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(allt[wl], indent=1))
