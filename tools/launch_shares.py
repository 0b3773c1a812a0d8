"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> kernel, launches, total us, share of the listed time, own kernel?
    python tools/launch_shares.py gpurun_out/launches.csv > profiles/rN_launch_shares_timed_region.csv"""
import collections
import csv
import sys

rows = collections.OrderedDict()
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    value = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = value / 1e3 if unit in ("ns", "nsecond") else value if unit in ("us", "usecond") else value * 1e3
    e = rows.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += us
total = sum(v[1] for v in rows.values()) or 1.0
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_us", "share_of_timed_region", "own_kernel"])
for name, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    own = int(any(k in name for k in ("hist_", "minmax", "segmented", "compute_encodings", "reset_kernel", "per_channel", "per_tensor",
                                      "broadcast_", "fold_", "tf_refresh", "ste_bwd", "lg_", "init_range", "entropy_", "pack_kernel")))
    w.writerow([name[:110], n, round(us, 1), round(us / total, 4), own])
