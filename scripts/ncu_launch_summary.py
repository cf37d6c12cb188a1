"""Summarise an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch,
`--csv`) of one train step: per-kernel totals and the conv family's DRAM traffic per launch.

    python scripts/ncu_launch_summary.py gpurun_out/launches.csv [profiles/rN_conv_traffic.json]"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = {}
for r in data:
    if len(r) <= vi:
        continue
    d = per.setdefault(int(r[0]), {"k": r[ki].split("(")[0].replace("b200::", "").replace("void ", "")})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi].startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)      # -> microseconds
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[mi]] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    a = agg[d["k"][:48]]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("%d launches, %.3f ms of (cold-cache, serialised) kernel time" % (len(per), tot / 1e3))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-50s n=%4d %9.1f us %5.1f%%  dram %8.1f MB read %8.1f MB written" % (k, a[0], a[1], 100 * a[1] / tot, a[2] / 1e6, a[3] / 1e6))
conv = [a for k, a in agg.items() if k.startswith("conv_igemm")]
if conv:
    n = sum(a[0] for a in conv)
    rd, wr, us = sum(a[2] for a in conv), sum(a[3] for a in conv), sum(a[1] for a in conv)
    out = {"kernel": " + ".join(sorted(k for k in agg if k.startswith("conv_igemm"))), "launches_per_step": n,
           "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr, "traffic_bytes_per_launch": (rd + wr) / n,
           "ncu_time_ms_per_step": us / 1e3, "share_of_kernel_time": us / tot, "source": sys.argv[1]}
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            json.dump(out, f, indent=1)
