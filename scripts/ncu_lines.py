"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.

Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > src.csv
       python scripts/ncu_lines.py src.csv [top_n]

Per line: share of warp instructions, share of stall samples and the line's three largest stall reasons.
"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hdr_i]
c_line, c_inst, c_samp = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
src_text = {}
for r in rows[:hdr_i]:
    if len(r) == 2 and r[0].isdigit():
        src_text[int(r[0])] = r[1]
inst, samp = defaultdict(int), defaultdict(int)
stalls = defaultdict(lambda: defaultdict(int))
for r in rows[hdr_i + 1:]:
    if len(r) != len(hdr) or r[0] == "Line No":
        if r and r[0] == "Line No":
            break
        continue
    try:
        ln = int(r[c_line])
    except ValueError:
        continue
    inst[ln] += int(r[c_inst] or 0)
    samp[ln] += int(r[c_samp] or 0)
    for i, name in stall_cols:
        try:
            stalls[ln][name] += int(r[i] or 0)
        except ValueError:
            pass
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp instructions %d, samples %d" % (ti, ts))
tot = defaultdict(int)
for ln in stalls:
    for k, v in stalls[ln].items():
        tot[k] += v
print("stall totals: " + "  ".join("%s %d" % kv for kv in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
for ln, v in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
    why = "  ".join("%s %d" % kv for kv in sorted(stalls[ln].items(), key=lambda kv: -kv[1])[:3] if kv[1])
    print("%5d  inst %5.1f%%  samples %5.1f%%  [%s]  %s" % (ln, 100.0 * inst[ln] / max(ti, 1), 100.0 * v / max(ts, 1), why,
                                                            src_text.get(ln, "").strip()[:90]))
