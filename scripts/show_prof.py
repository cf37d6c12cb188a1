"""Print the conv / wgrad rows of a bench.py --profile-json table (optionally only those matching a regex)."""
import json
import re
import sys

d = json.load(open(sys.argv[1]))
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
rows = []
for k, v in d["by_shape"].items():
    if pat is not None and not pat.search(k):
        continue
    rate = v["flops"] / v["ms"] / 1e9 if v["flops"] else v["bytes"] / v["ms"] / 1e6
    rows.append((v["ms"], k, v["calls"], v["ms"] / v["calls"] * 1e3, rate))
rows.sort(reverse=True)
for ms, k, c, us, r in rows[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print("%-66s c=%2d %7.1fus %7.0f %s ms=%.3f" % (k, c, us, r, "TF/s" if "igemm" in k or "wgrad" in k else "GB/s", ms))
print({k: round(v["ms"], 3) for k, v in sorted(d["by_kernel"].items(), key=lambda kv: -kv[1]["ms"])[:8]})
