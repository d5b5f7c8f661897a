"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share.

    python tools/launch_summary.py gpurun_out/launches_r01.csv [skip_first_n]
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val
    rows.append((name, ms))
rows = rows[skip:]
agg = collections.OrderedDict()
for name, ms in rows:
    short = re.sub(r"\(.*", "", name).replace("void amb::", "")
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.2f} ms total (cold-cache, serialised under ncu: compare SHARES)")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:10.3f} ms {100 * ms / tot:6.2f}%  n={n:5d}  avg={ms / n:8.4f} ms  {k}")
