"""Summarise an ncu report's source page: top SASS instructions by warp-stall samples with their dominant stall reason.

    python tools/ncu_source_top.py gpurun_out/attn_v2_full.ncu-rep [N]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:200])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for idx, r in enumerate(rows[1:]):
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    tot += n
    st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    data.append((n, idx, r[ix["Source"]].strip(), st, r[ix["Instructions Executed"]]))
print("total samples", tot)
agg = {}
for n, idx, src, st, ie in data:
    for v, c in st:
        pass
# per-reason totals
reason_tot = {c: 0 for c in stall_cols}
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    for c in stall_cols:
        try:
            reason_tot[c] += int(r[ix[c]] or 0)
        except ValueError:
            pass
print("by reason:", sorted(((v, k) for k, v in reason_tot.items() if v), reverse=True)[:10])
for n, idx, src, st, ie in sorted(data, reverse=True)[:topn]:
    print(f"{n:7d} {100.0 * n / max(tot, 1):5.1f}%  #{idx:5d} exec={ie:>9s}  {src[:70]:70s} {st}")
