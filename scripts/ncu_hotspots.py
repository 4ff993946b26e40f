"""Off-box: top stall locations of an ncu source page (gpurun_out/<name>_source.csv.gz, SASS view): cumulative samples by
instruction with the dominant stall reason; and totals per stall reason.   python scripts/ncu_hotspots.py <file.csv.gz> [N]"""
import csv
import gzip
import sys

rows = list(csv.reader(gzip.open(sys.argv[1], "rt")))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
by_reason = {h: sum(int(r[idx[h]] or 0) for r in body) for h in stall_cols}
print(rows[0][1][:150])
print("total samples", tot, "| instructions executed", sum(int(r[idx["Instructions Executed"]] or 0) for r in body))
print("by reason:", ", ".join(f"{k[6:]} {100.0 * v / max(tot, 1):.1f}%" for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:9]))
top = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:n]
for i in sorted(top):
    r = body[i]
    s = int(r[idx["# Samples"]] or 0)
    reason = max(stall_cols, key=lambda h: int(r[idx[h]] or 0))
    print(f"{i:6d} {100.0 * s / tot:5.1f}%  {reason[6:]:18s} {r[idx['Source']].strip()[:110]}")
