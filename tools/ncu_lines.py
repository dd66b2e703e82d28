"""Per-source-line shares of executed warp instructions and stall samples from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` (development aid)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = {}
cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[2] != "-":
        continue
    try:
        ln, inst, st = int(r[0]), int(r[7]), int(r[4])
    except ValueError:
        continue
    agg[(cur, ln)] = (inst, st, r[1].strip()[:100])
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("total warp instructions %d, stall samples %d" % (ti, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %4d  inst %5.1f%%  stall %5.1f%%  %s" % (k[0], k[1], 100 * v[0] / ti, 100 * v[1] / ts, v[2]))
