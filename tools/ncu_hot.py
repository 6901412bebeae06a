"""Dev tool: top stall sites of one kernel from an ncu source page.  usage: ncu -i X.ncu-rep --page source --csv > f.csv; python tools/ncu_hot.py f.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("total samples", tot, "| by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 50 > tot})
top = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:n]
for i in sorted(top):
    r = body[i]
    s = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {100.0 * s / tot:5.1f}%  {r[ix['Source']].strip():70s} exec={r[ix['Instructions Executed']]:>9s} {why}")
