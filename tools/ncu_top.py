"""Top stall-sampled SASS instructions of an `ncu --page source --csv` export.  usage: ncu_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
samp = ix["# Samples"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
body = [r for r in rows[2:] if len(r) > samp and r[samp].isdigit()]
tot = sum(int(r[samp]) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][samp]))[:n]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0), reverse=True)[:3]
    print(f"{i:5d} {int(r[samp]):5d} {100*int(r[samp])/tot:5.1f}%  {r[ix['Source']][:90]:90s} {st}")
