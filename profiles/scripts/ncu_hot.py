"""Hot spots of one kernel from `ncu --page source --csv` (SASS view): samples per instruction group.
python ncu_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
total = sum(int(r[isamp] or 0) for r in body)
print("total samples", total, "instructions", len(body))
agg = {}
for r in body:
    for i in stall:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:top]
for i in sorted(order):
    r = body[i]
    why = sorted(((int(r[j] or 0), hdr[j][6:]) for j in stall), reverse=True)[:3]
    print(f"{i:5d} {int(r[isamp] or 0):7d} {100 * int(r[isamp] or 0) / max(total, 1):5.1f}%  exec {r[iexec]:>10s}  {r[isrc].strip()[:70]:70s} {why}")
