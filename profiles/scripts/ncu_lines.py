"""Hot source lines from `ncu --page source --print-source cuda,sass --csv`: the SASS rows (samples, instructions
executed) are summed per CUDA source line they follow.  python ncu_lines.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg, fname, line, text = {}, None, None, ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0].strip().isdigit():
        line, text = int(r[0]), r[1][:100]
        agg.setdefault((fname, line), [text, 0, 0, {}])
        continue
    if r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
        try:
            samples, instr = int(r[6]), int(r[7])
        except ValueError:
            continue
        a = agg.setdefault((fname, line), [text, 0, 0, {}])
        a[1] += samples
        a[2] += instr
        op = r[3].split()[0] if not r[3].strip().startswith("@") else r[3].split()[1]
        a[3][op.split(".")[0]] = a[3].get(op.split(".")[0], 0) + instr
tot_s = sum(a[1] for a in agg.values()) or 1
tot_i = sum(a[2] for a in agg.values()) or 1
print("total samples", tot_s, "instructions executed", tot_i)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    ops = ",".join(f"{k}:{v * 100 // tot_i}%" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{f:16s} {ln:4d} {100 * a[1] / tot_s:5.1f}% samples {100 * a[2] / tot_i:5.1f}% instr  {a[0][:70]:70s} {ops}")
