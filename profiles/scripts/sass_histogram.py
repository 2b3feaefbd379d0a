"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass), the evidence for which hardware paths
the kernels use: DMMA.8x8x4 (FP64 tensor pipe), DFMA/DMUL/DADD (FP64 vector), UBLKCP (bulk async copy), LDS/STS/LDG/STG.
usage: python profiles/scripts/sass_histogram.py [out.txt]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = os.path.join(ROOT, "fiat_b200", "csrc", "libfiat_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ("DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ATOMS")
out = [f"# SASS opcode histogram of fiat_b200/csrc/libfiat_b200.so (sm_100a), {len(kernels)} kernels; cuobjdump -sass",
       "# family counts: " + " ".join(KEY)]
total = collections.Counter()
by_family = collections.defaultdict(collections.Counter)
for name, cnt in kernels.items():
    fam = re.sub(r"<.*", "", demangle(name).replace("void ", ""))
    for op, n in cnt.items():
        by_family[fam][op] += n
        total[op] += n
for fam, cnt in sorted(by_family.items()):
    ninst = sum(cnt.values())
    keys = " ".join(f"{k}={sum(n for op, n in cnt.items() if op.startswith(k))}" for k in KEY)
    out.append(f"{fam:18s} instantiations={sum(1 for n in kernels if re.sub(r'<.*', '', demangle(n).replace('void ', '')) == fam):3d} instructions={ninst:8d}  {keys}")
out.append("# whole library, top opcodes")
for op, n in total.most_common(40):
    out.append(f"{n:9d} {op}")
text = "\n".join(out) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)
