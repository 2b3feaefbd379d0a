"""Summaries (profiles/rNN_ncu_*.txt) and profiles/traffic.json from the `ncu --set full` reports pulled into
gpurun_out/ by capture_rN.sh.  traffic = dram__bytes_read.sum + dram__bytes_write.sum over the launches of one
step, keyed workload|kernel path|points per launch like bench.py looks it up; every entry is stamped with the hash of
the source files its kernel is compiled from, as they were when the capture was taken (bench.py quotes an entry only
while that stamp matches: editing another kernel's file does not void it, editing the kernel's does).
usage: python profiles/scripts/make_traffic.py [r02]   (run right after pulling the captures, before editing csrc/)"""
import csv, hashlib, json, os, subprocess, sys
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LATTICE = ["lattice.cuh", "expansion.cuh"]
MMA = ["kernels.cuh", "expansion.cuh"]
VALS = ["vals.cuh", "small.cuh", "expansion.cuh"]
CAPS = {  # report -> (workload, kernel path, batch, source files of the kernel)
    "lattice_p8": ("p8_tet_o2", "lattice", 1 << 20, LATTICE), "mma_p8": ("p8_tet_o2", "simplex", 1 << 20, MMA),
    "mma_p8_spectral": ("p8_spectral_tet_o2", "simplex", 1 << 20, MMA), "lattice_p3": ("p3_tri_o1", "lattice", 1 << 20, LATTICE),
    "cells_walkington": ("walkington_tet_o2", "simplex", 1 << 20, ["cells_reg.cuh", "expansion.cuh"]),
    "mma_n2curl": ("n2curl4_tet_o1", "simplex", 1 << 20, MMA), "vals_hct": ("hct_o2", "simplex", 10_000_000, VALS),
    "vals_ps6": ("ps6_o2", "simplex", 10_000_000, VALS), "vals_ps12": ("ps12_o2", "simplex", 10_000_000, VALS),
    "tensor_hex": ("gll_q10_hex_o1", "tensor", 1 << 20, MMA),
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
src = os.path.join(ROOT, "fiat_b200", "csrc")


def stamp(files):
    h = hashlib.sha256()
    for fname in sorted(files):
        h.update(open(os.path.join(src, fname), "rb").read())
    return h.hexdigest()[:16]


traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step from `ncu --set full` "
                       f"captures (profiles/{TAG}_ncu_*.txt); key = workload|kernel path|points per launch; stamp = hash "
                       "of the kernel's source files at capture time"}
ONLY = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
if ONLY is not None:        # refresh some entries, keep the others as they are
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
for name, (workload, path, batch, files) in CAPS.items():
    if ONLY is not None and name not in ONLY:
        continue
    rep = os.path.join(ROOT, "gpurun_out", f"{TAG}_raw_{name}.csv")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    summary = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "scripts", "ncu_summary.py"), rep],
                             capture_output=True, text=True).stdout
    out = os.path.join(ROOT, "profiles", f"{TAG}_ncu_{name}.txt")
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, the launches of one bench step ({workload}, {batch} points); "
                f"raw metric page: {TAG}_raw_{name}.csv (gpurun_out/, not tracked)\n" + summary)
    raw = open(rep).read()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    total = 0.0
    for r in rows[2:]:
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            total += float(r[i]) * UNIT[units[i]]
    traffic[f"{workload}|{path}|{batch}"] = {"bytes": int(total), "launches": len(rows) - 2,
                                             "source": os.path.relpath(out, ROOT), "kernel_sources": files,
                                             "stamp": stamp(files)}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
