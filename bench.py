#!/usr/bin/env python
"""Benchmark of the basis-tabulation hot path (BASELINE.json: "tabulated values/s, P8 tet order-2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

The job is BASELINE's configs[1]: Lagrange P8 on the tetrahedron, all derivatives up to order 2 (1650 float64
values per point), at 10^8 uniform-random points per GPU (weak scaling: every rank tabulates its own contiguous
shard, no collective on the data path).  A *step* is one pass of the hot path over ceil(10^8 / K) of those points,
fed to the kernel in tiles of at most 2^20 points (13.8 GB of output per tile; the whole job's 1.32 TB does not fit
in HBM, so tiles land in a ring of device buffers).  Every tile of every step has its OWN points, generated on the
device before the timed region (2.4 GB of inputs for the job).

Printed JSON line (rank 0): see the contract in the task description.  `value` counts device-resident inputs;
`e2e` goes through the C ABI's host-buffer entry point (host points in, host tables out); `roofline` is the
tabulation kernel against the measured HBM peak (MEASURED_PEAKS.json); `cpu_baseline` is the reference itself
(oracle/_ref, materialised by __graft_entry__.build()) or, without it, its numpy port (oracle/) on the host cores;
`parity` compares the CUDA tables with that CPU pass on the same points, inside this run;
`other_workloads` are short legs for the other BASELINE configurations and the general (non-equispaced) P8 path;
`latency_us` is the per-call latency of the device drop-in at FIAT's real call sizes.
"""
import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy  # noqa: E402

WORKLOADS = {
    # name: (description file, order, cell kind, label)
    "p8_tet_o2": ("p8_tet", 2, "simplex3", "Lagrange P8 tetrahedron, tabulate(order=2)"),
    "p8_spectral_tet_o2": ("p8_spectral_tet", 2, "simplex3", "Lagrange P8 (spectral variant) tetrahedron, tabulate(order=2)"),
    "n2curl4_tet_o1": ("n2curl4_tet", 1, "simplex3", "Nedelec 2nd kind deg 4 tetrahedron, tabulate(order=1)"),
    "hct_o2": ("hct", 2, "simplex2", "HCT triangle, tabulate(order=2)"),
    "ps6_o2": ("ps6", 2, "simplex2", "Powell-Sabin 6 triangle, tabulate(order=2)"),
    "ps12_o2": ("ps12", 2, "simplex2", "Powell-Sabin 12 triangle, tabulate(order=2)"),
    "gll_q10_hex_o1": ("gll_q10_hex", 1, "cube3", "GLL Q10 hexahedron (flattened tensor product), tabulate(order=1)"),
    "p3_tri_o1": ("p3_tri", 1, "simplex2", "Lagrange P3 triangle, tabulate(order=1)"),
}
# points per kernel launch: the split-cell workloads are quoted at 10^7 points (BASELINE.json configs[3]), one launch
DEFAULT_BATCH = {"hct_o2": 10_000_000, "ps6_o2": 10_000_000, "ps12_o2": 10_000_000, "gll_q10_hex_o1": 1 << 20}
JOB_POINTS = {"p8_tet_o2": 100_000_000}       # BASELINE configs[1]: the metric is quoted on 10^8 points
FP64_PEAK_TFLOPS = 37.06      # measured here: profiles/microbench/fp64_peaks.txt (DMMA m8n8k4, B200)
FALLBACK_HBM_GBS = 6650.0


def load_desc(name):
    from fiat_b200 import description
    return description.load(os.path.join(ROOT, "tests", "golden", f"desc_{name}.npz"))


def host_points(kind, n, seed):
    rng = numpy.random.default_rng(seed)
    if kind.startswith("cube"):
        return rng.random((n, int(kind[-1])))
    sd = int(kind[-1])
    u = numpy.sort(rng.random((n, sd)), axis=1)
    return numpy.ascontiguousarray(numpy.diff(numpy.concatenate([numpy.zeros((n, 1)), u], axis=1), axis=1))


def device_points(kind, n, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sd = int(kind[-1])
    u = torch.rand((n, sd), generator=g, device=device, dtype=torch.float64)
    if kind.startswith("cube"):
        return u
    u, _ = torch.sort(u, dim=1)
    return torch.diff(torch.cat([torch.zeros((n, 1), device=device, dtype=torch.float64), u], dim=1), dim=1).contiguous()


def values_per_point(desc, order):
    from fiat_b200 import plan as planmod
    sd = planmod._cell_dim(desc)
    na = len(planmod.alpha_list(sd, order))

    def rows(d):
        if d["kind"] == "simplex":
            return int(d["coeffs"].shape[0] * d["coeffs"].shape[1])
        if d["kind"] == "flattened":
            return rows(d["element"])
        return rows(d["A"]) * rows(d["B"])
    return na * rows(desc), sd


def default_batch(workload, vpp):
    return DEFAULT_BATCH.get(workload) or max(1 << 14, min(1 << 20, int(14e9 // (8 * vpp)) // 4096 * 4096))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=50):
        self.index, self.rows, self.proc, self.period = index, [], None, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", str(self.period),
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        """Summary of the samples taken in [t0, t1] (perf_counter times); all samples if that window is empty."""
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.06] or [r for _, r in self.rows]
        return self._summary(rows)

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)

    def _summary(self, rows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        busy = [v for v in sm if v > 0.5 * (mx[0] if mx else 0)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(rows)}


# ---- CPU side: the reference itself (oracle/_ref) or its numpy port (oracle/) ------------------------------------

def set_blas_threads(n=None):
    """Give the CPU arm every host core whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1).
    -> threads actually in use."""
    n = n or os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return 1


LIVE_ELEMENTS = {
    "p8_tet_o2": lambda F, S: F.Lagrange(S(3), 8),
    "p8_spectral_tet_o2": lambda F, S: F.Lagrange(S(3), 8, variant="spectral"),
    "n2curl4_tet_o1": lambda F, S: F.NedelecSecondKind(S(3), 4),
    "hct_o2": lambda F, S: F.HsiehCloughTocher(S(2)),
    "ps6_o2": lambda F, S: F.QuadraticPowellSabin6(S(2)),
    "ps12_o2": lambda F, S: F.QuadraticPowellSabin12(S(2)),
    "p3_tri_o1": lambda F, S: F.Lagrange(S(2), 3),
}


def cpu_tabulator(workload, desc, order):
    """-> (callable pts -> dict of tables, kind, note).  kind 'reference': the unmodified reference under oracle/_ref
    (FIAT's own element.tabulate); 'port': the numpy restatement in oracle/fiat_oracle.py."""
    try:
        from oracle.make_ref import import_reference
        FIAT = import_reference()
        if FIAT is not None and workload in LIVE_ELEMENTS:
            from FIAT.reference_element import ufc_simplex
            element = LIVE_ELEMENTS[workload](FIAT, ufc_simplex)
            return (lambda pts: element.tabulate(order, pts)), "reference", \
                "FIAT element.tabulate of the reference itself (oracle/_ref), numpy/OpenBLAS"
    except Exception as exc:        # a broken copy must not take the bench down: fall back to the port and say so
        print("live reference unavailable:", exc, file=sys.stderr)
    from oracle import fiat_oracle
    return (lambda pts: fiat_oracle.tabulate(desc, order, pts)), "port", \
        "numpy/OpenBLAS port of the reference algorithm (oracle/fiat_oracle.py)"


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation on the host cores, all threads it can use.

    Every step tabulates a bounded sample of the workload; the sample is sized from a short probe so that the whole
    `--steps K --warmup W` run stays within about two minutes."""
    if rank != 0:
        return
    cores = set_blas_threads()
    dname, order, kind, label = WORKLOADS[args.workload]
    desc = load_desc(dname)
    vpp, _ = values_per_point(desc, order)
    tabulate, cpu_kind, note = cpu_tabulator(args.workload, desc, order)
    probe = host_points(kind, 2000, 98)
    tabulate(probe[:200])
    t0 = time.perf_counter()
    tabulate(probe)
    probe_rate = len(probe) / (time.perf_counter() - t0)                # points/s
    npts = int(probe_rate * 90.0 / max(args.steps + args.warmup, 1))
    npts = max(500, min(args.cpu_points, npts))
    pts = host_points(kind, npts, 99)
    for _ in range(args.warmup):
        tabulate(pts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tabulate(pts)
    dt = time.perf_counter() - t0
    thr = args.steps * npts * vpp / dt
    line = {
        "impl": "reference", "metric": "tabulated values/s", "value": thr, "unit": "values/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "points_per_step": npts, "values_per_point": vpp},
        "cpu_baseline": {"value": thr, "unit": "values/s", "cores": cores, "kind": cpu_kind,
                         "sample": f"{npts} uniform-random points per step x {args.steps} steps; {note}; "
                                   f"BLAS threads pinned to {cores} by this script (host has {os.cpu_count()} CPUs)"},
        "e2e": {"value": thr, "unit": "values/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_and_parity(workload, desc, order, kind, npts, vpp, tab, flags, device):
    """One pass of the CPU arm over `npts` points of the workload, timed; the same points then go through the CUDA
    path and the two results are compared (the oracle as the checker, inside the run)."""
    import torch
    from oracle.tolerance import tolerance
    cores = set_blas_threads()
    tabulate, cpu_kind, note = cpu_tabulator(workload, desc, order)
    pts = host_points(kind, npts, 99)
    tabulate(pts[: max(64, npts // 50)])     # warm-up
    t0 = time.perf_counter()
    want = tabulate(pts)
    secs = time.perf_counter() - t0
    base = {"value": npts * vpp / secs, "unit": "values/s", "cores": cores, "kind": cpu_kind,
            "sample": f"{npts} points of the same workload, one pass ({secs:.1f} s); {note}"}
    worst, ok = 0.0, True
    got = tab.tabulate(order, torch.as_tensor(pts, device=device), flags=flags)
    for alpha, w in want.items():
        g = got[tuple(alpha)].cpu().numpy()
        rel = float(abs(g - w).max() / max(abs(w).max(), 1e-300))
        worst = max(worst, rel)
        ok = ok and rel <= tolerance(desc, alpha)
    parity = {"points": npts, "against": cpu_kind, "max_rel_error": worst, "within_tolerance": bool(ok),
              "tolerance": "1e-12 * max|ref| per derivative table (1e-10 for order >= 2 at degree >= 8)"}
    parity.update(mask_parity(desc, kind, tab, device))
    return base, parity


def mask_parity(desc, kind, tab, device, npts=1 << 16):
    """Split-cell elements: subcell bitmasks of the first 2^16 bench points against the oracle's binning."""
    if desc.get("kind") != "simplex" or int(desc.get("ncells", 1)) < 2:
        return {}
    from oracle import fiat_oracle
    pts = device_points(kind, npts, 1234, device)
    host = pts.cpu().numpy()
    bad = 0
    for unique in (False, True):
        near = fiat_oracle.locate_cells(desc, host, unique=unique)
        want = sum(near[c].astype(numpy.int64) << c for c in range(near.shape[0]))
        mask = tab.locate_subcells(pts, unique).cpu().numpy().astype(numpy.int64)
        bad += int((mask != want).sum())
    return {"mask_points": npts, "mask_mismatches": bad}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs next to its GPU so that the pinned host buffers of the end-to-end
    path are NUMA-local (matters when 8 ranks copy results back at once)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def csrc_stamp(sources):
    """Hash of the kernel's own source files (under fiat_b200/csrc): a profiles/traffic.json entry is only quoted
    while the files its kernel is compiled from are the ones the ncu capture was taken at."""
    h = hashlib.sha256()
    for name in sorted(sources):
        h.update(open(os.path.join(ROOT, "fiat_b200", "csrc", name), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(key):
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        hit = table.get(key)
        if hit and hit.get("stamp") == csrc_stamp(hit["kernel_sources"]):
            return hit.get("bytes")
    except Exception:
        pass
    return None


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def timed_launches(tab, order, tiles, ring, flags, steps, device, barrier=None):
    """`steps` passes over the list of point tiles (device tensors), tile i into ring[i % len(ring)];
    -> (ms per step on the device, kernel launches)."""
    import torch
    from fiat_b200 import _lib
    lib = _lib.load()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.fiatb200_launch_count()
    if barrier:
        barrier()
    e0.record()
    for s in range(steps):
        for i, pts in enumerate(tiles[s % len(tiles)]):
            tab.tabulate_into(ring[i % len(ring)], order, pts, flags=flags)
    e1.record()
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / steps, lib.fiatb200_launch_count() - n0


def run_leg(workload, device, sampler, batch=None, flags=0, steps=10, parity_pts=4096):
    """Short leg for one of the other BASELINE configurations: `steps` launches of one batch (device-resident
    points), its roofline fraction, the kernels it ran on, clocks during the leg and a parity check against the
    oracle on `parity_pts` of its points."""
    import torch
    from fiat_b200 import plan as planmod
    from fiat_b200.api import Tabulator
    from oracle import fiat_oracle
    from oracle.tolerance import tolerance
    dname, order, kind, label = WORKLOADS[workload]
    desc = load_desc(dname)
    vpp, sd = values_per_point(desc, order)
    batch = batch or default_batch(workload, vpp)
    na = len(planmod.alpha_list(sd, order))
    time.sleep(2.5)          # every leg starts from an idle GPU, like the headline (the power cap's running average settles)
    t0 = time.perf_counter()
    tab = Tabulator(desc, device)
    pts = device_points(kind, batch, 4242, device)
    out = torch.empty((na, vpp // na, batch), dtype=torch.float64, device=device)
    tab.tabulate_into(out, order, pts, flags=flags)
    torch.cuda.synchronize(device)
    first_call_s = time.perf_counter() - t0
    for _ in range(3):
        tab.tabulate_into(out, order, pts, flags=flags)
    torch.cuda.synchronize(device)
    reps = max(steps, int(0.25 / max(8e-12 * vpp * batch / 6.5, 1e-6)))      # at least ~0.25 s of device time
    reps = min(reps, 20000)
    w0 = time.perf_counter()
    ms, launches = timed_launches(tab, order, [[pts]], [out], flags, reps, device)
    w1 = time.perf_counter()
    peak, _ = hbm_peak()
    bpp = 8 * vpp + 8 * sd
    achieved = bpp * batch / (ms * 1e-3) / 1e9
    n = min(parity_pts, batch)
    want = fiat_oracle.tabulate(desc, order, pts[:n].cpu().numpy())
    worst, ok = 0.0, True
    for j, (alpha, w) in enumerate(want.items()):
        g = out[j, :, :n].reshape(w.shape).cpu().numpy()
        rel = float(abs(g - w).max() / max(abs(w).max(), 1e-300))
        worst, ok = max(worst, rel), ok and rel <= tolerance(desc, alpha)
    parity = {"points": n, "against": "port", "max_rel_error": worst, "within_tolerance": bool(ok)}
    parity.update(mask_parity(desc, kind, tab, device))
    del out
    torch.cuda.empty_cache()
    return {"workload": label, "points_per_launch": batch, "flags": flags, "value": batch * vpp / (ms * 1e-3),
            "unit": "values/s", "ms_per_launch": ms / max(launches // reps, 1), "launches_per_step": launches // reps,
            "steps": reps, "kernels": tab.kernel_names(order, None, flags), "bytes_per_point": bpp,
            "roofline_frac": achieved / peak, "achieved_gbs": achieved, "first_call_s": first_call_s,
            "clocks": sampler.window(w0, w1), "parity": parity}


def latency_legs(device):
    """Per-call latency of the device drop-in at FIAT's real call sizes (finat/fiat_elements.py:69: 10^1..10^3
    points) on BASELINE configs[0] (Lagrange P3 triangle, order 1): host points in, device tables out, synchronised;
    and the first call (element description -> plan -> first launch)."""
    import torch
    from fiat_b200.api import Tabulator
    dname, order, kind, _ = WORKLOADS["p3_tri_o1"]
    desc = load_desc(dname)
    t0 = time.perf_counter()
    tab = Tabulator(desc, device)
    tab.tabulate(order, host_points(kind, 10, 5))
    torch.cuda.synchronize(device)
    out = {"first_call_ms": (time.perf_counter() - t0) * 1e3}
    for n in (10, 1000, 10000):
        pts = host_points(kind, n, 6)
        for _ in range(20):
            tab.tabulate(order, pts)
        torch.cuda.synchronize(device)
        reps = 200
        t0 = time.perf_counter()
        for _ in range(reps):
            tab.tabulate(order, pts)
        torch.cuda.synchronize(device)
        out[f"host_points_{n}"] = (time.perf_counter() - t0) / reps * 1e6
        dpts = torch.as_tensor(pts, device=device)
        t0 = time.perf_counter()
        for _ in range(reps):
            tab.tabulate(order, dpts)
        torch.cuda.synchronize(device)
        out[f"device_points_{n}"] = (time.perf_counter() - t0) / reps * 1e6
    # first call on the headline element (P8 tet, order 2): 1000 points take the quick plan (thread-per-point kernels,
    # no plan-time optimisation, api.QUICK_NPTS), 2^16 points build the streaming plans (lattice check, packing, ...)
    dname, order, kind, _ = WORKLOADS["p8_tet_o2"]
    desc = load_desc(dname)
    for label, n in (("first_call_ms_p8_tet_o2_1000_points", 1000), ("first_call_ms_p8_tet_o2_65536_points", 1 << 16)):
        pts = host_points(kind, n, 7)
        t0 = time.perf_counter()
        Tabulator(desc, device).tabulate(order, pts)
        torch.cuda.synchronize(device)
        out[label] = (time.perf_counter() - t0) * 1e3
    out["unit"] = "us per call (first_call_ms* in ms)"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fiat_b200", choices=["fiat_b200", "reference"])
    ap.add_argument("--workload", default="p8_tet_o2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="points per kernel launch (0 = workload default)")
    ap.add_argument("--job-points", type=int, default=0, help="points per GPU over all steps (0 = workload default)")
    ap.add_argument("--flags", type=int, default=0, help="kernel selection flags (testing)")
    ap.add_argument("--e2e-points", type=int, default=1 << 16)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-points", type=int, default=100000)
    ap.add_argument("--eval-functions", type=int, default=1, help="functions evaluated by the e2e_evaluate leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / parity leg")
    ap.add_argument("--no-legs", action="store_true", help="skip other_workloads / latency legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from fiat_b200.api import Tabulator
    from fiat_b200 import plan as planmod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    dname, order, kind, label = WORKLOADS[args.workload]
    desc = load_desc(dname)
    vpp, sd = values_per_point(desc, order)
    bytes_per_point = 8 * vpp + 8 * sd
    tile = args.batch or default_batch(args.workload, vpp)
    warmup = max(args.warmup, 3)
    # points of one step on this GPU: the job (10^8 points per GPU for P8) spread over the K steps, in tiles
    job = args.job_points or JOB_POINTS.get(args.workload, 0)
    step_pts = max(tile, -(-job // max(args.steps, 1))) if job else tile
    step_pts = -(-step_pts // 4096) * 4096
    ntiles = -(-step_pts // tile)
    tab = Tabulator(desc, device)
    na = len(planmod.alpha_list(sd, order))
    # inputs resident before the timed region: every tile of every step has its own points (up to 2.4 GB); when the
    # run has more steps than distinct point sets fit in 4 GB the sets are reused cyclically
    nsets = max(1, min(args.steps, int(4e9 // (8 * sd * step_pts))))
    tiles = []
    for s in range(nsets):
        allpts = device_points(kind, step_pts, 1234 + 7919 * rank + 104729 * s, device)
        tiles.append([allpts[i * tile:(i + 1) * tile] for i in range(ntiles)])
    free, _ = torch.cuda.mem_get_info(device)
    ring_n = max(1, min(ntiles, 4, int(0.8 * free // (8 * vpp * tile))))
    ring = [torch.empty((na, vpp // na, tile), dtype=torch.float64, device=device) for _ in range(ring_n)]

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    for w in range(warmup):
        for i, pts in enumerate(tiles[w % nsets]):
            tab.tabulate_into(ring[i % ring_n], order, pts, flags=args.flags)
    barrier()

    # ---- end to end through the host-buffer entry point: host points in, host tables out ----
    ne = min(args.e2e_points, tile)
    hp = torch.empty((ne, sd), dtype=torch.float64, pin_memory=True)
    hp.copy_(tiles[0][0][:ne].cpu())
    ho = torch.empty((na, vpp // na, ne), dtype=torch.float64, pin_memory=True)

    def e2e_run(points, out):
        for _ in range(2):
            tab.tabulate_host(order, points, out=out, chunk_pts=1 << 14, flags=args.flags)
        barrier()
        times = []
        for _ in range(args.e2e_steps):
            t0 = time.perf_counter()
            tab.tabulate_host(order, points, out=out, chunk_pts=1 << 14, flags=args.flags)
            times.append(time.perf_counter() - t0)
        secs = sum(times) / len(times)
        if world > 1:
            t = torch.tensor([secs], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return world * ne * vpp / secs, times

    e2e_value, e2e_times = e2e_run(hp.numpy(), ho.numpy())
    pageable_pts = numpy.array(hp.numpy(), copy=True)
    pageable_out = numpy.empty((na, vpp // na, ne))
    pageable_out.fill(0.0)                       # touch the pages before the timed calls
    e2e_pageable, _ = e2e_run(pageable_pts, pageable_out)
    print("e2e step times (ms):", ["%.1f" % (t * 1e3) for t in e2e_times], file=sys.stderr)

    # ---- fused consumer end to end: host points in, nfunc x npts function values / derivatives out ----
    e2e_eval = None
    if hasattr(tab, "evaluate_host"):
        nfunc = args.eval_functions
        ndofs = planmod.num_dofs_of(desc)
        coef = numpy.random.default_rng(17).standard_normal((nfunc, ndofs))
        neval = min(1 << 20, tile)
        epts = torch.empty((neval, sd), dtype=torch.float64, pin_memory=True)
        epts.copy_(tiles[0][0][:neval].cpu())
        ncomp = vpp // (na * ndofs)
        eout = torch.empty((na, nfunc * ncomp, neval), dtype=torch.float64, pin_memory=True)
        for _ in range(2):
            tab.evaluate_host(coef, order, epts.numpy(), out=eout.numpy())
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            tab.evaluate_host(coef, order, epts.numpy(), out=eout.numpy())
        secs = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([secs], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        e2e_eval = {"value": world * neval / secs, "unit": "points/s", "functions": nfunc, "points_per_step": neval,
                    "tabulated_values_equivalent_per_s": world * neval * vpp / secs,
                    "h2d_bytes_per_step": int(neval * sd * 8 + coef.size * 8), "d2h_bytes_per_step": int(na * nfunc * ncomp * neval * 8),
                    "what": "Tabulator.evaluate_host: u_f = sum_i c[f,i] D^alpha phi_i at host points, host result; "
                            "the (ndofs x npts) tables are never written"}

    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    time.sleep(0.2)
    w0 = time.perf_counter()
    ms_step, launches = timed_launches(tab, order, tiles, ring, args.flags, args.steps, device, barrier)
    w1 = time.perf_counter()
    ms = ms_step * args.steps
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * step_pts * vpp / (ms * 1e-3)

    if rank == 0:
        peak, peak_src = hbm_peak()
        clocks = sampler.window(w0, w1)
        # the roofline line is quoted on all launches of a step together: algorithmic bytes of the step's points over
        # the device time of the step (one launch per tile, or one per derivative table for per-alpha splits)
        per_step = max(launches, 1) / args.steps
        achieved = bytes_per_point * step_pts / (ms / args.steps * 1e-3) / 1e9
        kernel = tab.kernel_path(order, args.flags)
        kernels = tab.kernel_names(order, None, args.flags)
        traffic = measured_traffic(f"{args.workload}|{kernel}|{tile}")
        line = {
            "metric": "tabulated values/s", "value": value, "unit": "values/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": label, "points_per_gpu_per_step": step_pts, "points_per_launch": tile,
                       "launches_per_step": ntiles, "values_per_point": vpp,
                       "total_points": world * args.steps * step_pts, "distinct_point_sets": nsets,
                       "l2": "every launch streams %.1f GB of output through L2 (126 MB) into a ring of %d device buffers; "
                             "inputs (%.0f MB per step, own points for every tile) are >> L2 as well; no separate flush"
                             % (8 * vpp * tile / 1e9, ring_n, 8 * sd * step_pts / 1e6),
                       "sharding": "contiguous point shards, one rank per GPU, no collective",
                       "kernel": kernel, "kernels_per_step": kernels},
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "values/s", "h2d_bytes_per_step": int(ne * sd * 8),
                    "d2h_bytes_per_step": int(ne * vpp * 8), "points_per_step": ne,
                    "buffers": "pinned caller buffers", "pageable_value": e2e_pageable,
                    "pageable_note": "same call with pageable numpy arrays (what tabulate_host allocates by default)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_point": bytes_per_point, "algorithmic_bytes_per_launch": bytes_per_point * tile,
                         "kernel_ms": ms / max(launches, 1), "launches_per_step": per_step, "kernel": kernel,
                         "traffic_source": "profiles/traffic.json (ncu --set full), quoted only while the hash of the kernel's source files matches",
                         "fp64_peak_tflops_measured": FP64_PEAK_TFLOPS},
            "clocks": clocks,
        }
        if e2e_eval:
            line["e2e_evaluate"] = e2e_eval
        del ring, tiles
        torch.cuda.empty_cache()
        if world == 1 and not args.no_cpu:
            base, parity = cpu_baseline_and_parity(args.workload, desc, order, kind, args.cpu_points, vpp, tab,
                                                   args.flags, device)
            line["cpu_baseline"], line["parity"] = base, parity
        if world == 1 and not args.no_legs:
            legs = []
            plan = [("p8_tet_o2", None, 4), ("p8_spectral_tet_o2", None, 0), ("p3_tri_o1", 10000, 0), ("p3_tri_o1", 1 << 20, 0),
                    ("n2curl4_tet_o1", None, 0), ("hct_o2", None, 0), ("ps6_o2", None, 0), ("ps12_o2", None, 0),
                    ("gll_q10_hex_o1", None, 0)]
            for wl, batch, flags in plan:
                if wl == args.workload and flags == args.flags and batch is None:
                    continue
                try:
                    legs.append(run_leg(wl, device, sampler, batch, flags))
                except Exception as exc:       # a failing leg is reported, not hidden
                    legs.append({"workload": wl, "error": f"{type(exc).__name__}: {exc}"})
            line["other_workloads"] = legs
            line["latency_us"] = latency_legs(device)
        sampler.stop()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
