#!/usr/bin/env python
"""Benchmark of the basis-tabulation hot path (BASELINE.json: "tabulated values/s, P8 tet order-2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch of synthetic points: `--batch` uniform-random
points of the reference tetrahedron per GPU (weak scaling: every rank tabulates its own contiguous
shard, no collective on the data path), Lagrange P8, all derivatives up to order 2
(1650 float64 values per point).  The full 10^8-point job does not fit in HBM (1.32 TB of output),
so the output of every step goes to the same device buffer (13.8 GB per GPU at the default batch,
>> the 126 MB L2, which therefore cannot absorb the stores); the default K = 96 steps of 2^20
points is the 10^8-point job.

Printed JSON line (rank 0): see the contract in the task description.  `value` counts device-resident
inputs; `e2e` goes through the C ABI's host-buffer entry point (pinned host points in, host result
out); `roofline` is the tabulation kernel against the measured HBM peak (MEASURED_PEAKS.json) with
the FP64 picture beside it; `cpu_baseline` is the numpy port of the reference (oracle/) on the host.
"""
import argparse
import json
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
    "n2curl4_tet_o1": ("n2curl4_tet", 1, "simplex3", "Nedelec 2nd kind deg 4 tetrahedron, tabulate(order=1)"),
    "hct_o2": ("hct", 2, "simplex2", "HCT triangle, tabulate(order=2)"),
    "ps6_o2": ("ps6", 2, "simplex2", "Powell-Sabin 6 triangle, tabulate(order=2)"),
    "ps12_o2": ("ps12", 2, "simplex2", "Powell-Sabin 12 triangle, tabulate(order=2)"),
    "gll_q10_hex_o1": ("gll_q10_hex", 1, "cube3", "GLL Q10 hexahedron (flattened tensor product), tabulate(order=1)"),
    "p3_tri_o1": ("p3_tri", 1, "simplex2", "Lagrange P3 triangle, tabulate(order=1)"),
}
DEFAULT_BATCH = {"hct_o2": 10_000_000, "ps6_o2": 10_000_000, "ps12_o2": 10_000_000, "gll_q10_hex_o1": 1 << 20}
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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        busy = [v for v in sm if v > 0.5 * (mx[0] if mx else 0)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_port_throughput(desc, order, kind, npts, vpp, repeats=1):
    """values/s of the numpy port of the reference (oracle/) on this host."""
    from oracle import fiat_oracle
    pts = host_points(kind, npts, 99)
    fiat_oracle.tabulate(desc, order, pts[: max(64, npts // 50)])     # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        fiat_oracle.tabulate(desc, order, pts)
        best = min(best, time.perf_counter() - t0)
    return npts * vpp / best, best


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


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count() or 1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU algorithm (numpy port in oracle/) on the host cores.

    Every step tabulates a bounded sample of the workload; the sample is sized from a short probe so
    that the whole `--steps K --warmup W` run stays within about two minutes.
    """
    if rank != 0:
        return
    from oracle import fiat_oracle
    dname, order, kind, label = WORKLOADS[args.workload]
    desc = load_desc(dname)
    vpp, _ = values_per_point(desc, order)
    probe_pts = 2000
    probe_rate, _ = cpu_port_throughput(desc, order, kind, probe_pts, vpp)          # values/s
    budget_s = 90.0
    npts = int(probe_rate / vpp * budget_s / max(args.steps + args.warmup, 1))
    npts = max(500, min(args.cpu_points, npts))
    pts = host_points(kind, npts, 99)
    for _ in range(args.warmup):
        fiat_oracle.tabulate(desc, order, pts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fiat_oracle.tabulate(desc, order, pts)
    dt = time.perf_counter() - t0
    thr = args.steps * npts * vpp / dt
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "tabulated values/s", "value": thr, "unit": "values/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "points_per_step": npts, "values_per_point": vpp},
        "cpu_baseline": {"value": thr, "unit": "values/s", "cores": cores, "kind": "port",
                         "sample": f"{npts} uniform-random points per step x {args.steps} steps, numpy/OpenBLAS port of "
                                   "the reference algorithm (oracle/fiat_oracle.py); the Python reference itself "
                                   "cannot travel to the GPU box"},
        "e2e": {"value": thr, "unit": "values/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fiat_b200", choices=["fiat_b200", "reference"])
    ap.add_argument("--workload", default="p8_tet_o2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="points per GPU per step (0 = workload default)")
    ap.add_argument("--flags", type=int, default=0, help="kernel selection flags (testing)")
    ap.add_argument("--e2e-points", type=int, default=1 << 16)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-points", type=int, default=100000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from fiat_b200 import _lib
    from fiat_b200.api import Tabulator

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
    # default batch: ~13.8 GB of output per step (P8: 2^20 points); the split-cell workloads are quoted at
    # 10^7 points (BASELINE.json configs[3]), which is one launch and 4.3-5.8 GB of output; the hexahedron
    # streams 2^20 points (44.7 GB) per step (measured: 0.92 of the HBM peak against 0.82 at 327 680 points)
    batch = args.batch or DEFAULT_BATCH.get(args.workload) or \
        max(1 << 14, min(1 << 20, int(14e9 // (8 * vpp)) // 4096 * 4096))
    tab = Tabulator(desc, device)
    from fiat_b200 import plan as planmod
    na = len(planmod.alpha_list(sd, order))
    pts = device_points(kind, batch, 1234 + rank, device)
    out = torch.empty((na, vpp // na, batch), dtype=torch.float64, device=device)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    for _ in range(max(args.warmup, 3)):
        tab.tabulate_into(out, order, pts, flags=args.flags)
    barrier()
    # end to end through the host-buffer entry point: pinned host points in, host result out
    ne = min(args.e2e_points, batch)
    hp = torch.empty((ne, sd), dtype=torch.float64, pin_memory=True)
    hp.copy_(pts[:ne].cpu())
    ho = torch.empty((na, vpp // na, ne), dtype=torch.float64, pin_memory=True)
    for _ in range(2):
        tab.tabulate_host(order, hp.numpy(), out=ho.numpy(), chunk_pts=1 << 14, flags=args.flags)
    barrier()
    e2e_times = []
    for _ in range(args.e2e_steps):
        t0 = time.perf_counter()
        tab.tabulate_host(order, hp.numpy(), out=ho.numpy(), chunk_pts=1 << 14, flags=args.flags)
        e2e_times.append(time.perf_counter() - t0)
    print("e2e step times (ms):", ["%.1f" % (t * 1e3) for t in e2e_times], file=sys.stderr)
    e2e_s = sum(e2e_times) / len(e2e_times)
    if world > 1:
        t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * ne * vpp / e2e_s

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.fiatb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        tab.tabulate_into(out, order, pts, flags=args.flags)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.fiatb200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * batch * vpp / (ms * 1e-3)

    if rank == 0:
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
        except Exception:
            hbm_peak, peak_src = FALLBACK_HBM_GBS, "fallback"
        # a step is one tabulation of the batch: one kernel launch, or one launch per derivative table when
        # the element is split into per-alpha derived elements (plan.alpha_split); the roofline line is
        # quoted on all launches of a step together (algorithmic bytes of the step / time of the step)
        per_step = max(launches, 1) / args.steps
        ms_launch = ms / args.steps
        achieved = bytes_per_point * batch / (ms_launch * 1e-3) / 1e9
        kernel = tab.kernel_path(order, args.flags)
        kernels = tab.kernel_names(order, None, args.flags)
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                f"{args.workload}|{kernel}|{batch}", {}).get("bytes")
        except Exception:
            traffic = None
        line = {
            "metric": "tabulated values/s", "value": value, "unit": "values/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": label, "points_per_gpu_per_step": batch, "values_per_point": vpp,
                       "total_points": world * args.steps * batch,
                       "l2": "every step streams %.1f GB of output through L2 (126 MB), which also evicts the %.0f MB of "
                             "input points between steps; no separate flush" % (8 * vpp * batch / 1e9, 8 * sd * batch / 1e6),
                       "sharding": "contiguous point shards, one rank per GPU, no collective",
                       "kernel": kernel, "kernels_per_step": kernels},
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "values/s", "h2d_bytes_per_step": int(ne * sd * 8),
                    "d2h_bytes_per_step": int(ne * vpp * 8), "points_per_step": ne},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_point": bytes_per_point, "algorithmic_bytes_per_launch": bytes_per_point * batch,
                         "kernel_ms": ms_launch, "launches_per_step": per_step, "kernel": kernel,
                         "fp64_peak_tflops_measured": FP64_PEAK_TFLOPS},
            "clocks": clocks,
        }
        if not args.no_cpu:
            thr, secs = cpu_port_throughput(desc, order, kind, args.cpu_points, vpp)
            line["cpu_baseline"] = {"value": thr, "unit": "values/s", "cores": blas_threads(), "kind": "port",
                                    "sample": f"{args.cpu_points} points of the same workload, one pass ({secs:.1f} s), "
                                              "numpy/OpenBLAS port of the reference algorithm"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
