#!/bin/sh
# Round-2 multi-GPU evidence (gpurun --gpus 2): one point array sharded over the devices of ONE process
# (fiat_b200.tabulate_sharded) and the bench under torchrun with one rank per GPU.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_multi_gpus.txt
timeout 600 python -m pytest tests/test_gpu_bench_parity.py -m gpu -q -k "sharded" > gpurun_out/r02_multi_tests.txt 2>&1
tail -3 gpurun_out/r02_multi_tests.txt
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 1500 gpurun_out/r02_bench_n$N.json
python - <<'PY'
import time, torch, numpy, sys, os
sys.path.insert(0, os.getcwd())
import bench, fiat_b200
desc = bench.load_desc("p8_tet")
n = torch.cuda.device_count()
pts = bench.host_points("simplex3", n * (1 << 18), 5)
hp = torch.as_tensor(pts).pin_memory()
fiat_b200.tabulate_sharded(desc, 2, hp)            # plans on every device
for d in range(n): torch.cuda.synchronize(d)
t0 = time.perf_counter()
for _ in range(5):
    shards = fiat_b200.tabulate_sharded(desc, 2, hp)
for d in range(n): torch.cuda.synchronize(d)
dt = (time.perf_counter() - t0) / 5
print(f"tabulate_sharded: {n} devices, {len(pts)} points, {dt * 1e3:.2f} ms per call, {len(pts) * 1650 / dt / 1e9:.1f} Gval/s in one process")
PY
