#!/bin/sh
# usage: quick_bench.sh "<bench args>" workload [workload...]   -> compact lines in gpurun_out/quick.txt
ARGS="$1"; shift
: > gpurun_out/quick.txt
for w in "$@"; do
  python bench.py --steps 30 --no-cpu --e2e-points 1024 --workload $w $ARGS 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', d['config']['kernel'], '+'.join(sorted(set(d['config']['kernels_per_step']))), len(d['config']['kernels_per_step']), round(d['value']/1e9,1), 'Gval/s frac', round(d['roofline']['frac'],3), 'ms', round(d['ms_per_step'],4))" >> gpurun_out/quick.txt
done
cat gpurun_out/quick.txt
