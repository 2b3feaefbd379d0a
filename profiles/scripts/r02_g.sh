#!/bin/sh
# Round-2 step g: tests of the new cases; store-epilogue probes of the tile kernel; int8 peak; pullback / trace tests.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_g_tests.txt 2>&1
tail -5 gpurun_out/r02_g_tests.txt
q() {  # label env workload flags
  env $2 timeout 300 python bench.py --steps 20 --no-cpu --no-legs --e2e-points 1024 --e2e-steps 1 --workload $3 --flags $4 2>gpurun_out/r02_g_err_$1.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); ev=d.get('e2e_evaluate') or {}
print('$1', '$2', '$3', d['config']['kernel'], round(d['value']/1e9,1), 'Gval/s frac', round(d['roofline']['frac'],3), 'ms/launch', round(d['roofline']['kernel_ms'],4), 'e2e_eval Mpt/s', round(ev.get('value',0)/1e6,1), d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_g_exp.txt
}
: > gpurun_out/r02_g_exp.txt
q p8 A=1 p8_tet_o2 4
q p8_noshfl FIATB200_MMA_SKIP=16 p8_tet_o2 4
q p8_plain FIATB200_MMA_SKIP=32 p8_tet_o2 4
q p8_noST FIATB200_MMA_SKIP=8 p8_tet_o2 4
q n2 A=1 n2curl4_tet_o1 0
q n2_noshfl FIATB200_MMA_SKIP=16 n2curl4_tet_o1 0
q n2_plain FIATB200_MMA_SKIP=32 n2curl4_tet_o1 0
q n2_noST FIATB200_MMA_SKIP=8 n2curl4_tet_o1 0
cat gpurun_out/r02_g_exp.txt
python profiles/microbench/int8_peak.py > gpurun_out/int8_peak.txt 2>&1; cat gpurun_out/int8_peak.txt
